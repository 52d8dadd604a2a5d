"""Per-opcode executed-instruction mix and stall samples from an `ncu --page source --csv` export (scratch tool)."""
import collections
import csv
import sys

rd = csv.reader(open(sys.argv[1]))
next(rd)
hdr = next(rd)
ci = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); smp = collections.Counter(); wav = collections.Counter()
tot = 0
for row in rd:
    if len(row) < len(hdr):
        continue
    ins = row[ci["Source"]].split()
    if not ins:
        continue
    op = ins[0] if not ins[0].startswith("@") else ins[1]
    op = ".".join(op.split(".")[:3]) if op.startswith("F2F") else op.split(".")[0]
    n = int(row[ci["Instructions Executed"]])
    ex[op] += n; tot += n
    smp[op] += int(row[ci["# Samples"]])
    wav[op] += int(row[ci["L1 Wavefronts Shared"]] or 0)
base = ex["DMMA"] / 11.0   # 8x8 blocks at d = 20 (11 DMMAs each); pass another divisor as argv[2]
if len(sys.argv) > 2:
    base = ex["DMMA"] / float(sys.argv[2])
print("total warp-instructions %.3e; per 8x8 block %.1f" % (tot, tot / base))
ts = sum(smp.values())
for op, n in ex.most_common(28):
    print("%-14s %6.1f per block  %5.1f %% of instr  %5.1f %% of stall samples  smem wavefronts/block %.1f" % (op, n / base, 100.0 * n / tot, 100.0 * smp[op] / ts, wav[op] / base))
