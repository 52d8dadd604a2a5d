"""Compressed opcode-class sequence of one kernel's SASS (scratch tool): D = FP64, M = DMMA, F = FP32, X = F2F, L = LDS, G = LDG, . = other."""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False; seq = []
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if not m:
        continue
    op = m.group(2)
    c = "M" if op == "DMMA" else "D" if op in ("DFMA", "DMUL", "DADD", "DSETP") else "X" if op == "F2F" else \
        "F" if op in ("FFMA", "FMUL", "FADD") else "L" if op == "LDS" else "G" if op == "LDG" else "B" if op in ("BRA", "BSSY", "BSYNC") else "."
    seq.append(c)
s = "".join(seq)
for i in range(0, len(s), 160):
    print(s[i:i + 160])
