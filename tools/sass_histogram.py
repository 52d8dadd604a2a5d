"""SASS evidence (no GPU needed): per-kernel instruction histogram of libmcmcgpu.so from `cuobjdump -sass`, and registers /
spills / shared memory from `nvcc -Xptxas -v` on every translation unit.  Writes profiles/sass_r02.txt.

What to look for: DMMA (FP64 tensor-core MMA; tcgen05 has no FP64 kind, so DMMA.8x8x4 is the tensor path of this problem),
UBLKCP (cp.async.bulk: 1-D bulk TMA on pre-packed tile images -- no tensor map, hence no UTMALDG), SYNCS (mbarrier ops)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mcmc.jl_b200", "libmcmcgpu.so")
CSRC = os.path.join(ROOT, "mcmc.jl_b200", "csrc")
KEYS = ["DMMA", "UBLKCP", "UTMALDG", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "LDG", "STG", "STS", "ATOM", "RED", "BAR", "SHFL", "IMAD", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    hist, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); hist[cur] = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur:
            op = m.group(1); full = op + (m.group(2) or "")
            hist[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k):
                    hist[cur][k] += 1
            if op == "DMMA":
                hist[cur]["shape:" + full] += 1
    dm = demangle(list(hist))
    regs = {}
    for src, flags in [("k1_regress.cu", []), ("transition.cu", ["-fmad=false"]), ("fused_chain.cu", ["-fmad=false"]), ("stats.cu", ["-fmad=false"]),
                       ("population.cu", ["-fmad=false"]), ("zv.cu", ["-fmad=false"]), ("engine.cu", [])]:
        p = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v", "-c",
                            os.path.join(CSRC, src), "-o", "/dev/null"] + flags, capture_output=True, text=True, cwd=CSRC)
        name = None
        for line in p.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = m.group(1); regs[name] = {}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and name:
                regs[name].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", line)
            if m and name:
                regs[name].update(regs=int(m.group(1)), smem=int(m.group(2) or 0))
    out = [__doc__.strip(), "", "kernel | instructions | " + " | ".join(KEYS) + " | regs | spill st/ld (B) | static smem (B)"]
    for fn, h in hist.items():
        if h["total"] < 8:
            continue
        r = regs.get(fn, {})
        out.append(" | ".join([dm[fn][:110], str(h["total"])] + [str(h[k]) for k in KEYS] +
                              [str(r.get("regs", "?")), f"{r.get('spill_st', '?')}/{r.get('spill_ld', '?')}", str(r.get("smem", "?"))]))
        shapes = {k: v for k, v in h.items() if k.startswith("shape:")}
        if shapes:
            out.append("    " + ", ".join(f"{k[6:]} x{v}" for k, v in shapes.items()))
    tot = collections.Counter()
    for h in hist.values():
        tot.update({k: h[k] for k in KEYS})
    out += ["", "library totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]), f"UTMALDG (tensor-map TMA) occurrences: {tot['UTMALDG']}"]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", "sass_r02.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-3:]))


if __name__ == "__main__":
    main()
