"""Measure cuBLAS FP64 GEMM throughput (the FP64 roofline denominator) on the current GPU."""
import json, sys, torch
def measure(n=8192, reps=10):
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    for _ in range(2): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2.0 * n ** 3 / best / 1e9
    # sustained: back to back for ~3 s
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    k = max(3, int(3000 / best)); e0.record()
    for _ in range(k): torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    sus = 2.0 * n ** 3 * k / e0.elapsed_time(e1) / 1e9
    return {"fp64_dgemm_tflops": burst, "fp64_dgemm_tflops_sustained": sus, "n": n}
if __name__ == "__main__":
    print(json.dumps(measure(int(sys.argv[1]) if len(sys.argv) > 1 else 8192)))
