// probe 2: does DMMA share the FP64 pipe with DFMA? what does an LDS.64 per DMMA cost?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: DMMA only (regs). 1: DMMA + LDS.64 per DMMA. 2: DMMA(regs) + NF DFMA per DMMA. 3: LDS + DFMA
template <int MODE, int NF>
__global__ void __launch_bounds__(256, 2) k(double* out, int iters, double a, double bb) {
  __shared__ double sm[32 * 108];
  for (int i = threadIdx.x; i < 32 * 108; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double* p = sm + g * 108 + t;
  double c[16], f[8];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8; i++) f[i] = 1.0 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double b = bb;
      if (MODE == 1 || MODE == 3) b = p[(i * 4 + (it & 3) * 32) ];
      dmma884(c[2 * i], c[2 * i + 1], a, b);
      if (MODE >= 2) {
#pragma unroll
        for (int q = 0; q < NF; q++) f[(i * NF + q) & 7] = fma(f[(i * NF + q) & 7], a, bb);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
#pragma unroll
  for (int i = 0; i < 8; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize()); float best = 1e30f;
  for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}
int main() {
  int nsm = 148; double* out; CK(cudaMalloc(&out, 8 * nsm * 2 * 256 * 4));
  const int iters = 20000; int blocks = nsm * 2, threads = 256; double nw = (double)blocks * threads / 32;
  float ms;
#define RUN(M, NF, name) ms = timeit([&] { k<M, NF><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }); \
  printf("%-28s DMMA %.2f TFLOP/s  (+DFMA %.2f TFLOP/s)\n", name, nw * 8 * 512.0 * iters / ms / 1e9, (M >= 2) ? nw * 8 * NF * 64.0 * iters / ms / 1e9 : 0.0);
  RUN(0, 0, "dmma regs");
  RUN(1, 0, "dmma + lds64");
  RUN(2, 1, "dmma + 1 dfma");
  RUN(2, 2, "dmma + 2 dfma");
  RUN(2, 4, "dmma + 4 dfma");
  RUN(2, 8, "dmma + 8 dfma");
  RUN(3, 1, "dmma + lds + 1 dfma");
  RUN(3, 2, "dmma + lds + 2 dfma");
  return 0;
}
