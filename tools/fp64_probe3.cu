// probe 3: what do double <-> float conversions cost, alone and next to DMMA / DFMA?  (K1's probit link converts 8 values per element)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// per iteration: ND DMMAs, NF DFMAs, NC double->float->double round trips (2 conversions each), NS FFMAs
template <int ND, int NF, int NC, int NS, int NI = 0>
__global__ void __launch_bounds__(256, 2) k(double* out, int iters, double a, double bb) {
  double c[16], f[8], v[8];
  float s[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8; i++) { f[i] = 1.0 + i; v[i] = 1.0 + 1e-3 * i + threadIdx.x; s[i] = 1.0f + i; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ND; i++) dmma884(c[2 * (i & 7)], c[2 * (i & 7) + 1], a, bb);
#pragma unroll
    for (int q = 0; q < NF; q++) f[q & 7] = fma(f[q & 7], a, bb);
#pragma unroll
    for (int q = 0; q < NC; q++) { float x = (float)v[q & 7]; x = x * 1.0000001f; v[q & 7] = (double)x; }
#pragma unroll
    for (int q = 0; q < NS; q++) s[q & 7] = fmaf(s[q & 7], 1.0000001f, 1e-9f);
#pragma unroll
    for (int q = 0; q < NI; q++) u[q & 7] = (u[q & 7] ^ (u[(q + 1) & 7] >> 3)) + 0x9E3779B9u;
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r += c[i];
#pragma unroll
  for (int i = 0; i < 8; i++) r += f[i] + v[i] + s[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize()); float best = 1e30f;
  for (int r = 0; r < 3; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}
int main() {
  int nsm = 148; double* out; CK(cudaMalloc(&out, 8 * nsm * 2 * 256 * 4));
  const int iters = 20000; int blocks = nsm * 2, threads = 256;
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  float ms;
  // cycles per iteration per SM sub-partition: 16 warps per SM = 4 per sub-partition
#define RUN(ND, NF, NC, NS) RUNI(ND, NF, NC, NS, 0)
#define RUNI(ND, NF, NC, NS, NI) ms = timeit([&] { k<ND, NF, NC, NS, NI><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }); \
  printf("int-ops %2d DMMA %2d DFMA %2d conv-pairs %2d FFMA %2d : %7.1f cycles per iteration per warp-slot (4 warps share a sub-partition: x4 = pipe cycles)\n", 2 * NI, ND, NF, NC, NS, ms * 1e-3 * clk * 1e3 / iters / 4.0);
  RUN(8, 0, 0, 0);
  RUN(0, 32, 0, 0);
  RUN(0, 0, 8, 0);
  RUN(0, 0, 16, 0);
  RUN(0, 0, 0, 64);
  RUN(8, 0, 8, 0);
  RUN(8, 0, 16, 0);
  RUN(0, 32, 8, 0);
  RUN(0, 32, 16, 0);
  RUN(8, 32, 16, 64);
  RUN(8, 32, 0, 64);
  RUN(8, 0, 0, 64);
  RUN(8, 0, 0, 128);
  RUNI(0, 0, 0, 0, 32);
  RUNI(8, 0, 0, 0, 32);
  RUNI(8, 0, 0, 0, 64);
  RUN(4, 0, 0, 64);
  RUN(2, 0, 0, 64);
  return 0;
}
