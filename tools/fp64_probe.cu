// FP64 throughput probe for B200 (sm_100a): DFMA vs DMMA shapes, and libm transcendentals.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(c[2 * i], c[2 * i + 1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void k_dmma16816(double* out, int iters, double av, double bv) {
  double c[16], a[8], b[4];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = av + i;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = bv + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) dmma16816(c + 4 * i, a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__global__ void k_dmma1688(double* out, int iters, double av, double bv) {
  double c[16], a[4], b[2];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = av + i;
  b[0] = bv; b[1] = bv + 1;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) dmma1688(c + 4 * i, a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// transcendental throughput (results/s): 4 independent chains per thread
template <int OP>
__global__ void k_trans(double* out, int iters, double x0) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = x0 + 1e-3 * threadIdx.x + 0.1 * i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (OP == 0) x[i] = exp(-x[i]) + 0.5;            // exp
      if (OP == 1) x[i] = log(x[i]) + 2.0;             // log
      if (OP == 2) x[i] = 1.0 / (1.0 + x[i]);          // div
      if (OP == 3) x[i] = erfc(x[i]) + 0.5;            // erfc
      if (OP == 4) x[i] = sqrt(x[i]) + 1.0;            // sqrt
      if (OP == 5) x[i] = log1p(x[i]) + 0.5;           // log1p
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms=%d cc=%d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 32 * 1024));
  const int iters = 20000;
  for (int wpb : {4, 8, 16, 32}) {
    for (int bps : {1, 2}) {
      if (wpb * bps > 64) continue;
      int threads = wpb * 32, blocks = nsm * bps;
      double nthr = (double)threads * blocks;
      float ms;
      ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk=%2d blk/sm=%d  DFMA      %.2f TFLOP/s\n", wpb, bps, nthr * 16 * 2.0 * iters / ms / 1e9);
      ms = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk=%2d blk/sm=%d  DMMA884   %.2f TFLOP/s\n", wpb, bps, (nthr / 32) * 8 * 256 * 2.0 * iters / ms / 1e9);
      ms = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk=%2d blk/sm=%d  DMMA1688  %.2f TFLOP/s\n", wpb, bps, (nthr / 32) * 4 * 1024 * 2.0 * iters / ms / 1e9);
      ms = timeit([&] { k_dmma16816<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf("warps/blk=%2d blk/sm=%d  DMMA16816 %.2f TFLOP/s\n", wpb, bps, (nthr / 32) * 4 * 2048 * 2.0 * iters / ms / 1e9);
    }
  }
  {
    int threads = 256, blocks = nsm * 8; double nthr = (double)threads * blocks; int it2 = 2000;
    const char* names[] = {"exp", "log", "div", "erfc", "sqrt", "log1p"};
    float ms;
    ms = timeit([&] { k_trans<0><<<blocks, threads>>>(out, it2, 0.3); }); printf("%-6s %.1f Gop/s\n", names[0], nthr * 4 * it2 / ms / 1e6);
    ms = timeit([&] { k_trans<1><<<blocks, threads>>>(out, it2, 0.3); }); printf("%-6s %.1f Gop/s\n", names[1], nthr * 4 * it2 / ms / 1e6);
    ms = timeit([&] { k_trans<2><<<blocks, threads>>>(out, it2, 0.3); }); printf("%-6s %.1f Gop/s\n", names[2], nthr * 4 * it2 / ms / 1e6);
    ms = timeit([&] { k_trans<3><<<blocks, threads>>>(out, it2, 0.3); }); printf("%-6s %.1f Gop/s\n", names[3], nthr * 4 * it2 / ms / 1e6);
    ms = timeit([&] { k_trans<4><<<blocks, threads>>>(out, it2, 0.3); }); printf("%-6s %.1f Gop/s\n", names[4], nthr * 4 * it2 / ms / 1e6);
    ms = timeit([&] { k_trans<5><<<blocks, threads>>>(out, it2, 0.3); }); printf("%-6s %.1f Gop/s\n", names[5], nthr * 4 * it2 / ms / 1e6);
  }
  return 0;
}
