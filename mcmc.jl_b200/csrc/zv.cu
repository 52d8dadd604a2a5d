// zv.cu -- zero-variance control variates on the device (src/stats/zv.jl:8-66).
// One CTA per chain: feature means, the covariance block [cov(features) | cov(features, x_i)] accumulated over the
// stored draws and gradients in the reference's order (two-pass, 1/(n-1)), a cooperative Gauss-Jordan elimination
// with partial pivoting in shared memory (the reference calls inv(); a = -inv(C) Sigma), then zv = x + features * a.
// Every matrix entry is owned by one thread and updated with the same operations in the same order as the CPU
// restatement, so the coefficients agree bit for bit.  Compiled with -fmad=false.
#include "zv.h"

namespace mg {

int64_t zv_features(int64_t d, int order) { return order == 1 ? d : d * (d + 3) / 2; }
bool zv_supported(int64_t d, int order) {
  if (order != 1 && order != 2) return false;
  const int64_t k = zv_features(d, order);
  return d >= 1 && (size_t)(k * (k + d) + (k + d)) * sizeof(double) <= 200 * 1024;
}

__device__ __forceinline__ double zv_feature(int64_t p, int64_t d, const double* x, const double* g, int64_t st) {
  // zv.jl:16,48-56: z = -grad/2; [z, 2*z.*x - 1, x_i*z_j + x_j*z_i (i<j)]; x[j*st], g[j*st] are this draw's entries
  if (p < d) return (-g[p * st]) / 2.0;
  if (p < 2 * d) { int64_t j = p - d; return (2.0 * ((-g[j * st]) / 2.0)) * x[j * st] - 1.0; }
  int64_t l = p - 2 * d, i = 0;
  while (l >= d - 1 - i) { l -= d - 1 - i; i++; }
  int64_t j = i + 1 + l;
  return x[i * st] * ((-g[j * st]) / 2.0) + x[j * st] * ((-g[i * st]) / 2.0);
}

__global__ void __launch_bounds__(256) zv_kernel(const double* __restrict__ samples, const double* __restrict__ grads,
                                                 int64_t S, int64_t d, int64_t C, int64_t Cp, int order, double* zv_out,
                                                 double* a_out, int32_t* status) {
  extern __shared__ double sm[];
  const int64_t c = blockIdx.x;
  const int64_t k = (order == 1) ? d : d * (d + 3) / 2, w = k + d;
  double* mean = sm;            // [w]
  double* M = sm + w;           // [k][w]
  __shared__ int piv_row;
  __shared__ int singular;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int64_t dst = d * Cp;   // stride between consecutive draws
  const double* xs = samples + c;
  const double* gs = grads + c;
  if (tid == 0) singular = 0;
  for (int64_t p = tid; p < w; p += nth) {
    double s = 0.0;
    for (int64_t t = 0; t < S; t++)
      s += (p < k) ? zv_feature(p, d, xs + t * dst, gs + t * dst, Cp) : xs[t * dst + (p - k) * Cp];
    mean[p] = s / (double)S;
  }
  __syncthreads();
  for (int64_t e = tid; e < k * w; e += nth) {                  // cov([features x]) (zv.jl:19,58)
    const int64_t p = e / w, q = e % w;
    double s = 0.0;
    for (int64_t t = 0; t < S; t++) {
      const double fp = zv_feature(p, d, xs + t * dst, gs + t * dst, Cp) - mean[p];
      const double fq = ((q < k) ? zv_feature(q, d, xs + t * dst, gs + t * dst, Cp) : xs[t * dst + (q - k) * Cp]) - mean[q];
      s += fp * fq;
    }
    M[e] = s / (double)(S - 1);
  }
  __syncthreads();
  for (int64_t col = 0; col < k; col++) {                       // Gauss-Jordan, partial pivoting (zv.jl:20-22,59-61)
    if (tid == 0) {
      int64_t piv = col; double best = fabs(M[col * w + col]);
      for (int64_t r = col + 1; r < k; r++) if (fabs(M[r * w + col]) > best) { best = fabs(M[r * w + col]); piv = r; }
      piv_row = (int)piv;
      if (!(best > 0.0)) singular = 1;
    }
    __syncthreads();
    if (singular) break;
    const int64_t piv = piv_row;
    if (piv != col)
      for (int64_t q = tid; q < w; q += nth) { double tmp = M[col * w + q]; M[col * w + q] = M[piv * w + q]; M[piv * w + q] = tmp; }
    __syncthreads();
    const double pv = M[col * w + col];
    __syncthreads();
    for (int64_t q = tid; q < w; q += nth) M[col * w + q] = M[col * w + q] / pv;
    __syncthreads();
    // every other row: M[r][q] -= M[r][col] * M[col][q]; the multipliers are read before anything is overwritten
    for (int64_t e = tid; e < k * w; e += nth) {
      const int64_t r = e / w, q = e % w;
      if (r == col || q == col) continue;
      M[e] = M[e] - M[r * w + col] * M[col * w + q];
    }
    __syncthreads();
    for (int64_t r = tid; r < k; r += nth) if (r != col) M[r * w + col] = M[r * w + col] - M[r * w + col] * 1.0;   // column -> 0 (pivot is 1)
    __syncthreads();
  }
  if (tid == 0) status[c] = singular;
  if (singular) return;
  for (int64_t e = tid; e < k * d; e += nth) {                  // a = -precision * sigma
    const int64_t p = e / d, i = e % d;
    a_out[e * Cp + c] = -M[p * w + k + i];
  }
  if (zv_out) {                                                 // zvChain = x + features * a (zv.jl:25,63)
    for (int64_t e = tid; e < S * d; e += nth) {
      const int64_t t = e / d, i = e % d;
      double s = 0.0;
      for (int64_t p = 0; p < k; p++) s += zv_feature(p, d, xs + t * dst, gs + t * dst, Cp) * (-M[p * w + k + i]);
      zv_out[(t * d + i) * Cp + c] = xs[t * dst + i * Cp] + s;
    }
  }
}

cudaError_t launch_zv(const double* samples, const double* grads, int64_t S, int64_t d, int64_t C, int64_t Cp, int order,
                      double* zv_out, double* a_out, int32_t* status, cudaStream_t st) {
  const int64_t k = zv_features(d, order);
  size_t smem = sizeof(double) * (size_t)(k * (k + d) + (k + d));
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(zv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  zv_kernel<<<(unsigned)C, 256, smem, st>>>(samples, grads, S, d, C, Cp, order, zv_out, a_out, status);
  return cudaGetLastError();
}

}  // namespace mg
