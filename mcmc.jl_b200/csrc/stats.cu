// stats.cu -- K3: the src/stats estimators on the device-resident draws.
//
// One thread per (chain, parameter) series; the draws are chain-minor ([kept][param][chain]) so a
// warp's loads at a fixed draw index are one coalesced 256-byte segment.  Restated:
//   mean              src/stats/mean.jl:6          sum / n
//   mcvar_iid         src/stats/var.jl:7-8         unbiased variance (two-pass) / n
//   mcvar_bm          src/stats/var.jl:20-26       batch means
//   mcvar_imse/ipse   src/stats/var.jl:45-75,95-116  Geyer initial monotone / positive sequence over the
//                     demeaned, 1/n-normalised autocovariance (StatsBase.acf(...; correlation=false))
//   ess / actime      src/stats/ess.jl:6-19
// The reference computes all n lags (O(n^2)) and then reads only those before the first non-positive
// pair sum; here lags are produced LAGS_PER_PASS at a time from a register window and the scan stops
// at the truncation point, which gives the same value with O(n * m) work.
// Compiled with -fmad=false (sums in the reference's order).
#include "stats.h"

namespace mg {

constexpr int LW = 16;  // lags per pass (8 Geyer pairs)

__global__ void __launch_bounds__(128) stats_kernel(const double* __restrict__ samples, int64_t S, int64_t d, int64_t C,
                                                    int64_t Cp, int vtype, int64_t maxlag, int64_t batchlen,
                                                    double* mean_o, double* viid_o, double* var_o, double* ess_o,
                                                    double* act_o) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (c >= C) return;
  const double* x = samples + j * Cp + c;     // x[t] at x[t * d * Cp]
  const int64_t st = d * Cp;
  const double n = (double)S;
  // mean (mean.jl:6) and variance (Base.var: two-pass, n-1)
  double s = 0.0;
#pragma unroll 8
  for (int64_t t = 0; t < S; t++) s += x[t * st];            // loads are independent: 8 in flight per thread
  const double mu = s / n;
  double ss = 0.0;
#pragma unroll 8
  for (int64_t t = 0; t < S; t++) { double v = x[t * st] - mu; ss += v * v; }
  const double viid = (ss / (double)(S - 1)) / n;   // var.jl:7-8
  double v = CUDART_NAN;
  if (vtype == MCMCGPU_VAR_IID) {
    v = viid;
  } else if (vtype == MCMCGPU_VAR_BM) {
    // var.jl:20-26
    const int64_t nb = S / batchlen;
    if (nb > 1) {
      double sm = 0.0;
      for (int64_t b = 0; b < nb; b++) {
        double bs = 0.0;
        for (int64_t t = 0; t < batchlen; t++) bs += x[(b * batchlen + t) * st];
        sm += bs / (double)batchlen;
      }
      const double mb = sm / (double)nb;
      double sv = 0.0;
      for (int64_t b = 0; b < nb; b++) {
        double bs = 0.0;
        for (int64_t t = 0; t < batchlen; t++) bs += x[(b * batchlen + t) * st];
        double dv = bs / (double)batchlen - mb;
        sv += dv * dv;
      }
      v = (double)batchlen * (sv / (double)(nb - 1)) / (double)(nb * batchlen);
    }
  } else {
    // Geyer IMSE / IPSE (var.jl:45-75 / :95-116)
    const int64_t k = (maxlag - 1 >= 0) ? (maxlag - 1) / 2 : -1;   // floor((maxlag-1)/2)
    if (k >= 0) {
      const bool monotone = (vtype == MCMCGPU_VAR_IMSE);
      double acv0 = 0.0, gsum = 0.0, gprev = 0.0;
      int64_t jj = 0;          // Geyer pair index
      bool stop = false;
      for (int64_t lag0 = 0; !stop && jj <= k; lag0 += LW) {
        // autocovariances at lags lag0 .. lag0+LW-1 in one pass over the series
        double a[LW];
#pragma unroll
        for (int l = 0; l < LW; l++) a[l] = 0.0;
        // window w[l] = x[t + lag0 + l] - mu, slid along t
        double w[LW];
#pragma unroll
        for (int l = 0; l < LW; l++) { int64_t idx = lag0 + l; w[l] = (idx < S) ? x[idx * st] - mu : 0.0; }
        for (int64_t t = 0; t + lag0 < S; t++) {
          const double xt = x[t * st] - mu;
#pragma unroll
          for (int l = 0; l < LW; l++) a[l] += xt * w[l];   // terms beyond the end are exact zeros
#pragma unroll
          for (int l = 0; l + 1 < LW; l++) w[l] = w[l + 1];
          int64_t nx = t + lag0 + LW;
          w[LW - 1] = (nx < S) ? x[nx * st] - mu : 0.0;
        }
#pragma unroll
        for (int l = 0; l < LW; l += 2) {
          if (stop || jj > k) break;
          double c0 = a[l] / n, c1 = a[l + 1] / n;
          if (lag0 + l == 0) acv0 = c0;
          double g = c0 + c1;                       // var.jl:57
          if (g <= 0) { stop = true; break; }      // :58-61 (m = j)
          if (monotone && jj >= 1 && g > gprev) g = gprev;   // :65-71
          gsum += g; gprev = g;
          jj++;
        }
      }
      v = (-acv0 + 2.0 * gsum) / n;                 // :74
    }
  }
  const int64_t o = j * Cp + c;
  if (mean_o) mean_o[o] = mu;
  if (viid_o) viid_o[o] = viid;
  if (var_o) var_o[o] = v;
  if (ess_o) ess_o[o] = n * viid / v;               // ess.jl:9
  if (act_o) act_o[o] = v / viid;                   // ess.jl:18
}

cudaError_t launch_stats(const double* samples, int64_t S, int64_t d, int64_t C, int64_t Cp, int vtype, int64_t maxlag,
                         int64_t batchlen, double* mean, double* var_iid, double* var, double* ess, double* actime,
                         cudaStream_t st) {
  dim3 grid((unsigned)((C + 127) / 128), (unsigned)d);
  stats_kernel<<<grid, 128, 0, st>>>(samples, S, d, C, Cp, vtype, maxlag, batchlen, mean, var_iid, var, ess, actime);
  return cudaGetLastError();
}

__global__ void accept_rate_kernel(const uint8_t* accept, int64_t S, int64_t C, int64_t Cp, double* rate) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  long long s = 0;
  for (int64_t t = 0; t < S; t++) s += accept[t * Cp + c];
  rate[c] = (double)s * 100.0 / (double)S;          // summary.jl:13
}
cudaError_t launch_accept_rate(const uint8_t* accept, int64_t S, int64_t C, int64_t Cp, double* rate, cudaStream_t st) {
  accept_rate_kernel<<<(unsigned)((C + 127) / 128), 128, 0, st>>>(accept, S, C, Cp, rate);
  return cudaGetLastError();
}

}  // namespace mg
