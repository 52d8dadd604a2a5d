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
// pair sum; here the lags are produced a window at a time and the scan stops at the truncation point:
// the same value with O(n m) work.
//
// Bound: HBM (8 S d C algorithmic bytes per pass).  Two passes over the draws in the common case:
//   pass 1  stats_mean_kernel   the mean (serial sum, the reference's order) and, for batch means, the mean of
//                               the batch means;
//   pass 2  stats_var_kernel    the centred sum of squares (== lag 0) together with the first NP_WIN Geyer pair sums
//                               (16 lags) from one register ring (Geyer's scan usually stops inside it: a noise-level
//                               pair sum is negative with probability 1/2), or the batch-means variance.
//   Only a series whose pair sums are all still positive goes on, NP_WIN pairs per further pass (stats_more_warp_kernel / stats_more_lane_kernel).
// Round 1 made four passes for IMSE (mean, variance, and a 16-lag window pass re-reading x[t] and x[t + lag]) at 120
// registers per thread (16 warps per SM) and shifted the window through registers (~30 moves per element: issue-bound).
// The rings below are rotated by unrolling (static register names, no moves), the next chunk's loads are issued before
// the current chunk's arithmetic, and the kernels are specialised per estimator (32-80 registers; the windows beyond
// the first window live in their own kernel and touch only the unfinished series).
// Compiled with -fmad=false: the mean, the variance and lag 0 are summed exactly as the reference does (mul, then add);
// the pair sums use an explicit fma on v_t (v_{t+2j} + v_{t+2j+1}) (1e-16 per term, agreement with the oracle 1e-13): a
// quarter of the FP64 instructions of two separate mul + add lags.
#include "stats.h"

namespace mg {

constexpr int ST_THREADS = 128;
constexpr int NP_WIN = 8;     // Geyer pairs (2 lags each) per window: the first window comes with the variance

// ---- pass 1: mean (mean.jl:6), and the mean of the batch means (var.jl:20-26) ----
__global__ void __launch_bounds__(ST_THREADS) stats_mean_kernel(const double* __restrict__ samples, int64_t S, int64_t d, int64_t C,
                                                                 int64_t Cp, int64_t batchlen, double* __restrict__ mean_o,
                                                                 double* __restrict__ bmean_o) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (c >= C) return;
  const double* x = samples + j * Cp + c;     // x[t] at x[t * st]
  const int64_t st = d * Cp;
  double s = 0.0;
  constexpr int U = 16;                        // independent loads in flight per thread
  int64_t t = 0;
  if (bmean_o == nullptr) {
    for (; t + U <= S; t += U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = x[(t + u) * st];
#pragma unroll
      for (int u = 0; u < U; u++) s += v[u];
    }
    for (; t < S; t++) s += x[t * st];
  } else {
    const int64_t nb = S / batchlen;
    double sm = 0.0, bs = 0.0;
    int64_t inb = 0, b = 0;
    auto take = [&](double v) {
      s += v;
      if (b < nb) {
        bs += v;
        if (++inb == batchlen) { sm += bs / (double)batchlen; bs = 0.0; inb = 0; b++; }
      }
    };
    for (; t + U <= S; t += U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = x[(t + u) * st];
#pragma unroll
      for (int u = 0; u < U; u++) take(v[u]);
    }
    for (; t < S; t++) take(x[t * st]);
    bmean_o[j * Cp + c] = sm / (double)nb;
  }
  mean_o[j * Cp + c] = s / (double)S;
}

// Geyer's scan only needs the PAIR sums Gamma_j = gamma_{2j} + gamma_{2j+1} (var.jl:57), and
//   n Gamma_j = sum_t v_t (v_{t+2j} + v_{t+2j+1}) = sum_t v_t s_{t+2j},   v = x - mu,  s_u = v_u + v_{u+1}  (zeros beyond the end),
// i.e. one multiply-add per pair and element instead of two.  One window produces NP pairs starting at the even lag lag0:
// a ring sr[0 .. 2 NP) holds s at positions t + lag0 .. t + lag0 + 2 NP - 1 and is rotated by unrolling (static register
// names, no moves); it is refilled from ONE stream of x (position p + 1 completes s_p), the multiplier v_t is a second
// stream that trails the first by lag0 + 2 NP elements (L1 / L2 hits for the first window).  FIRST also accumulates
// a0 = sum v^2 as mul + add: the reference's variance and gamma_0, bit for bit.
template <int NP, bool FIRST>
__device__ __forceinline__ void acov_pairs(const double* __restrict__ x, int64_t st, int64_t S, double mu, int64_t lag0, int64_t t_begin,
                                           int64_t t_end, double& a0, double (&G)[NP]) {
  // sums over t in [t_begin, t_end) (the whole series: [0, S - lag0))
  constexpr int R = 2 * NP;
#pragma unroll
  for (int j = 0; j < NP; j++) G[j] = 0.0;
  a0 = 0.0;
  double sr[R], nx[R];
  auto cent = [&](int64_t idx) -> double { return (idx < S) ? x[idx * st] - mu : 0.0; };   // beyond the end: exact zeros
  double vcur = cent(t_begin + lag0);          // v at the position whose s is formed next
#pragma unroll
  for (int u = 0; u < R; u++) { const double vn = cent(t_begin + lag0 + u + 1); sr[u] = vcur + vn; vcur = vn; }
  // software pipeline: the ring refill values of chunk k + 1 (the DRAM stream) are requested at the top of chunk k and
  // consumed one iteration later, so a full chunk of arithmetic separates each of these loads from its use (left to itself
  // the compiler sinks every load next to its use: 2 loads in flight per thread, 1.7 TB/s).  The multipliers v_t were read
  // 2 NP + lag0 elements earlier as ring values: L1 / L2 hits, loaded in place.
#pragma unroll
  for (int u = 0; u < R; u++) nx[u] = cent(t_begin + lag0 + R + u + 1);
  for (int64_t t0 = t_begin; t0 < t_end; t0 += R) {
    double cn[R];
#pragma unroll
    for (int u = 0; u < R; u++) cn[u] = nx[u];
#pragma unroll
    for (int u = 0; u < R; u++) nx[u] = cent(t0 + R + lag0 + R + u + 1);
#pragma unroll
    for (int u = 0; u < R; u++) {
      const double v = (t0 + u < t_end) ? x[(t0 + u) * st] - mu : 0.0;
      if (FIRST) a0 += v * v;                                       // == sum (x - mu)^2 of Base.var, same roundings
#pragma unroll
      for (int j = 0; j < NP; j++) G[j] = fma(v, sr[(u + 2 * j) % R], G[j]);
      sr[u] = vcur + cn[u]; vcur = cn[u];
    }
  }
}

// The first window (lag0 = 0) needs no look-ahead at all: n Gamma_j = sum_u s_u v_{u-2j}, so ONE stream of x suffices --
// when v_u arrives, s_{u-1} = v_{u-1} + v_u is complete and meets the ring of the last 2 NP values (every other one).
// a0 = sum v^2 is accumulated as mul + add in the order of the draws: the reference's variance and gamma_0, bit for bit.
// The raw values of the next chunk are requested one iteration ahead (software pipeline, see acov_pairs).
template <int NP>
__device__ __forceinline__ void acov_first(const double* __restrict__ x, int64_t st, int64_t S, double mu, double& a0, double (&G)[NP]) {
  constexpr int R = 2 * NP;
#pragma unroll
  for (int j = 0; j < NP; j++) G[j] = 0.0;
  a0 = 0.0;
  double ring[R], nx[R];
#pragma unroll
  for (int i = 0; i < R; i++) { ring[i] = 0.0; nx[i] = (i < S) ? x[(int64_t)i * st] : mu; }   // beyond the end: v = 0 exactly
  double vprev = 0.0;
  for (int64_t u0 = 0; u0 <= S; u0 += R) {       // u = S closes s_{S-1} = v_{S-1} + 0
    double cur[R];
#pragma unroll
    for (int i = 0; i < R; i++) cur[i] = nx[i];
#pragma unroll
    for (int i = 0; i < R; i++) { const int64_t idx = u0 + R + i; nx[i] = (idx < S) ? x[idx * st] : mu; }
#pragma unroll
    for (int i = 0; i < R; i++) {
      const double vu = cur[i] - mu;
      const double sp = vprev + vu;                                 // s_{u-1}
      a0 += vu * vu;                                                // == sum (x - mu)^2 of Base.var, same roundings
#pragma unroll
      for (int j = 0; j < NP; j++) G[j] = fma(sp, (j == 0) ? vprev : ring[(i - 1 - 2 * j + 2 * R) % R], G[j]);
      ring[i] = vu;
      vprev = vu;
    }
  }
}

// Geyer scan over the pairs of one window (var.jl:56-71); returns true when the sequence is truncated
template <int NP>
__device__ __forceinline__ bool geyer_window(const double (&G)[NP], double n, int64_t k, bool monotone, int64_t& jj, double& gsum,
                                             double& gprev) {
#pragma unroll
  for (int l = 0; l < NP; l++) {
    if (jj > k) return true;
    double g = G[l] / n;                         // var.jl:57 (c0 + c1, formed here as one sum)
    if (g <= 0) return true;                     // :58-61 (m = j)
    if (monotone && jj >= 1 && g > gprev) g = gprev;   // :65-71
    gsum += g; gprev = g;
    jj++;
  }
  return jj > k;
}

// ---- pass 2 (and the rare further passes): variance / batch means / Geyer ----
template <int VT>
__global__ void __launch_bounds__(ST_THREADS, 3) stats_var_kernel(const double* __restrict__ samples, int64_t S, int64_t d, int64_t C,
                                                                int64_t Cp, int64_t maxlag, int64_t batchlen,
                                                                const double* __restrict__ mean_i, const double* __restrict__ bmean_i,
                                                                double* viid_o, double* var_o, double* ess_o, double* act_o,
                                                                double* __restrict__ more, int32_t* __restrict__ more_list,
                                                                unsigned int* __restrict__ more_count) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (c >= C) return;
  const double* x = samples + j * Cp + c;
  const int64_t st = d * Cp;
  const double n = (double)S;
  const double mu = mean_i[j * Cp + c];
  const int64_t o = j * Cp + c;
  double ss = 0.0, v = CUDART_NAN;
  if (VT == MCMCGPU_VAR_IID || VT == MCMCGPU_VAR_BM) {
    constexpr int U = 16;
    const int64_t nb = (VT == MCMCGPU_VAR_BM) ? S / batchlen : 0;
    const double mb = (VT == MCMCGPU_VAR_BM) ? bmean_i[j * Cp + c] : 0.0;
    double sv = 0.0, bs = 0.0;
    int64_t inb = 0, b = 0;
    auto take = [&](double xv) {
      const double dv = xv - mu;
      ss += dv * dv;
      if (VT == MCMCGPU_VAR_BM && b < nb) {
        bs += xv;
        if (++inb == batchlen) { const double e = bs / (double)batchlen - mb; sv += e * e; bs = 0.0; inb = 0; b++; }
      }
    };
    int64_t t = 0;
    for (; t + U <= S; t += U) {
      double w[U];
#pragma unroll
      for (int u = 0; u < U; u++) w[u] = x[(t + u) * st];
#pragma unroll
      for (int u = 0; u < U; u++) take(w[u]);
    }
    for (; t < S; t++) take(x[t * st]);
    if (VT == MCMCGPU_VAR_BM && nb > 1) v = (double)batchlen * (sv / (double)(nb - 1)) / (double)(nb * batchlen);   // var.jl:20-26
  } else {
    // Geyer IMSE / IPSE (var.jl:45-75 / :95-116)
    const int64_t k = (maxlag - 1 >= 0) ? (maxlag - 1) / 2 : -1;   // floor((maxlag-1)/2)
    const bool monotone = (VT == MCMCGPU_VAR_IMSE);
    double acv0 = 0.0, gsum = 0.0, gprev = 0.0;
    int64_t jj = 0;
    bool stop;
    {
      double G[NP_WIN];
      acov_first<NP_WIN>(x, st, S, mu, ss, G);
      acv0 = ss / n;                                                // gamma_0 (var.jl:53)
      stop = (k < 0) || geyer_window<NP_WIN>(G, n, k, monotone, jj, gsum, gprev);
    }
    // a series whose pair sums are still positive goes on in the stats_more_* kernels (their windows need two load streams and
    // twice the registers: kept out of this kernel so that the common case runs at full occupancy)
    const int64_t plane = d * Cp;
    more[o] = stop ? 0.0 : 1.0;
    if (!stop) {
      more_list[atomicAdd(more_count, 1u)] = (int32_t)o;            // compact list of the unfinished series
      more[plane + o] = acv0; more[2 * plane + o] = gsum; more[3 * plane + o] = gprev; more[4 * plane + o] = (double)jj;
      more[5 * plane + o] = (ss / (double)(S - 1)) / n;
      return;
    }
    if (k >= 0) v = (-acv0 + 2.0 * gsum) / n;                      // :74
  }
  const double viid = (ss / (double)(S - 1)) / n;                  // var.jl:7-8
  if (VT == MCMCGPU_VAR_IID) v = viid;
  if (viid_o) viid_o[o] = viid;
  if (var_o) var_o[o] = v;
  if (ess_o) ess_o[o] = n * viid / v;               // ess.jl:9
  if (act_o) act_o[o] = v / viid;                   // ess.jl:18
}

// The windows beyond the first NP_WIN pairs, for the series stats_var_kernel left unfinished.
// Few unfinished series (HMC(0.75) on the 3-D Normal: 4.4 % survive the first 8 pairs): they are first copied out of the
// chain-minor draws into a compact TIME-CONTIGUOUS buffer (a series is strided by d * Cp doubles in the draws: every element
// costs a whole DRAM sector, paid here once), then ONE WARP PER SERIES walks its copy: the lanes split the time axis, each
// streaming its own contiguous segment, and the partial pair sums are combined by a butterfly (every lane holds the same
// sum and takes the same decision).  A thread per series would leave the other 31 lanes of almost every warp idle for whole
// passes over the strided draws (measured: 7.9 ms and 8 GB of DRAM reads for those windows at 65 536 x 3 series x 9000 draws).
__global__ void stats_gather_kernel(const double* __restrict__ samples, int64_t S, int64_t st, const int32_t* __restrict__ more_list,
                                    unsigned int cnt, double* __restrict__ buf) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t e = warp; e < (int64_t)cnt; e += nwarps) {
    const double* x = samples + more_list[e];
    double* y = buf + e * S;
    int64_t t = lane;
    for (; t + 96 < S; t += 128) {               // 4 independent strided loads per lane in flight
      const double v0 = x[t * st], v1 = x[(t + 32) * st], v2 = x[(t + 64) * st], v3 = x[(t + 96) * st];
      y[t] = v0; y[t + 32] = v1; y[t + 64] = v2; y[t + 96] = v3;
    }
    for (; t < S; t += 32) y[t] = x[t * st];
  }
}

__global__ void __launch_bounds__(ST_THREADS) stats_more_warp_kernel(const double* __restrict__ buf, int64_t S, int64_t d, int64_t Cp,
                                                                      int64_t maxlag, int monotone, const double* __restrict__ mean_i,
                                                                      const double* __restrict__ more, const int32_t* __restrict__ more_list,
                                                                      const unsigned int* __restrict__ more_count, double* viid_o,
                                                                      double* var_o, double* ess_o, double* act_o) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t plane = d * Cp, st = 1, k = (maxlag - 1) / 2;
  const double n = (double)S;
  const unsigned int cnt = *more_count;
  for (int64_t e = warp; e < (int64_t)cnt; e += nwarps) {
    const int64_t o = more_list[e];                // o = j * Cp + c
    const double* x = buf + e * S;                 // the series' time-contiguous copy
    const double mu = mean_i[o];
    double acv0 = more[plane + o], gsum = more[2 * plane + o], gprev = more[3 * plane + o];
    int64_t jj = (int64_t)more[4 * plane + o];
    const double viid = more[5 * plane + o];
    bool stop = false;
    for (int64_t lag0 = 2 * NP_WIN; !stop; lag0 += 2 * NP_WIN) {
      const int64_t T = S - lag0, seg = (T > 0) ? (T + 31) / 32 : 0;
      const int64_t tb = lane * seg, te = (tb + seg < T) ? tb + seg : T;
      double G[NP_WIN], unused;
      if (tb < te) acov_pairs<NP_WIN, false>(x, st, S, mu, lag0, tb, te, unused, G);
      else {
#pragma unroll
        for (int j = 0; j < NP_WIN; j++) G[j] = 0.0;
      }
#pragma unroll
      for (int j = 0; j < NP_WIN; j++) {
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) G[j] += __shfl_xor_sync(0xffffffffu, G[j], w);
      }
      stop = geyer_window<NP_WIN>(G, n, k, monotone != 0, jj, gsum, gprev);
    }
    if (lane == 0) {
      const double v = (-acv0 + 2.0 * gsum) / n;        // var.jl:74
      if (viid_o) viid_o[o] = viid;
      if (var_o) var_o[o] = v;
      if (ess_o) ess_o[o] = n * viid / v;               // ess.jl:9
      if (act_o) act_o[o] = v / viid;                   // ess.jl:18
    }
  }
}

// Many unfinished series (slowly mixing chains): one thread per series, coalesced across the chains of a warp
__global__ void __launch_bounds__(ST_THREADS) stats_more_lane_kernel(const double* __restrict__ samples, int64_t S, int64_t d, int64_t C, int64_t Cp,
                                                                 int64_t maxlag, int monotone, const double* __restrict__ mean_i,
                                                                 const double* __restrict__ more, double* viid_o, double* var_o,
                                                                 double* ess_o, double* act_o) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (c >= C) return;
  const int64_t o = j * Cp + c, plane = d * Cp;
  if (more[o] == 0.0) return;
  const double* x = samples + j * Cp + c;
  const int64_t st = d * Cp;
  const double n = (double)S, mu = mean_i[o];
  const int64_t k = (maxlag - 1) / 2;
  double acv0 = more[plane + o], gsum = more[2 * plane + o], gprev = more[3 * plane + o];
  int64_t jj = (int64_t)more[4 * plane + o];
  const double viid = more[5 * plane + o];
  bool stop = false;
  for (int64_t lag0 = 2 * NP_WIN; !stop; lag0 += 2 * NP_WIN) {
    double G[NP_WIN], unused;
    acov_pairs<NP_WIN, false>(x, st, S, mu, lag0, 0, S - lag0, unused, G);
    stop = geyer_window<NP_WIN>(G, n, k, monotone != 0, jj, gsum, gprev);
  }
  const double v = (-acv0 + 2.0 * gsum) / n;        // var.jl:74
  if (viid_o) viid_o[o] = viid;
  if (var_o) var_o[o] = v;
  if (ess_o) ess_o[o] = n * viid / v;               // ess.jl:9
  if (act_o) act_o[o] = v / viid;                   // ess.jl:18
}

cudaError_t launch_stats(const double* samples, int64_t S, int64_t d, int64_t C, int64_t Cp, int vtype, int64_t maxlag,
                         int64_t batchlen, double* mean, double* var_iid, double* var, double* ess, double* actime,
                         double* scratch, unsigned int* n_unfinished, cudaStream_t st) {
  // `mean` must be a device buffer [d][Cp] (pass 2 reads it); scratch: [d][Cp] doubles for batch means,
  // STATS_SCRATCH_PLANES * d * Cp + 2 for IMSE / IPSE (6 planes of scan state, the list of unfinished series, its length)
  double* bmean_scratch = scratch;
  const int64_t plane = d * Cp;
  int32_t* more_list = scratch ? reinterpret_cast<int32_t*>(scratch + 6 * plane) : nullptr;
  unsigned int* more_count = scratch ? reinterpret_cast<unsigned int*>(scratch + 7 * plane) : nullptr;
  const bool geyer = (vtype == MCMCGPU_VAR_IMSE || vtype == MCMCGPU_VAR_IPSE);
  if (n_unfinished) *n_unfinished = 0;
  if (geyer) {
    cudaError_t e0 = cudaMemsetAsync(more_count, 0, sizeof(unsigned int), st);
    if (e0 != cudaSuccess) return e0;
  }
  dim3 grid((unsigned)((C + ST_THREADS - 1) / ST_THREADS), (unsigned)d);
  const bool bm = (vtype == MCMCGPU_VAR_BM);
  stats_mean_kernel<<<grid, ST_THREADS, 0, st>>>(samples, S, d, C, Cp, batchlen, mean, bm ? bmean_scratch : nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (!var_iid && !var && !ess && !actime) return cudaSuccess;
  switch (vtype) {
    case MCMCGPU_VAR_IID: stats_var_kernel<MCMCGPU_VAR_IID><<<grid, ST_THREADS, 0, st>>>(samples, S, d, C, Cp, maxlag, batchlen, mean, nullptr, var_iid, var, ess, actime, nullptr, nullptr, nullptr); break;
    case MCMCGPU_VAR_BM: stats_var_kernel<MCMCGPU_VAR_BM><<<grid, ST_THREADS, 0, st>>>(samples, S, d, C, Cp, maxlag, batchlen, mean, bmean_scratch, var_iid, var, ess, actime, nullptr, nullptr, nullptr); break;
    case MCMCGPU_VAR_IMSE: stats_var_kernel<MCMCGPU_VAR_IMSE><<<grid, ST_THREADS, 0, st>>>(samples, S, d, C, Cp, maxlag, batchlen, mean, nullptr, var_iid, var, ess, actime, scratch, more_list, more_count); break;
    case MCMCGPU_VAR_IPSE: stats_var_kernel<MCMCGPU_VAR_IPSE><<<grid, ST_THREADS, 0, st>>>(samples, S, d, C, Cp, maxlag, batchlen, mean, nullptr, var_iid, var, ess, actime, scratch, more_list, more_count); break;
    default: return cudaErrorInvalidValue;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (geyer && n_unfinished) {      // how many series go on: the caller sizes the compact copy from it
    e = cudaMemcpyAsync(n_unfinished, more_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  return e;
}

cudaError_t launch_stats_more(const double* samples, int64_t S, int64_t d, int64_t C, int64_t Cp, int vtype, int64_t maxlag,
                              const double* mean, double* var_iid, double* var, double* ess, double* actime, double* scratch,
                              unsigned int cnt, double* gather, cudaStream_t st) {
  if (cnt == 0) return cudaSuccess;
  const int64_t plane = d * Cp;
  const int32_t* more_list = reinterpret_cast<const int32_t*>(scratch + 6 * plane);
  const unsigned int* more_count = reinterpret_cast<const unsigned int*>(scratch + 7 * plane);
  const int mono = vtype == MCMCGPU_VAR_IMSE ? 1 : 0;
  if (!gather) {
    dim3 grid((unsigned)((C + ST_THREADS - 1) / ST_THREADS), (unsigned)d);
    stats_more_lane_kernel<<<grid, ST_THREADS, 0, st>>>(samples, S, d, C, Cp, maxlag, mono, mean, scratch, var_iid, var, ess, actime);
    return cudaGetLastError();
  }
  const unsigned blocks = (unsigned)((cnt + 3) / 4 < 148 * 16 ? (cnt + 3) / 4 : 148 * 16);     // 4 warps per block
  stats_gather_kernel<<<blocks, ST_THREADS, 0, st>>>(samples, S, d * Cp, more_list, cnt, gather);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  stats_more_warp_kernel<<<blocks, ST_THREADS, 0, st>>>(gather, S, d, Cp, maxlag, mono, mean, scratch, more_list, more_count, var_iid, var, ess, actime);
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
