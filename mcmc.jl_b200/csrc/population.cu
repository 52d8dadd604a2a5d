// population.cu -- population runners on the device, closed-form families.
//   SeqMC         src/runners/SeqMC.jl:39-122         one thread mutates one particle per target; weights, the
//                                                      variance trigger, the cumulative sum and the multinomial
//                                                      resampling (prefix sum + binary search) stay on the device
//   SerialTempMC  src/runners/SerialTempMC.jl:31-85   one thread = one tempering replica, whole run in one launch
// Both are built on reset(task, pars) followed by one consume(task): the sampler re-evaluates the log-target at
// pars (RWM.jl:49, MALA.jl:75-78, HMC.jl:114-116) and runs one loop body.  Compiled with -fmad=false.
#include "population.h"
#include "families.cuh"

namespace mg {

template <int FAM, int D>
__device__ __forceinline__ void reset_and_step(const PopTasks& T, int t, const double (&pars)[D], const double (&z)[D],
                                               double u, double (&ppars)[D], double& plt, double& lt0,
                                               unsigned long long& nev) {
  ModelDev M; M.family = FAM; M.N = 0; M.d = T.d; M.hyper[0] = T.hyper[t][0]; M.hyper[1] = T.hyper[t][1]; M.hyper[2] = 0; M.hyper[3] = 0;
  M.series = nullptr;
  const int d = T.d;
  const int kind = T.kind[t];
  double prop[D], grad[D], pg[D];
  bool acc;
  double lt, l2;
  if (kind == MCMCGPU_RWM) {
    lt = Family<FAM, D>::evalallg(M, nullptr, d, pars, grad);
#pragma unroll
    for (int j = 0; j < D; j++) prop[j] = pars[j] + z[j] * (1.0 * T.scale[t]);   // model.scale = ones (RWM.jl:52,59)
    l2 = Family<FAM, D>::evalallg(M, nullptr, d, prop, pg);
    double ratio = l2 - lt;
    acc = ratio > 0 || ratio > log(u);                                           // RWM.jl:63
    nev += 2;
  } else if (kind == MCMCGPU_MALA) {
    const double h = T.scale[t], sq = sqrt(h), lc = log(MG_TWO_PI * h) / 2.0;
    double mean[D], qno = 0.0, qon = 0.0;
    lt = Family<FAM, D>::evalallg(M, nullptr, d, pars, grad);
#pragma unroll
    for (int j = 0; j < D; j++) { mean[j] = pars[j] + (h / 2.0) * grad[j]; prop[j] = mean[j] + sq * z[j]; }   // MALA.jl:98-100
    l2 = Family<FAM, D>::evalallg(M, nullptr, d, prop, pg);
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) { double tt = mean[j] - prop[j]; qno += -(tt * tt) / (2.0 * h) - lc; }
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) { double m2 = prop[j] + (h / 2.0) * pg[j]; double tt = m2 - pars[j]; qon += -(tt * tt) / (2.0 * h) - lc; }
    double ratio = l2 + qon - lt - qno;
    acc = ratio > 0 || ratio > log(u);                                           // MALA.jl:107-108
    nev += 2;
  } else {  // HMC with fixed nLeaps (HMC.jl:136-158)
    const double eps = T.scale[t];
    double mom[D];
    lt = Family<FAM, D>::evalallg(M, nullptr, d, pars, grad);
    double mm = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) { if (j < d) mm += z[j] * z[j]; mom[j] = z[j]; prop[j] = pars[j]; pg[j] = grad[j]; }
    const double H0 = -lt + 0.5 * mm;
    l2 = lt;
    for (int l = 0; l < T.nleaps[t]; l++) {
#pragma unroll
      for (int j = 0; j < D; j++) mom[j] += (0.5 * pg[j]) * eps;
#pragma unroll
      for (int j = 0; j < D; j++) prop[j] += eps * mom[j];
      l2 = Family<FAM, D>::evalallg(M, nullptr, d, prop, pg);
#pragma unroll
      for (int j = 0; j < D; j++) mom[j] += (0.5 * pg[j]) * eps;
    }
    double m2 = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) m2 += mom[j] * mom[j];
    const double H = -l2 + 0.5 * m2;
    acc = u < exp(H0 - H);
    nev += 1 + (unsigned long long)T.nleaps[t];
  }
#pragma unroll
  for (int j = 0; j < D; j++) ppars[j] = acc ? prop[j] : pars[j];
  plt = acc ? l2 : lt;
  lt0 = lt;
}

// ---- SeqMC ------------------------------------------------------------------------------------------
template <int FAM, int D>
__global__ void __launch_bounds__(128) seqmc_mutate_kernel(const SeqArgs A) {
  // SeqMC.jl:66-72: force the task to the particle, take one step, update the weight with the old-vs-new target ratio
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.npart) return;
  const int d = A.T.d, t = (int)A.target;
  const int64_t gn = (int64_t)A.rank * A.npart + n;     // global particle id (Philox key / injected-draw index)
  const int64_t k = ((A.iter - 1) * A.T.nt + t) * A.gpart + gn;
  const uint32_t pstep = (uint32_t)((A.iter - 1) * A.T.nt + t + 1);
  double pars[D], z[D], pp[D];
#pragma unroll
  for (int j = 0; j < D; j++) pars[j] = (j < d) ? A.pars[j * A.Np + n] : 0.0;
  double u;
  if (A.inj_normals) {
#pragma unroll
    for (int j = 0; j < D; j++) z[j] = (j < d) ? A.inj_normals[k * d + j] : 0.0;
    u = A.inj_uniforms[k];
  } else {
#pragma unroll
    for (int b = 0; 2 * b < D; b++) {
      double z0 = 0.0, z1 = 0.0;
      if (2 * b < d) philox_normal_pair(A.seed, (uint64_t)gn, pstep, (uint32_t)b, z0, z1);
      z[2 * b] = z0;
      if (2 * b + 1 < D) z[2 * b + 1] = (2 * b + 1 < d) ? z1 : 0.0;
    }
    u = philox_uniform(A.seed, (uint64_t)gn, pstep);
  }
  double plt, ll0;
  unsigned long long nev = 0;
  reset_and_step<FAM, D>(A.T, t, pars, z, u, pp, plt, ll0, nev);
#pragma unroll
  for (int j = 0; j < D; j++) if (j < d) A.pars[j * A.Np + n] = pp[j];
  A.logW[n] += ll0 - A.logtarget[n];                                             // :70
  A.logtarget[n] = plt;                                                          // :71
  atomicAdd(A.nevals, nev);
}

// accessors over the population: local arrays on one GPU, the all-gathered [rank][d+2][Np] image otherwise
__device__ __forceinline__ double g_logW(const SeqArgs& A, int64_t gn) {
  if (A.nranks == 1) return A.logW[gn];
  const int64_t r = gn / A.npart, l = gn % A.npart;
  return A.gathered[(r * (A.T.d + 2) + A.T.d + 1) * A.Np + l];
}
__device__ __forceinline__ double g_lt(const SeqArgs& A, int64_t gn) {
  if (A.nranks == 1) return A.logtarget[gn];
  const int64_t r = gn / A.npart, l = gn % A.npart;
  return A.gathered[(r * (A.T.d + 2) + A.T.d) * A.Np + l];
}
__device__ __forceinline__ double g_par(const SeqArgs& A, int j, int64_t gn) {
  if (A.nranks == 1) return A.pars[j * A.Np + gn];
  const int64_t r = gn / A.npart, l = gn % A.npart;
  return A.gathered[(r * (A.T.d + 2) + j) * A.Np + l];
}

__global__ void seqmc_pack_kernel(const SeqArgs A) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.npart) return;
  const int d = A.T.d;
  for (int j = 0; j < d; j++) A.sendbuf[j * A.Np + n] = A.pars[j * A.Np + n];
  A.sendbuf[d * A.Np + n] = A.logtarget[n];
  A.sendbuf[(d + 1) * A.Np + n] = A.logW[n];
}

__global__ void __launch_bounds__(1024) seqmc_resample_kernel(const SeqArgs A) {
  // SeqMC.jl:74-89.  One block (every rank runs it on the same gathered population).  The sums that decide and define
  // the resampling are taken sequentially by one thread, in the reference's order (var two-pass, cumsum), so that the
  // resampled indices are reproducible bit for bit; the binary searches and the gather are done by all threads.
  __shared__ int do_res;
  const int64_t gp = A.gpart;
  for (int64_t n = threadIdx.x; n < gp; n += blockDim.x) A.W[n] = exp(g_logW(A, n));         // :76
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int64_t n = 0; n < gp; n++) s += A.W[n];
    const double mean = s / (double)gp;
    double ss = 0.0;
    for (int64_t n = 0; n < gp; n++) { double v = A.W[n] - mean; ss += v * v; }
    const double var = ss / (double)(gp - 1);
    do_res = (var < A.trigger) ? 1 : 0;                                                      // :77
    if (do_res) {
      double c = 0.0;
      for (int64_t n = 0; n < gp; n++) { c += A.W[n]; A.cp[n] = c / s; }                     // :78 cumsum(W) / sum(W)
      atomicAdd(A.nres, 1ull);
    }
  }
  __syncthreads();
  if (!do_res) return;
  const int d = A.T.d;
  const int64_t kbase = ((A.iter - 1) * A.T.nt + A.target) * gp;
  const uint32_t pstep = (uint32_t)((A.iter - 1) * A.T.nt + A.target + 1);
  for (int64_t n = threadIdx.x; n < A.npart; n += blockDim.x) {                              // this rank's slots
    const int64_t gn = (int64_t)A.rank * A.npart + n;
    double l;
    if (A.inj_res) l = A.inj_res[kbase + gn];
    else {
      u4 o = philox4x32_10((uint32_t)gn, (uint32_t)((uint64_t)gn >> 32), pstep, 0xFFFFFFFEu, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
      l = u01(o.x, o.y);
    }
    int64_t lo = 0, hi = gp - 1;                                                             // :82 findfirst(p -> p >= l, cp)
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (A.cp[mid] >= l) hi = mid; else lo = mid + 1; }
    for (int j = 0; j < d; j++) A.pars_tmp[j * A.Np + n] = g_par(A, j, lo);                  // :84 pars = pars[rs]
    A.lt_tmp[n] = g_lt(A, lo);                                                               // :86
  }
  __syncthreads();
  for (int64_t n = threadIdx.x; n < A.npart; n += blockDim.x) {
    for (int j = 0; j < d; j++) A.pars[j * A.Np + n] = A.pars_tmp[j * A.Np + n];
    A.logtarget[n] = A.lt_tmp[n];
    A.logW[n] = 0.0;                                                                         // :85
  }
}

// mutation made by the wave engine (any family, any d): take over its one-step result and update the weights (:70-71)
__global__ void seqmc_apply_kernel(const SeqArgs A, const double* __restrict__ ppars, const double* __restrict__ plt,
                                   const double* __restrict__ ll0) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.npart) return;
  for (int j = 0; j < A.T.d; j++) A.pars[j * A.Np + n] = ppars[j * A.Np + n];
  A.logW[n] += ll0[n] - A.logtarget[n];                                          // :70
  A.logtarget[n] = plt[n];                                                       // :71
}
cudaError_t launch_seqmc_apply(const SeqArgs& A, const double* ppars, const double* plt, const double* ll0, cudaStream_t st) {
  seqmc_apply_kernel<<<(unsigned)((A.npart + 127) / 128), 128, 0, st>>>(A, ppars, plt, ll0);
  return cudaGetLastError();
}

__global__ void seqmc_store_kernel(const SeqArgs A) {
  // SeqMC.jl:92-100: logtarget = zeros; after burn-in store every particle and its weight
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.npart) return;
  A.logtarget[n] = 0.0;
  if (A.iter > A.burnin) {
    const int64_t pos = (A.iter - A.burnin - 1) * A.npart + n;
    for (int j = 0; j < A.T.d; j++) A.samples[pos * A.T.d + j] = A.pars[j * A.Np + n];
    A.weights[pos] = exp(A.logW[n]);
  }
}

template <int FAM>
static cudaError_t mutate_d(const SeqArgs& A, cudaStream_t st) {
  unsigned blocks = (unsigned)((A.npart + 127) / 128);
  int d = A.T.d;
  if (d <= 1) seqmc_mutate_kernel<FAM, 1><<<blocks, 128, 0, st>>>(A);
  else if (d <= 2) seqmc_mutate_kernel<FAM, 2><<<blocks, 128, 0, st>>>(A);
  else if (d <= 4) seqmc_mutate_kernel<FAM, 4><<<blocks, 128, 0, st>>>(A);
  else seqmc_mutate_kernel<FAM, 8><<<blocks, 128, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_seqmc_mutate(const SeqArgs& A, cudaStream_t st) {
  switch (A.T.family) {
    case MCMCGPU_FAM_NORMAL_FN: return mutate_d<MCMCGPU_FAM_NORMAL_FN>(A, st);
    case MCMCGPU_FAM_NORMAL_DSL: return mutate_d<MCMCGPU_FAM_NORMAL_DSL>(A, st);
    case MCMCGPU_FAM_ABS_NORMAL: return mutate_d<MCMCGPU_FAM_ABS_NORMAL>(A, st);
  }
  return cudaErrorInvalidValue;
}
cudaError_t launch_seqmc_resample(const SeqArgs& A, cudaStream_t st) {
  seqmc_resample_kernel<<<1, 1024, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_seqmc_pack(const SeqArgs& A, cudaStream_t st) {
  seqmc_pack_kernel<<<(unsigned)((A.npart + 127) / 128), 128, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_seqmc_store(const SeqArgs& A, cudaStream_t st) {
  seqmc_store_kernel<<<(unsigned)((A.npart + 127) / 128), 128, 0, st>>>(A);
  return cudaGetLastError();
}

// ---- SerialTempMC ------------------------------------------------------------------------------------
template <int FAM, int D>
__global__ void __launch_bounds__(128) serialtemp_kernel(const TempArgs A) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.nrep) return;
  const PopTasks& T = A.T;
  const int d = T.d, nt = T.nt;
  const int64_t S = A.steps - A.burnin;
  unsigned long long nev = 0;
  auto normals = [&](int64_t col, double (&z)[D]) {
    if (A.inj_normals) {
#pragma unroll
      for (int j = 0; j < D; j++) z[j] = (j < d) ? A.inj_normals[(c * (A.steps + 2) + col) * d + j] : 0.0;
    } else {
#pragma unroll
      for (int b = 0; 2 * b < D; b++) {
        double z0 = 0.0, z1 = 0.0;
        if (2 * b < d) philox_normal_pair(A.seed, (uint64_t)c, (uint32_t)(col + 1), (uint32_t)b, z0, z1);   // sampler step = column + 1
        z[2 * b] = z0;
        if (2 * b + 1 < D) z[2 * b + 1] = (2 * b + 1 < d) ? z1 : 0.0;
      }
    }
  };
  // sampler uniforms sit at Philox step column + 1 (the step index of a one-step run, as mcmcgpu_run_serialtemp_models
  // draws them); the pick / swap uniforms of iteration i at step i in their own blocks
  auto uni = [&](const double* inj, int64_t stride, int64_t col, uint32_t block, int shift = 0) -> double {
    if (inj) return inj[c * stride + col];
    u4 o = philox4x32_10((uint32_t)c, (uint32_t)((uint64_t)c >> 32), (uint32_t)(col + shift), block, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
    return u01(o.x, o.y);
  };
  // every task is started from its model.init (:44); a non-finite start is the samplers' assertion
  double state[D], ppars[D], pars[D], pp2[D], z[D], g[D];
  int bad = 0;
  for (int t = 0; t < nt; t++) {
    ModelDev M; M.family = FAM; M.N = 0; M.d = d; M.hyper[0] = T.hyper[t][0]; M.hyper[1] = T.hyper[t][1]; M.hyper[2] = 0; M.hyper[3] = 0; M.series = nullptr;
#pragma unroll
    for (int j = 0; j < D; j++) pars[j] = (j < d) ? A.inits[t * d + j] : 0.0;
    if (!isfinite(Family<FAM, D>::evalallg(M, nullptr, d, pars, g))) bad = 1;
  }
  A.status[c] = bad;
  if (bad) return;
#pragma unroll
  for (int j = 0; j < D; j++) pars[j] = (j < d) ? A.inits[j] : 0.0;
  int at = 0;
  double plt, lt0, logtarget;
  normals(0, z);
  reset_and_step<FAM, D>(T, 0, pars, z, uni(A.inj_uniforms, A.steps + 2, 0, 0xFFFFFFFFu, 1), state, plt, lt0, nev);   // :44
#pragma unroll
  for (int j = 0; j < D; j++) pars[j] = state[j];
  normals(1, z);
  reset_and_step<FAM, D>(T, 0, pars, z, uni(A.inj_uniforms, A.steps + 2, 1, 0xFFFFFFFFu, 1), ppars, plt, lt0, nev);   // :51
#pragma unroll
  for (int j = 0; j < D; j++) state[j] = ppars[j];
  logtarget = lt0;
  for (int64_t i = 1; i <= A.steps; i++) {
    normals(i + 1, z);
    const double u = uni(A.inj_uniforms, A.steps + 2, i + 1, 0xFFFFFFFFu, 1);
    if (i % A.swap_period == 0) {                                                 // :57
      int at2 = (int)floor(uni(A.inj_pick, A.steps + 1, i, 0xFFFFFFFDu) * (double)(nt - 1));   // :59
      if (at2 > nt - 2) at2 = nt - 2;
      if (at2 >= at) at2 += 1;                                                    // :60
      double plt2, lt02;
      reset_and_step<FAM, D>(T, at2, pars, z, u, pp2, plt2, lt02, nev);           // :62-63
      if (uni(A.inj_swap, A.steps + 1, i, 0xFFFFFFFCu) < exp(logtarget - lt02 + 0.0 - 0.0)) {   // :64
        at = at2;
#pragma unroll
        for (int j = 0; j < D; j++) { ppars[j] = pp2[j]; state[j] = pp2[j]; }
        logtarget = lt02;
      }
    } else {                                                                      // :68
#pragma unroll
      for (int j = 0; j < D; j++) pars[j] = state[j];
      reset_and_step<FAM, D>(T, at, pars, z, u, ppars, plt, lt0, nev);
#pragma unroll
      for (int j = 0; j < D; j++) state[j] = ppars[j];
      logtarget = lt0;
    }
    if (i > A.burnin) {                                                           // :73-76
      const int64_t k = i - A.burnin - 1;
      for (int j = 0; j < d; j++) A.samples[(c * S + k) * d + j] = ppars[j];
      if (A.at) A.at[c * S + k] = at;
    }
  }
  atomicAdd(A.nevals, nev);
}

template <int FAM>
static cudaError_t temp_d(const TempArgs& A, cudaStream_t st) {
  unsigned blocks = (unsigned)((A.nrep + 127) / 128);
  int d = A.T.d;
  if (d <= 1) serialtemp_kernel<FAM, 1><<<blocks, 128, 0, st>>>(A);
  else if (d <= 2) serialtemp_kernel<FAM, 2><<<blocks, 128, 0, st>>>(A);
  else if (d <= 4) serialtemp_kernel<FAM, 4><<<blocks, 128, 0, st>>>(A);
  else serialtemp_kernel<FAM, 8><<<blocks, 128, 0, st>>>(A);
  return cudaGetLastError();
}
cudaError_t launch_serialtemp(const TempArgs& A, cudaStream_t st) {
  switch (A.T.family) {
    case MCMCGPU_FAM_NORMAL_FN: return temp_d<MCMCGPU_FAM_NORMAL_FN>(A, st);
    case MCMCGPU_FAM_NORMAL_DSL: return temp_d<MCMCGPU_FAM_NORMAL_DSL>(A, st);
    case MCMCGPU_FAM_ABS_NORMAL: return temp_d<MCMCGPU_FAM_ABS_NORMAL>(A, st);
  }
  return cudaErrorInvalidValue;
}

// ---- SerialTempMC over arbitrary models: replicas regrouped by the task they consume at this step ----------------------
// (the mutation itself is a one-step run of the wave engine per task, engine.cu: mcmcgpu_run_serialtemp_models)
__device__ __forceinline__ double temp_uniform(const TempMArgs& A, const double* inj, int64_t stride, int64_t r, int64_t col, uint32_t block) {
  if (inj) return inj[r * stride + col];
  const uint64_t g = (uint64_t)(A.rep_offset + r);
  u4 o = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)col, block, (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
  return u01(o.x, o.y);
}

__global__ void temp_plan_kernel(const TempMArgs A, int64_t i, int swap_step) {
  // SerialTempMC.jl:57-63 / :68: which task this replica consumes at iteration i, started at s.pars
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.nrep) return;
  const int at = A.at[r];
  if (swap_step) {
    int at2 = (int)floor(temp_uniform(A, A.inj_pick, A.steps + 1, r, i, 0xFFFFFFFDu) * (double)(A.nt - 1));   // :59
    if (at2 > A.nt - 2) at2 = A.nt - 2;
    if (at2 >= at) at2 += 1;                                                                                  // :60
    A.sel[r] = at2;                      // reset to s.pars (unchanged), consume                               // :62-63
  } else {
    for (int j = 0; j < A.d; j++) A.pars[j * A.Rp + r] = A.state[j * A.Rp + r];
    A.sel[r] = at;
  }
}

__global__ void temp_gather_kernel(const TempMArgs A, const int64_t* __restrict__ idx, int64_t n, int64_t Cp, double* __restrict__ start,
                                   int64_t* __restrict__ chain_ids) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Cp) return;
  const int64_t r = (k < n) ? idx[k] : -1;
  for (int j = 0; j < A.d; j++) start[j * Cp + k] = (r >= 0) ? A.pars[j * A.Rp + r] : 0.0;
  chain_ids[k] = (r >= 0) ? A.rep_offset + r : 0;
}

__global__ void temp_scatter_kernel(const TempMArgs A, const int64_t* __restrict__ idx, int64_t n, int64_t Cp, const double* __restrict__ ppars,
                                    const double* __restrict__ lt0) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t r = idx[k];
  for (int j = 0; j < A.d; j++) A.res_pp[j * A.Rp + r] = ppars[j * Cp + k];
  A.res_lt0[r] = lt0[k];
}

__global__ void temp_update_kernel(const TempMArgs A, int64_t i, int swap_step) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.nrep) return;
  const int d = A.d;
  if (swap_step) {
    const double lt02 = A.res_lt0[r];
    if (temp_uniform(A, A.inj_swap, A.steps + 1, r, i, 0xFFFFFFFCu) < exp(A.logtarget[r] - lt02 + 0.0 - 0.0)) {   // :64; logW stays 0 (:48,:70)
      A.at[r] = A.sel[r];                                                                                         // :65
      for (int j = 0; j < d; j++) { const double v = A.res_pp[j * A.Rp + r]; A.ppars[j * A.Rp + r] = v; A.state[j * A.Rp + r] = v; }
      A.logtarget[r] = lt02;
    }
  } else {                                                                                                        // :68-71
    for (int j = 0; j < d; j++) { const double v = A.res_pp[j * A.Rp + r]; A.ppars[j * A.Rp + r] = v; A.state[j * A.Rp + r] = v; }
    A.logtarget[r] = A.res_lt0[r];
  }
  if (i > A.burnin) {                                                                                             // :73-76
    const int64_t S = A.steps - A.burnin, k = i - A.burnin - 1;
    for (int j = 0; j < d; j++) A.samples[(r * S + k) * d + j] = A.ppars[j * A.Rp + r];
    if (A.at_out) A.at_out[r * S + k] = A.at[r];
  }
}

cudaError_t launch_temp_plan(const TempMArgs& A, int64_t i, bool swap_step, cudaStream_t st) {
  temp_plan_kernel<<<(unsigned)((A.nrep + 127) / 128), 128, 0, st>>>(A, i, swap_step ? 1 : 0);
  return cudaGetLastError();
}
cudaError_t launch_temp_gather(const TempMArgs& A, const int64_t* idx, int64_t n, int64_t Cp, double* start, int64_t* chain_ids, cudaStream_t st) {
  temp_gather_kernel<<<(unsigned)((Cp + 127) / 128), 128, 0, st>>>(A, idx, n, Cp, start, chain_ids);
  return cudaGetLastError();
}
cudaError_t launch_temp_scatter(const TempMArgs& A, const int64_t* idx, int64_t n, int64_t Cp, const double* ppars, const double* lt0, cudaStream_t st) {
  temp_scatter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A, idx, n, Cp, ppars, lt0);
  return cudaGetLastError();
}
cudaError_t launch_temp_update(const TempMArgs& A, int64_t i, bool swap_step, cudaStream_t st) {
  temp_update_kernel<<<(unsigned)((A.nrep + 127) / 128), 128, 0, st>>>(A, i, swap_step ? 1 : 0);
  return cudaGetLastError();
}

bool pop_supported(int family, int64_t d) {
  return (family == MCMCGPU_FAM_NORMAL_FN || family == MCMCGPU_FAM_NORMAL_DSL || family == MCMCGPU_FAM_ABS_NORMAL) && d >= 1 && d <= 8;
}

}  // namespace mg
