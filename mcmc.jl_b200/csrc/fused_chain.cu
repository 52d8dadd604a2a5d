// fused_chain.cu -- engine "FUSED": one thread runs one whole chain (all MCMC steps) for the
// closed-form likelihood families (Normal via function, Normal via DSL, Ornstein-Uhlenbeck).
// Proposal generation (Philox4x32-10 + Box-Muller, or host-injected draws), the leapfrog / Langevin /
// random-walk move, the Metropolis test, EmpMCTuner / dual-averaging adaptation and the SerialMC
// keep/thin logic all stay in registers; the only global traffic is the kept draws, written
// chain-minor ([kept][param][chain]) so every store instruction is fully coalesced.
//
// Reference loop bodies restated here: src/samplers/RWM.jl:43-72, MALA.jl:65-126, HMC.jl:81-175,
// HMCDA.jl:51-143; keep/thin logic src/runners/SerialMC.jl:37-85.  Compiled with -fmad=false.
#include "common.cuh"
#include "fused_chain.h"
#include "families.cuh"

namespace mg {

template <int D>
__device__ __forceinline__ double dotd(const double (&a)[D], int d) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < D; j++) if (j < d) s += a[j] * a[j];
  return s;
}

// (the streaming variant carries 7 D more doubles: capped at the 146 registers that keep 7 CTAs = 14 warps on an SM, the
//  residency 65 536 chains need for a single round of CTAs)
template <int FAM, int D, bool STREAM>
__global__ void __launch_bounds__(FUSED_THREADS, (STREAM && D <= 4) ? 7 : 0) fused_chain_kernel(const FusedArgs A) {
  extern __shared__ double sh_series[];
  const ModelDev& M = A.M;
  const SamplerDev& S = A.S;
  const RunnerDev& R = A.R;
  constexpr int d = D;            // launch_d instantiates every d up to FUSED_MAX_D: no runtime dimension tests in the step loop
  if (FAM == MCMCGPU_FAM_OU) {
    for (int64_t t = threadIdx.x; t < M.N; t += blockDim.x) sh_series[t] = M.series[t];
    __syncthreads();
  }
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= R.C) return;
  const int64_t Cp = R.Cp;
  const uint64_t gchain = (uint64_t)(R.chain_offset + c);
  const int64_t burnin = R.first - 1;

  auto draw_normals = [&](int64_t i, double (&z)[D]) {
    if (A.inj_normals) {
#pragma unroll
      for (int j = 0; j < D; j++) z[j] = (j < d) ? A.inj_normals[(i * d + j) * Cp + c] : 0.0;
    } else {
#pragma unroll
      for (int b = 0; 2 * b < D; b++) {
        double z0 = 0.0, z1 = 0.0;
        if (2 * b < d) philox_normal_pair(R.seed, gchain, (uint32_t)i, (uint32_t)b, z0, z1);
        z[2 * b] = z0;
        if (2 * b + 1 < D) z[2 * b + 1] = (2 * b + 1 < d) ? z1 : 0.0;
      }
    }
  };
  auto draw_uniform = [&](int64_t i) -> double {
    return A.inj_uniforms ? A.inj_uniforms[i * Cp + c] : philox_uniform(R.seed, gchain, (uint32_t)i);
  };
  int64_t kept = 0;
  // keep / thin bookkeeping without a 64-bit modulo or multiply per step: the step index of the next kept draw and the running
  // offsets of its slots (i runs 1, 2, ...: `i == next_keep` is in(i, first:step:last), SerialMC.jl:49)
  int64_t next_keep = R.first;
  int64_t off_v = c, off_s = c;                 // offsets into the [S][d][Cp] and [S][Cp] arrays
  const int64_t row_v = (int64_t)d * Cp;
  // streamed summaries (stream_stats): per parameter the serial sum of the draws (the mean of src/stats/mean.jl:6, bit for bit),
  // one-pass sums shifted by the first kept draw for Base.var, and the batch sums of mcvar_bm (var.jl:20-26)
  // (a template parameter: the 7 D accumulators must not cost the draw-storing kernel its registers)
  constexpr bool streaming = STREAM;
  constexpr int DS = STREAM ? D : 1;
  double st_k[DS], st_sx[DS], st_sy[DS], st_syy[DS], st_bs[DS], st_sb[DS], st_sbb[DS];
  int64_t st_inb = 0, st_nb = 0, st_acc = 0;
  const int64_t st_nbmax = streaming ? R.S / A.stream_batchlen : 0;
#pragma unroll
  for (int j = 0; j < DS; j++) { st_k[j] = 0.0; st_sx[j] = 0.0; st_sy[j] = 0.0; st_syy[j] = 0.0; st_bs[j] = 0.0; st_sb[j] = 0.0; st_sbb[j] = 0.0; }
  auto store = [&](int64_t i, const double (&pp)[D], double plt, const double (&pg)[D], bool has_grad, bool acc,
                   double eps, int nl) {
    if (i != next_keep || i > R.last) return;
    next_keep += R.step;
    if (streaming) {
      if (kept == 0) {
#pragma unroll
        for (int j = 0; j < DS; j++) st_k[j] = pp[j];
      }
      const bool inbatch = st_nb < st_nbmax;
#pragma unroll
      for (int j = 0; j < DS; j++) {
        const double y = pp[j] - st_k[j];
        st_sx[j] += pp[j];
        st_sy[j] += y;
        st_syy[j] += y * y;
        if (inbatch) st_bs[j] += y;
      }
      if (inbatch && ++st_inb == A.stream_batchlen) {
#pragma unroll
        for (int j = 0; j < DS; j++) { const double b = st_bs[j] / (double)A.stream_batchlen; st_sb[j] += b; st_sbb[j] += b * b; st_bs[j] = 0.0; }
        st_inb = 0; st_nb++;
      }
      st_acc += acc ? 1 : 0;
      if (A.eps) A.eps[off_s] = eps;
      if (A.nleaps) A.nleaps[off_s] = nl;
      kept++; off_s += Cp;
    } else {
#pragma unroll
      for (int j = 0; j < D; j++) if (j < d) A.samples[off_v + (int64_t)j * Cp] = pp[j];
      if (A.grads) {
#pragma unroll
        for (int j = 0; j < D; j++) if (j < d) A.grads[off_v + (int64_t)j * Cp] = has_grad ? pg[j] : CUDART_NAN;
      }
      A.accept[off_s] = acc ? 1 : 0;
      if (A.logtarget) A.logtarget[off_s] = plt;
      if (A.eps) A.eps[off_s] = eps;
      if (A.nleaps) A.nleaps[off_s] = nl;
      kept++; off_v += row_v; off_s += Cp;
    }
  };

  double pars[D], grad[D];
#pragma unroll
  for (int j = 0; j < D; j++) pars[j] = (j < d) ? (R.init_per_chain ? A.init[j * Cp + c] : A.init[j]) : 0.0;
  long long nev = 0;
  double lt = Family<FAM, D>::evalallg(M, sh_series, d, pars, grad);
  nev++;
  if (!isfinite(lt)) {  // "Initial values out of model support" RWM.jl:55 MALA.jl:85 HMC.jl:121 HMCDA.jl:88
    A.status[c] = 1;
    return;
  }
  A.status[c] = 0;

  if (S.kind == MCMCGPU_RWM) {
    // RWM.jl:43-72
    double sc[D];
#pragma unroll
    for (int j = 0; j < D; j++) sc[j] = (j < d) ? A.scale[j] * S.scale : 0.0;  // :52
    for (int64_t i = 1; i <= R.last; i++) {
      double z[D], prop[D], pg[D];
      draw_normals(i, z);
#pragma unroll
      for (int j = 0; j < D; j++) prop[j] = pars[j] + z[j] * sc[j];            // :59
      double plt = Family<FAM, D>::evalallg(M, sh_series, d, prop, pg);         // :60 (eval; gradient unused)
      nev++;
      double ratio = plt - lt;                                                  // :62
      bool acc = ratio > 0;                                                     // :63 `ratio > 0 || ratio > log(rand())`
      if (!acc) acc = ratio > log_lean_normal(draw_uniform(i));
      if (acc) {
        store(i, prop, plt, pg, false, true, CUDART_NAN, 0);
#pragma unroll
        for (int j = 0; j < D; j++) pars[j] = prop[j];
        lt = plt;
      } else {
        store(i, pars, lt, pg, false, false, CUDART_NAN, 0);
      }
    }
  } else if (S.kind == MCMCGPU_MALA) {
    // MALA.jl:65-126
    double tune_step = S.scale;
    long long accepted = 0, proposed = 0;
    for (int64_t i = 1; i <= R.last; i++) {
      double h;
      if (S.tuner_on) { proposed += 1; h = tune_step; } else h = S.scale;       // :91-96
      double z[D], mean[D], prop[D], pg[D];
      draw_normals(i, z);
      const double sq = sqrt(h);
#pragma unroll
      for (int j = 0; j < D; j++) mean[j] = pars[j] + (h / 2.0) * grad[j];      // :98
#pragma unroll
      for (int j = 0; j < D; j++) prop[j] = mean[j] + sq * z[j];                // :100
      double plt = Family<FAM, D>::evalallg(M, sh_series, d, prop, pg);         // :101
      nev++;
      const double lc = log(MG_TWO_PI * h) / 2.0;
      double qno = 0.0;                                                         // :103
#pragma unroll
      for (int j = 0; j < D; j++) if (j < d) { double t = mean[j] - prop[j]; qno += -(t * t) / (2.0 * h) - lc; }
#pragma unroll
      for (int j = 0; j < D; j++) mean[j] = prop[j] + (h / 2.0) * pg[j];        // :104
      double qon = 0.0;                                                         // :105
#pragma unroll
      for (int j = 0; j < D; j++) if (j < d) { double t = mean[j] - pars[j]; qon += -(t * t) / (2.0 * h) - lc; }
      double ratio = plt + qon - lt - qno;                                      // :107
      bool acc = ratio > 0;                                                     // :108
      if (!acc) acc = ratio > log_lean_normal(draw_uniform(i));
      if (acc) {
        store(i, prop, plt, pg, true, true, h, 0);
#pragma unroll
        for (int j = 0; j < D; j++) { pars[j] = prop[j]; grad[j] = pg[j]; }
        lt = plt;
        if (S.tuner_on) accepted += 1;
      } else {
        store(i, pars, lt, grad, true, false, h, 0);
      }
      if (S.tuner_on && i <= burnin && (i % S.adapt_step) == 0) {               // :116-118, adapt! :36-39
        double rate = (double)accepted / (double)proposed;
        tune_step *= (1.0 / (1.0 + exp(-11.0 * (rate - S.target_rate))) + 0.5);
        accepted = 0; proposed = 0;
      }
    }
    if (A.final_eps) A.final_eps[c] = S.tuner_on ? tune_step : S.scale;
  } else if (S.kind == MCMCGPU_RAM) {
    // RAM.jl:41-80 (Vihola 2012): proposal pars + S*rvec, then S = chol(S (I + eta (alpha - rate) r r'/|r|^2) S')'
    double Sm[D][D], Am[D][D], Bm[D][D];
#pragma unroll
    for (int a = 0; a < D; a++)
#pragma unroll
      for (int b = 0; b < D; b++) Sm[a][b] = (a == b && a < d) ? A.scale[a] * S.scale : 0.0;   // :50,55
    for (int64_t i = 1; i <= R.last; i++) {
      double z[D], prop[D], pg[D];
      draw_normals(i, z);                                                       // :59
#pragma unroll
      for (int a = 0; a < D; a++) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < D; b++) if (b < d) acc += Sm[a][b] * z[b];
        prop[a] = pars[a] + acc;                                                // :60
      }
      double plt = Family<FAM, D>::evalallg(M, sh_series, d, prop, pg);         // :61
      nev++;
      double ratio = plt - lt;                                                  // :63
      double tr = 0.0;
#pragma unroll
      for (int a = 0; a < D; a++) if (a < d) tr += Sm[a][a];                    // "scale" => trace(S)
      bool acc = ratio > 0 || ratio > log_lean_normal(draw_uniform(i));                // :64
      if (acc) {
        store(i, prop, plt, pg, false, true, tr, 0);
#pragma unroll
        for (int j = 0; j < D; j++) pars[j] = prop[j];
        lt = plt;
      } else {
        store(i, pars, lt, pg, false, false, tr, 0);
      }
      double eta = (double)d * pow((double)i, -2.0 / 3.0);                      // :74
      if (!(eta < 1.0)) eta = 1.0;
      const double er = exp(ratio);
      const double al = isnan(er) ? 0.0 : (er < 1.0 ? er : 1.0);               // min(1, exp(ratio)); NaN => 0 (documented)
      double zz = 0.0;
#pragma unroll
      for (int a = 0; a < D; a++) if (a < d) zz += z[a] * z[a];
#pragma unroll
      for (int a = 0; a < D; a++)
#pragma unroll
        for (int b = 0; b < D; b++) Am[a][b] = ((a == b) ? 1.0 : 0.0) + (z[a] * z[b]) / zz * eta * (al - S.rate);   // :76
#pragma unroll
      for (int a = 0; a < D; a++)
#pragma unroll
        for (int b = 0; b < D; b++) {                                           // S * (I + SS)
          double s2 = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) if (k < d) s2 += Sm[a][k] * Am[k][b];
          Bm[a][b] = s2;
        }
#pragma unroll
      for (int a = 0; a < D; a++)
#pragma unroll
        for (int b = 0; b < D; b++) {                                           // ... * S'   (:77)
          double s2 = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) if (k < d) s2 += Bm[a][k] * Sm[b][k];
          Am[a][b] = s2;
        }
#pragma unroll
      for (int a = 0; a < D; a++)                                               // S = chol(SS)' (:78), lower factor row by row
#pragma unroll
        for (int b = 0; b < D; b++) {
          if (b > a || a >= d) { Sm[a][b] = 0.0; continue; }
          double s2 = Am[a][b];
#pragma unroll
          for (int k = 0; k < D; k++) if (k < b) s2 -= Sm[a][k] * Sm[b][k];
          Sm[a][b] = (a == b) ? sqrt(s2) : s2 / Sm[b][b];
        }
    }
  } else {
    // HMC.jl:106-175 and HMCDA.jl:72-143 share HMCSample / leapfrog (HMC.jl:81-102)
    const bool da = (S.kind == MCMCGPU_HMCDA);
    long long t_nleaps = S.nleaps; double t_step = S.scale; long long accepted = 0, proposed = 0;
    double leapStepDA = 1.0, mu = 0.0, dualLeapStep = 1.0, dualH = 0.0;
    if (da) {
      // HMCDA.jl:90-94.  The pre-loop randn (step-0 draw) and the leapfrog inside initializeHMCDAStep
      // (HMCDA.jl:51-69) have no observable effect: state0.H is NaN there (HMC.jl:88), so p = NaN,
      // a = -1, the while test is false and the function always returns 1.0.
      leapStepDA = 1.0;
      mu = log(10.0 * leapStepDA);
    }
    for (int64_t i = 1; i <= R.last; i++) {
      long long nLeaps; double eps;
      if (da) {
        eps = leapStepDA;
        double nl = round(S.len / eps);                                         // HMCDA.jl:104
        if (!(nl >= 1.0)) nl = 1.0;
        if (nl > (double)S.max_leaps) nl = (double)S.max_leaps;
        nLeaps = (long long)nl;
      } else if (S.tuner_on) { proposed += 1; nLeaps = t_nleaps; eps = t_step; } // HMC.jl:129-134
      else { nLeaps = S.nleaps; eps = S.scale; }
      double m[D], p[D], g[D];
      draw_normals(i, m);                                                       // HMC.jl:136 / HMCDA.jl:100
      const double H0 = -lt + 0.5 * dotd<D>(m, d);                              // update! HMC.jl:91
      double plt = lt;
#pragma unroll
      for (int j = 0; j < D; j++) { p[j] = pars[j]; g[j] = grad[j]; }
      double rbacc[D];
#pragma unroll
      for (int j = 0; j < D; j++) rbacc[j] = 0.0;
      if (FAM == MCMCGPU_FAM_NORMAL_FN && !A.rb) {
        // grad = -2 p, and scaling by a power of two is exact: (0.5 * grad) * eps == (-p) * eps bit for bit, so the
        // trajectory needs neither the gradient nor its halving (6 instead of 9 FP64 instructions per dimension and
        // leapfrog); value and gradient are formed once at the end point
        for (long long l = 0; l < nLeaps; l++) {                                // leapfrog HMC.jl:93-102
#pragma unroll
          for (int j = 0; j < D; j++) m[j] += (-p[j]) * eps;
#pragma unroll
          for (int j = 0; j < D; j++) p[j] += eps * m[j];
#pragma unroll
          for (int j = 0; j < D; j++) m[j] += (-p[j]) * eps;
        }
        nev += nLeaps;
        plt = Family<FAM, D>::evalallg(M, sh_series, d, p, g);
      } else
      for (long long l = 0; l < nLeaps; l++) {                                  // leapfrog HMC.jl:93-102
#pragma unroll
        for (int j = 0; j < D; j++) m[j] += (0.5 * g[j]) * eps;
#pragma unroll
        for (int j = 0; j < D; j++) p[j] += eps * m[j];
        if (l + 1 == nLeaps || A.rb) plt = Family<FAM, D>::evalallg(M, sh_series, d, p, g);
        else Family<FAM, D>::gradonly(M, sh_series, d, p, g);
        nev++;
#pragma unroll
        for (int j = 0; j < D; j++) m[j] += (0.5 * g[j]) * eps;
        if (A.rb) {                                                             // storeLeaps: w = exp(H0 - H_leap) (mean.jl:19)
          const double wl = exp(H0 - (-plt + 0.5 * dotd<D>(m, d)));
#pragma unroll
          for (int j = 0; j < D; j++) rbacc[j] += wl * p[j];                    // mean.jl:27
        }
      }
      const double H = -plt + 0.5 * dotd<D>(m, d);
      const double e = exp(H0 - H);
      const double u = draw_uniform(i);
      bool acc; double pacc = 0.0;
      if (da) { pacc = isnan(e) ? 0.0 : (e < 1.0 ? e : 1.0); acc = u < pacc; }  // HMCDA.jl:120-121 (NaN => 0: documented)
      else acc = u < e;                                                          // HMC.jl:154
      if (A.rb && i == next_keep && i <= R.last) {                             // (sample + sum_k w_k pars_k)/(nleaps+1), mean.jl:24-30
#pragma unroll
        for (int j = 0; j < D; j++) if (j < d) A.rb[off_v + (int64_t)j * Cp] = ((acc ? p[j] : pars[j]) + rbacc[j]) / (double)(nLeaps + 1);
      }
      if (acc) {
        store(i, p, plt, g, true, true, eps, (int)nLeaps);
#pragma unroll
        for (int j = 0; j < D; j++) { pars[j] = p[j]; grad[j] = g[j]; }
        lt = plt;
        if (S.tuner_on) accepted += 1;
      } else {
        store(i, pars, lt, grad, true, false, eps, (int)nLeaps);
      }
      if (da) {
        if (i < burnin) {                                                       // HMCDA.jl:133-138
          double fi = (double)i;
          double eta = 1.0 / (fi + S.t0);
          dualH = (1.0 - eta) * dualH + eta * (S.rate - pacc);
          leapStepDA = exp(mu - sqrt(fi) * dualH / S.shrinkage);
          eta = pow(fi, -S.step);
          dualLeapStep = exp((1.0 - eta) * log(dualLeapStep) + eta * log(leapStepDA));
        } else {
          leapStepDA = dualLeapStep;                                            // :140
        }
      } else if (S.tuner_on && i <= burnin && (i % S.adapt_step) == 0) {        // HMC.jl:167-169, adapt! :39-43
        double rate = (double)accepted / (double)proposed;
        t_step *= (1.0 / (1.0 + exp(-11.0 * (rate - S.target_rate))) + 0.5);
        double cl = ceil(S.target_path / t_step);
        t_nleaps = (cl < (double)S.max_step) ? (long long)cl : (long long)S.max_step;
        accepted = 0; proposed = 0;
      }
    }
    if (A.final_eps) A.final_eps[c] = da ? leapStepDA : (S.tuner_on ? t_step : S.scale);
  }
  if (streaming && kept >= 2) {
    const double n = (double)kept, plane = 0.0;
    (void)plane;
    const int64_t P = (int64_t)d * Cp;
#pragma unroll
    for (int j = 0; j < DS; j++) if (j < d) {
      const int64_t o = (int64_t)j * Cp + c;
      const double viid = ((st_syy[j] - st_sy[j] * st_sy[j] / n) / (n - 1.0)) / n;                  // var.jl:7-8
      double vbm = CUDART_NAN;
      if (st_nb > 1) {
        const double nb = (double)st_nb;
        const double vb = (st_sbb[j] - st_sb[j] * st_sb[j] / nb) / (nb - 1.0);                      // var of the batch means
        vbm = (double)A.stream_batchlen * vb / (nb * (double)A.stream_batchlen);                    // var.jl:25
      }
      A.stream[o] = st_sx[j] / n;                                                                   // mean.jl:6
      A.stream[P + o] = viid;
      A.stream[2 * P + o] = vbm;
      A.stream[3 * P + o] = n * viid / vbm;                                                         // ess.jl:9
      A.stream[4 * P + o] = vbm / viid;                                                             // ess.jl:18
    }
    A.stream_accept[c] = (double)st_acc * 100.0 / n;                                                // summary.jl:13
  }
  // final state (resume) + evaluation count
  if (A.final_pars) {
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) A.final_pars[j * Cp + c] = pars[j];
  }
  atomicAdd(A.n_evals, (unsigned long long)nev);
}

template <int FAM, int D>
static cudaError_t launch_one(const FusedArgs& A, cudaStream_t st) {
  size_t smem = (FAM == MCMCGPU_FAM_OU) ? sizeof(double) * (size_t)A.M.N : 0;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fused_chain_kernel<FAM, D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_chain_kernel<FAM, D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int blocks = (int)((A.R.C + FUSED_THREADS - 1) / FUSED_THREADS);
  if (A.stream) fused_chain_kernel<FAM, D, true><<<blocks, FUSED_THREADS, smem, st>>>(A);
  else fused_chain_kernel<FAM, D, false><<<blocks, FUSED_THREADS, smem, st>>>(A);
  return cudaGetLastError();
}

template <int FAM>
static cudaError_t launch_d(const FusedArgs& A, cudaStream_t st) {
  int d = (int)A.M.d;
  if (d <= 1) return launch_one<FAM, 1>(A, st);
  if (d <= 2) return launch_one<FAM, 2>(A, st);
  if (d <= 3) return launch_one<FAM, 3>(A, st);
  if (d <= 4) return launch_one<FAM, 4>(A, st);
  if (d <= 5) return launch_one<FAM, 5>(A, st);
  if (d <= 6) return launch_one<FAM, 6>(A, st);
  if (d <= 7) return launch_one<FAM, 7>(A, st);
  if (d <= 8) return launch_one<FAM, 8>(A, st);
  return cudaErrorInvalidValue;
}

bool fused_supported(int family, int64_t d, int64_t N) {
  if (family == MCMCGPU_FAM_NORMAL_FN || family == MCMCGPU_FAM_NORMAL_DSL || family == MCMCGPU_FAM_ABS_NORMAL)
    return d >= 1 && d <= FUSED_MAX_D;
  if (family == MCMCGPU_FAM_OU) return d == 3 && N >= 2 && N * 8 <= 200 * 1024;
  return false;
}

cudaError_t launch_fused(const FusedArgs& A, cudaStream_t st) {
  switch (A.M.family) {
    case MCMCGPU_FAM_NORMAL_FN: return launch_d<MCMCGPU_FAM_NORMAL_FN>(A, st);
    case MCMCGPU_FAM_NORMAL_DSL: return launch_d<MCMCGPU_FAM_NORMAL_DSL>(A, st);
    case MCMCGPU_FAM_ABS_NORMAL: return launch_d<MCMCGPU_FAM_ABS_NORMAL>(A, st);
    case MCMCGPU_FAM_OU: return launch_one<MCMCGPU_FAM_OU, 3>(A, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace mg
