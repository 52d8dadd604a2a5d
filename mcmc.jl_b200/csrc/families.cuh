// families.cuh -- closed-form likelihood families evaluated inside a thread (engine FUSED and the
// closed-form evaluation kernel of engine WAVE).  Formulas and LLAcc support semantics as in
// README.md:60-72, examples/ornstein.jl:19-27, src/dsl/definitions/MCMCDerivRules.jl:57-64,
// src/dsl/definitions/AccumulatorDerivRules.jl:10-20, src/dsl/modelparser.jl:64-92.
#pragma once
#include "common.cuh"
namespace mg {

template <int FAM, int D>
struct Family;

// README.md:60,63: v -> -dot(v,v); grad v -> -2v
template <int D>
struct Family<MCMCGPU_FAM_NORMAL_FN, D> {
  static __device__ __forceinline__ double evalallg(const ModelDev& M, const double*, int d, const double (&v)[D],
                                                    double (&g)[D]) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) s += v[j] * v[j];
#pragma unroll
    for (int j = 0; j < D; j++) g[j] = -2.0 * v[j];
    return -s;
  }
  // interior leapfrogs: only the gradient is used (HMC.jl:97 evaluates both; the value is never read)
  static __device__ __forceinline__ double gradonly(const ModelDev&, const double*, int, const double (&v)[D], double (&g)[D]) {
#pragma unroll
    for (int j = 0; j < D; j++) g[j] = -2.0 * v[j];
    return CUDART_NAN;
  }
};

// README.md:67-72 `v ~ Normal(mu, sigma)`; gradient rule MCMCDerivRules.jl:57; LLAcc check
// AccumulatorDerivRules.jl:10-20 => (-Inf, zeros) (modelparser.jl:64-72)
template <int D>
struct Family<MCMCGPU_FAM_NORMAL_DSL, D> {
  static __device__ __forceinline__ double gradonly(const ModelDev& M, const double* x, int d, const double (&v)[D], double (&g)[D]) {
    return evalallg(M, x, d, v, g);   // the LLAcc support test needs the value
  }
  static __device__ __forceinline__ double evalallg(const ModelDev& M, const double*, int d, const double (&v)[D],
                                                    double (&g)[D]) {
    const double mu = M.hyper[0], sigma = M.hyper[1];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) s += logpdf_normal(v[j], mu, sigma);
    double acc = 0.0 + s;
    if (!isfinite(acc)) {
#pragma unroll
      for (int j = 0; j < D; j++) g[j] = 0.0;
      return -CUDART_INF;
    }
#pragma unroll
    for (int j = 0; j < D; j++) g[j] = (j < d) ? (mu - v[j]) / (sigma * sigma) : 0.0;
    return acc;
  }
};

// README.md:253-259 `y = abs(x); y ~ Normal(mu, sigma)`: abs rule dx += sign(x)*ds, Normal rule MCMCDerivRules.jl:57
template <int D>
struct Family<MCMCGPU_FAM_ABS_NORMAL, D> {
  static __device__ __forceinline__ double gradonly(const ModelDev& M, const double* x, int d, const double (&v)[D], double (&g)[D]) {
    return evalallg(M, x, d, v, g);   // the LLAcc support test needs the value
  }
  static __device__ __forceinline__ double evalallg(const ModelDev& M, const double*, int d, const double (&v)[D],
                                                    double (&g)[D]) {
    const double mu = M.hyper[0], sigma = M.hyper[1];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) if (j < d) s += logpdf_normal(fabs(v[j]), mu, sigma);
    double acc = 0.0 + s;
    if (!isfinite(acc)) {
#pragma unroll
      for (int j = 0; j < D; j++) g[j] = 0.0;
      return -CUDART_INF;
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
      double sg = (v[j] > 0.0) ? 1.0 : ((v[j] < 0.0) ? -1.0 : 0.0);
      g[j] = (j < d) ? sg * ((mu - fabs(v[j])) / (sigma * sigma)) : 0.0;
    }
    return acc;
  }
};

// examples/ornstein.jl:19-27; parameter vector (tau, sigma, mu); series staged in shared memory
template <int D>
struct Family<MCMCGPU_FAM_OU, D> {
  static __device__ __forceinline__ double gradonly(const ModelDev& M, const double* x, int d, const double (&v)[D], double (&g)[D]) {
    return evalallg(M, x, d, v, g);   // the LLAcc support test needs the value
  }
  static __device__ __forceinline__ double evalallg(const ModelDev& M, const double* x, int, const double (&v)[D],
                                                    double (&g)[D]) {
    static_assert(D >= 3, "OU has 3 parameters");
    const double tau = v[0], sigma = v[1], mu = v[2];
#pragma unroll
    for (int j = 0; j < D; j++) g[j] = 0.0;
    double acc = 0.0;
    acc = acc + logpdf_uniform(tau, 0.0, M.hyper[0]);   if (!isfinite(acc)) return -CUDART_INF;
    acc = acc + logpdf_uniform(sigma, 0.0, M.hyper[1]); if (!isfinite(acc)) return -CUDART_INF;
    acc = acc + logpdf_uniform(mu, 0.0, M.hyper[2]);    if (!isfinite(acc)) return -CUDART_INF;
    const double fac = exp(-1.0 / tau);
    const double omf = 1.0 - fac;
    const double lsig = log(sigma);
    double s = 0.0, dfac = 0.0, dsigma = 0.0, dmu = 0.0;
    const int64_t T = M.N;
    double xt = x[0];
    for (int64_t t = 0; t + 1 < T; t++) {
      double xn = x[t + 1];
      double resid = xn - xt * fac - mu * omf;
      double z = (resid - 0.0) / sigma;
      s += -(MG_LN_SQRT_2PI + 0.5 * z * z + lsig);
      double dres = (0.0 - resid) / (sigma * sigma);
      dsigma += ((resid - 0.0) * (resid - 0.0) / (sigma * sigma) - 1.0) / sigma;
      dfac += dres * (mu - xt);
      dmu += dres * (-omf);
      xt = xn;
    }
    acc = acc + s;
    if (!isfinite(acc)) return -CUDART_INF;
    g[0] = dfac * fac * (1.0 / (tau * tau));
    g[1] = dsigma;
    g[2] = dmu;
    return acc;
  }
};

}  // namespace mg
