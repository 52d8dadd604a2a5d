// k1_regress.cu -- K1: fused log-likelihood + gradient of the regression families for a tile of
// 64 chains x a range of rows, on the FP64 tensor cores (DMMA m8n8k4), X tiles staged by bulk
// async copies (TMA, cp.async.bulk + mbarrier) through a 4-deep shared-memory ring.
//
// What it replaces: per chain and per evaluation the reference runs `X * vars`
// (dgemv, examples/logistic_regression.jl:18), an elementwise logpdf/logcdf pass with an LLAcc sum
// (DistributionsExtensions.jl:51-58, AccumulatorDerivRules.jl:19-20) and the reverse sweep X' * r
// (MCMCDerivRules.jl:111, probit_regression.jl:39).  Here all chains share each X tile:
//
//   phase 1   eta^T[chain][row] = sum_j beta[chain][j] * X[row][j]        (DMMA, A = beta frags in registers)
//   epilogue  link function in registers: loglik term, r = d loglik / d eta
//   phase 2   G^T[chain][j]    += sum_row r[chain][row] * X[row][j]       (DMMA, A = the phase-1 accumulators)
//
// The phase-1 accumulator registers ARE the phase-2 A fragments: an m8n8k4 C fragment holds columns
// (2t, 2t+1) of row g (g = lane/4, t = lane%4) and an A fragment needs column t, so phase 2 simply
// sums over the rows in the order the registers already hold them (k is a summation index) and loads
// the matching rows of X for the B fragment.  eta and r never leave registers.
//
// Algorithmic work per (row, chain): 4*d flop (2d for eta, 2d for X'r); bound: FP64 tensor pipe.
#include "k1_regress.h"
#include "erfcx_table.h"
#include "probit_table.h"
#include "exp_table.h"
#include "log_table.h"
#include <cmath>
#include <cstdio>
#include <type_traits>

namespace mg {

constexpr int K1_STAGES = 2;
// probit at d <= 32: two resident CTAs (128 registers: no spills in the link, room for the whole central table |z| < 8) and
// a 4-deep ring (a tile lasts ~1300 cycles there, the refill of a slot starts when the SLOWEST warp leaves it); measured on
// cfg3's MALA wave: 3 CTAs x 2 stages 7.42 ms, 3 x 3 7.20, 2 x 3 7.10, 2 x 4 7.03
#ifndef K1_PROBIT_SMALL_STAGES
#define K1_PROBIT_SMALL_STAGES 4
#endif
#ifndef K1_PROBIT_SMALL_CTAS
#define K1_PROBIT_SMALL_CTAS 2
#endif
#ifndef K1_OTHER_SMALL_CTAS
#define K1_OTHER_SMALL_CTAS 3      // linear, logistic at d <= 32
#endif
#ifndef K1_OTHER_SMALL_STAGES
#define K1_OTHER_SMALL_STAGES 2
#endif
constexpr int K1_NR = 4;          // 8-row groups per phase-1 pass (independent DMMA accumulator chains)
constexpr int PH_IDLE_K1 = 98;  // phases >= PH_PAUSE (98) have no pending evaluation (transition.h)
constexpr int PH_LEAP_K1 = 3;   // PH_LEAP (transition.h)
// x / v for the prior gradient, as transition.cu's div_by_var: multiplication by the exact reciprocal when v is a power of
// two (inv and the test on v are formed once per thread; the per-element test on x is integer work on its exponent)
__device__ __forceinline__ bool k1_var_is_pow2(double v) {
  return ((__double_as_longlong(v) & 0x000FFFFFFFFFFFFFll) == 0) && v > 1e-150 && v < 1e150;
}
__device__ __forceinline__ double k1_div_by_var(double x, double v, double inv, bool vpow2) {
  const int ex = (__double2hiint(x) >> 20) & 0x7ff;                 // biased exponent: 2^-498 < |x| < 2^499, or x == 0
  const bool ok = vpow2 && ((ex > 0x20d && ex < 0x5f2) || (__double_as_longlong(x) << 1) == 0);
  if (ok) return __dmul_rn(x, inv);
  return __ddiv_rn(x, v);
}

// row permutation inside an 8-row group: column n of the phase-1 B fragment reads row PI[n]
// (bank-conflict-free for both phases when the row stride is 4 mod 8 doubles)
__device__ __forceinline__ int pi8(int n) { return (n < 4) ? n : (n ^ 1); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// slot-release counter of the tile ring: acq_rel at CTA scope, so every warp's (generic-proxy) reads of the tile are ordered
// before the last arriver's refill (release by each arriving lane 0 after __syncwarp, acquire by the one that sees WARPS-1)
__device__ __forceinline__ unsigned atom_add_acq_rel_cta(unsigned int* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- erfcx(u), 0 <= u < 26: degree-9 piecewise polynomial (tools/gen_erfcx_table.py, 2e-16 relative) ----------
// CUDA's erfcx costs ~100 FP64 instructions; the table costs 10 fused multiply-adds and 10 cached loads (the
// coefficient-major layout keeps a warp's gathers on one or two cache lines because |z| clusters around 0..3).
__device__ const double erfcx_tab[(ERFCX_DEG + 1) * ERFCX_NINT] = {ERFCX_TABLE_VALUES};
__device__ __forceinline__ double erfcx_fast(double u) {       // requires 0 <= u < ERFCX_UMAX
  // nearest grid point k/8 by the magic-number add (no double<->int conversion instructions on the FP64 pipe)
  const double m = fma(u, (double)ERFCX_INV_W, 6755399441055744.0);
  const int k = __double2loint(m);
  const double t = fma(m - 6755399441055744.0, -1.0 / ERFCX_INV_W, u);
  const double* c = erfcx_tab + k;
  double p = __ldg(c + ERFCX_DEG * ERFCX_NINT);
#pragma unroll
  for (int j = ERFCX_DEG - 1; j >= 0; j--) p = fma(p, t, __ldg(c + j * ERFCX_NINT));
  return p;
}

// ---- probit link: F(z) = log Phi(z), W(z) = phi(z)/Phi(z), |z| < 36.9 (tools/gen_probit_table.py) ----------------------
// Round 1 evaluated degree-9/10 piecewise polynomials: 11 table gathers and 25 FP64 instructions per element for value +
// derivative, and at small d the kernel was bound by the load/store unit (ncu, cfg3: LSU wavefronts 89 % of peak).  Behind
// the LSU sits the math dispatch port: a DMMA keeps it busy for ~12 of its 16 cycles (tools/fp64_probe3.cu: 64 FFMAs next to
// 8 DMMAs cost 33 extra cycles, 16 conversion pairs cost 2), so every FP64 (2 cycles), FP32 and integer instruction of the
// link adds to the DMMA time, while conversions and loads overlap with it.  Round 2: the table holds, per grid point
// c = k/128, only F(c) and W(c) in double and a FLOAT pair (e3, e4) -- three gathers --; between grid points, t = z - c:
//   s = c + W, W1 = -W s, W2 = -W1 (s + W) - W                      from W' = -W (z + W), in double (enter with t, t^2)
//   W(c+t) = W + t (W1 + t/2 (W2 + t Tp2)),            Tp2 = e3 + e4 t            in float (enters with t^3 <= 2^-24)
//   F(c+t) = F + t (W + t/2 (W1 + t/3 (W2 + t Tv6))),  Tv6 = 0.75 e3 + 0.6 e4 t = 0.6 Tp2 + 0.15 e3
// 14 FP64 + 6 FP32 instructions for value + derivative (the grid index is float work too), 9 + 3 for the derivative alone.  Error against mpmath (emulated
// operation by operation in the generator): F 2.1e-16 relative (z < 0) / 1.1e-16 absolute (z >= 0), W 7.6e-16 / 1.8e-15.
__device__ const unsigned long long probit_tab[3 * PROBIT_NINT] = {PROBIT_TAB_WORDS};   // [F | W | E], 8-byte words
// The tables sit in global memory (L1-resident); when the CTA has room the central part |z| < H (H = 8, else 4) -- where
// practically every element falls -- is copied to shared memory, where a gather costs one wavefront per group of lanes that
// hit one bank at different addresses instead of one per distinct 32-byte sector.  The copies hold the same numbers and the
// arithmetic is the same, so which copy a warp reads never changes a number.
__host__ __device__ constexpr int ps_k0(int H) { return (PROBIT_ZMAX - H) * PROBIT_INV_W - 1; }   // first grid point of the copy
__host__ __device__ constexpr int ps_n(int H) { return 2 * H * PROBIT_INV_W + 3; }                // grid points -H - 1/128 .. H + 1/128
__host__ __device__ constexpr size_t ps_bytes(int H) { return (size_t)ps_n(H) * 24; }             // F, W (double), E (float2)
// shared-memory bytes of a CTA without the probit tables, and the half-width of the copy the resident CTA count leaves room for
__host__ __device__ constexpr size_t k1_smem_base(int DK, int WARPS, int STAGES) {
  return sizeof(double) * ((size_t)(8 * WARPS) * (8 * DK + 4) + (size_t)STAGES * (K1_ROWS * (8 * DK + 4) + K1_ROWS)) +
         2 * STAGES * sizeof(uint64_t) + (EXP_NTAB + 2 * LOG_NINT) * sizeof(double);
}
__host__ __device__ constexpr int k1_probit_half(int DK, int WARPS, int STAGES, int CTAS) {
  return ((size_t)CTAS * (k1_smem_base(DK, WARPS, STAGES) + ps_bytes(8) + 1024) <= 233472) ? 8      // 228 KB per SM
       : ((size_t)CTAS * (k1_smem_base(DK, WARPS, STAGES) + ps_bytes(4) + 1024) <= 233472) ? 4 : 0;
}
template <bool SM>
__device__ __forceinline__ unsigned long long tab_ld(const unsigned long long* p) { return SM ? *p : __ldg(p); }
// T: the table, rows F | W | E of STRIDE 8-byte words each, positioned so that T[k] is F at grid point k (one address per
// element, the rows at compile-time offsets).  WANT_F / WANT_W: which of the two functions the caller uses (the other
// one's instructions are not generated)
template <bool SM, int STRIDE, bool WANT_F, bool WANT_W>
__device__ __forceinline__ void probit_eval(const unsigned long long* T, double z, double& f, double& w) {
  // grid index in FLOAT (conversions and FP32 instructions are cheaper than FP64 ones here): k = rint(128 z) from the low
  // mantissa bits of 1.5 * 2^23 + 128 z; t = z - k/128 is exact in double whichever neighbour a tie-near z picks
  const float mf = fmaf((float)z, (float)PROBIT_INV_W, 12582912.0f);
  const int k = __float_as_int(mf) - 0x4B400000 + PROBIT_ZMAX * PROBIT_INV_W;
  const double kd = (double)(mf - 12582912.0f);                  // 128 c, exact
  const double t = fma(kd, -1.0 / PROBIT_INV_W, z);               // exact
  const unsigned long long* Tk = T + k;
  const double W0 = __longlong_as_double((long long)tab_ld<SM>(Tk + STRIDE));
  const unsigned long long eb = tab_ld<SM>(Tk + 2 * STRIDE);
  const float e3 = __uint_as_float((unsigned)eb), e4 = __uint_as_float((unsigned)(eb >> 32));
  const double s = fma(kd, 1.0 / PROBIT_INV_W, W0);               // c + W
  const double W1 = -__dmul_rn(W0, s);
  const double W2 = fma(-W1, __dadd_rn(s, W0), -W0);              // -W1 s - W (1 + W1) = -W1 (s + W) - W
  const double th = 0.5 * t;
  const float tf = (float)t;
  const float Tp2 = fmaf(e4, tf, e3);
  if (WANT_W) w = fma(t, fma(th, fma(t, (double)Tp2, W2), W1), W0);
  if (WANT_F) {
    const double F0 = __longlong_as_double((long long)tab_ld<SM>(Tk));
    const double t3 = t * (1.0 / 3.0);
    const float Tv6 = fmaf(0.15f, e3, 0.6f * Tp2);                // 0.75 e3 + 0.6 e4 t
    f = fma(t, fma(th, fma(t3, fma(t, (double)Tv6, W2), W1), W0), F0);
  }
}

// ---- logistic link: exp from a 64-entry table of 2^(j/64) held in shared memory (tools/gen_exp_table.py) ---------------
// DMMA and the scalar FP64 instructions share one pipe, so the link is written for the fewest FP64 instructions:
//   exp(x) = 2^k T[j] (1 + q(r)), n = rint(64 x / ln 2) = 64 k + j, |r| <= ln2/128, q of degree 5: 10 FP64 instructions
//   (exp_lean: 17), 2.2e-16 relative against mpmath; the table read is one LDS.64 (the load/store pipe has room).
__device__ const double exp_tab_g[EXP_NTAB] = {EXP_TABLE_VALUES};
__device__ __forceinline__ double exp_tab64(double x, const double* T) {      // |x| < 700
  const double t = fma(x, 64.0 * 1.4426950408889634, 6755399441055744.0);
  const int n = __double2loint(t);
  const double nf = t - 6755399441055744.0;
  double r = fma(nf, -6.93147180369123816490e-01 / 64.0, x);
  r = fma(nf, -1.90821492927058770002e-10 / 64.0, r);
  double p = fma(8.33333333333333333333e-03, r, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  const double tj = T[n & (EXP_NTAB - 1)];
  const double e = fma(tj, p * r, tj);
  return __hiloint2double(__double2hiint(e) + ((n >> 6) << 20), __double2loint(e));   // * 2^k, |k| <= 1010: stays normal
}
// 1/d, d in [1, 1e305]: rcp.approx.ftz.f64 (relative error <= 2^-23) and one cubic step y (1 + e + e^2), e = 1 - d y
// (remaining error e^3 = 2^-69): 3 FP64 instructions instead of the 4 of two Newton steps
__device__ __forceinline__ double rcp_cubic(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d, y, 1.0);
  return fma(y, fma(e, e, e), y);
}

// log(x) from a 129-interval table of (1/c_j, log c_j) in shared memory (tools/gen_log_table.py): fdlibm's reduction to
// sqrt(2)/2 <= m < sqrt(2), u = m/c_j - 1 by one fused multiply-add, log1p(u) of degree 6: 11 FP64 instructions (log_lean: 28);
// 2.3e-16 relative, 2e-18 absolute where the logarithm vanishes (x -> 1).  Zero, denormal, negative, infinite and NaN
// arguments go to libm's log (they decide the support test).
__device__ const double log_tab_g[2 * LOG_NINT] = {LOG_TABLE_VALUES};
__device__ __forceinline__ double log_tab129(double x, const double2* T) {
  int hx = __double2hiint(x);
  if ((unsigned)(hx - 0x00100000) >= (unsigned)(0x7ff00000 - 0x00100000)) return log(x);
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;
  k += i >> 20;
  const int hm = hx | (i ^ 0x3ff00000);
  const double m = __hiloint2double(hm, __double2loint(x));
  const double2 rl = T[(hm - LOG_BASE_HI) >> LOG_SHIFT];
  const double u = fma(m, rl.x, -1.0);
  double p = fma(-1.0 / 6.0, u, 0.2);
  p = fma(p, u, -0.25);
  p = fma(p, u, 1.0 / 3.0);
  p = fma(p, u, -0.5);
  const double l1 = fma(p, u * u, u);
  const double dk = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;   // (double)k without a conversion
  return fma(dk, 6.93147180369123816490e-01, rl.y) + fma(dk, 1.90821492927058770002e-10, l1);
}

// ---- link functions: (eta, y) -> loglik term(s) and r = d loglik / d eta ----------------------
struct LinkOut { double ll1, ll2, r; bool bad; };

template <int FAM>
__device__ __forceinline__ LinkOut link(double eta, double y, const double* hy, bool need_ll) {
  LinkOut o; o.ll2 = 0.0; o.bad = false;
  if (FAM == MCMCGPU_FAM_LINEAR) {
    // examples/linear_regression.jl:16-17: resid = Y - X*vars; resid ~ Normal(0, sd)
    const double nsd = hy[1];
    double resid = y - eta;
    double z = (resid - 0.0) / nsd;
    o.ll1 = -(MG_LN_SQRT_2PI + 0.5 * z * z + hy[3]);  // hy[3] = log(noise_sd), precomputed
    o.bad = !isfinite(o.ll1);
    o.r = -((0.0 - resid) / (nsd * nsd));             // MCMCDerivRules.jl:57 through resid = Y - X*vars
  } else if (FAM == MCMCGPU_FAM_LOGISTIC) {
    // examples/logistic_regression.jl:18-19: prob = 1/(1+exp(sgn*eta)); Y ~ Bernoulli(prob)
    const double sgn = hy[1];
    double e = exp(sgn * eta);
    double den = 1.0 + e;
    double p = 1.0 / den;
    double omp = 1.0 - p;               // Bernoulli stores p0 = 1 - p1
    bool y1 = (y != 0.0);
    double arg = y1 ? p : omp;
    if (need_ll) { o.ll1 = log(arg); o.bad = !isfinite(o.ll1); }
    else { o.ll1 = 0.0; o.bad = !(arg > 0.0); }   // log(arg) finite <=> arg > 0 (arg <= 1 always; NaN => bad)
    // d/d eta of the term.  The reference's AD chain (MCMCDerivRules.jl:111 through /, +, exp, -)
    // is e*(-(1/(p-1+y))/(den*den))*sgn; algebraically -sgn*e*p (y=1) or sgn*p (y=0); the closed form
    // differs from the chain only by the chain's own rounding (<= eps/(1-p)).
    o.r = y1 ? (-sgn) * (e * p) : sgn * p;
  } else {  // PROBIT, examples/probit_regression.jl:26-41
    const double base = -(eta * eta + MG_LOG2PI) / 2.0;
    if ((y == 1.0 || y == 0.0) && fabs(eta) < 1e100) {
      // binary response, finite tails: the other term is an exact zero
      // one scaled complementary error function gives both log Phi(z) and the Mills-type ratio phi(z)/Phi(z):
      //   z < 0:  Phi = exp(-u^2) erfcx(u) / 2, u = -z/sqrt2  =>  log Phi = -u^2 + log(erfcx(u)/2),  phi/Phi = sqrt(2/pi)/erfcx(u)
      //   z >= 0: Phi = 1 - exp(-u^2) erfcx(u) / 2, u = z/sqrt2
      // (same functions as logcdf(normal, .) and exp(-(eta^2+log 2pi)/2 - logcdf) of probit_regression.jl:29,39-40,
      //  evaluated without the exp/log round trip; agreement with that form is ~1e-14 relative)
      bool y1 = (y == 1.0);
      const double z = y1 ? eta : -eta;
      const double u = fabs(z) * MG_SQRT1_2;
      double l, w;
      if (u < (double)ERFCX_UMAX) {
        // branch-free for |z| < 36.8: c = exp(-u^2) erfcx(u) / 2 is the tail mass; no lane of a warp diverges on the
        // sign of z (the signs are mixed in practically every warp)
        const double ex = erfcx_fast(u);
        const double e2 = exp(-(u * u));
        const double c = 0.5 * e2 * ex;
        const bool neg = z < 0.0;
        const double omc = 1.0 - c;
        l = log(neg ? c : omc);                                    // log Phi(z)
        // phi(z)/Phi(z): sqrt(2/pi)/erfcx(u) for z < 0, phi(z)/(1 - c) otherwise -- one division either way
        w = (neg ? 0.79788456080286535588 : 0.39894228040143267794 * e2) / (neg ? ex : omc);
      } else if (z < 0.0) {
        const double ex = erfcx(u);
        l = -(u * u) + log(0.5 * ex);                              // far lower tail: exp(-u^2) would underflow
        w = 0.79788456080286535588 / ex;
      } else {
        const double e2 = exp(-(u * u));                           // far upper tail: Phi(z) = 1 - (denormal or 0)
        l = log1p(-0.5 * e2 * erfcx(u));
        w = 0.39894228040143267794 * e2;
      }
      (void)base;
      o.ll1 = y1 ? l : 0.0;
      o.ll2 = y1 ? 0.0 : l;
      o.r = y1 ? w : -w;
    } else {
      double lp = log_ndtr(eta), lm = log_ndtr(-eta);
      o.ll1 = lp * y;
      o.ll2 = lm * (1.0 - y);
      o.r = y * exp(base - lp) - (1.0 - y) * exp(base - lm);
    }
  }
  return o;
}

// ---- the kernel -----------------------------------------------------------------------------
// Shared memory per CTA: beta tile [64 chains][S] (A fragments of phase 1), STAGES X tiles, barriers.
// 256 threads, <= 128 registers, two CTAs per SM: 16 warps keep the DMMA pipe fed while other warps
// are in the (latency-bound) link-function epilogue.
// CTA shape: WARPS consumer warps (8 chains each), a STAGES-deep tile ring, CTAS resident CTAs per SM.
template <int FAM, int DK, int NR, int WARPS, int STAGES, int CTAS>
__global__ void __launch_bounds__(WARPS * 32, CTAS) k1_kernel(const K1Args a) {
  constexpr int K1_WARPS = WARPS, K1_CHAINS = 8 * WARPS, K1_THREADS = 32 * WARPS, K1_STAGES = STAGES;   // shadow the defaults
  constexpr int S = 8 * DK + 4;
  constexpr int TILE_D = K1_ROWS * S + K1_ROWS;
  constexpr int NG = K1_ROWS / (8 * NR);      // row groups per tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* betas = reinterpret_cast<double*>(smem_raw);
  double* tiles = betas + K1_CHAINS * S;
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)K1_STAGES * TILE_D);
  unsigned int* released = reinterpret_cast<unsigned int*>(full + K1_STAGES);   // warps that have finished with a slot
  double* etab = reinterpret_cast<double*>(full + 2 * K1_STAGES);               // 2^(j/64) (logistic link)
  double2* ltab = reinterpret_cast<double2*>(etab + EXP_NTAB);                  // (1/c_j, log c_j) (logistic link)
  constexpr int PSH = (FAM == MCMCGPU_FAM_PROBIT) ? k1_probit_half(DK, WARPS, STAGES, CTAS) : 0;   // |z| < PSH sits in shared memory
  constexpr bool PSM = PSH > 0;
  constexpr int PS_N = ps_n(PSH), PS_K0 = ps_k0(PSH);
  unsigned long long* ptab = reinterpret_cast<unsigned long long*>(ltab + LOG_NINT);   // central part of the probit table: F | W | E

  if (a.remaining && *a.remaining == 0) return;
  const bool ybin = (FAM == MCMCGPU_FAM_PROBIT) && a.P.ynb != nullptr && (*a.P.ynb == 0u);   // every response is 0.0 or 1.0
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t chain0 = (int64_t)blockIdx.x * K1_CHAINS;
  const int64_t mychain = chain0 + warp * 8 + g;   // the chain this thread's fragments belong to
  const int split = blockIdx.y;
  const int64_t Cp = a.Cp;
  const int d = (int)a.P.d;

  // skip chain tiles whose chains have no pending evaluation (asynchronous HMCDA trajectories)
  int alive = 1;
  if (a.phase) {
    const int64_t pc = chain0 + (tid & (K1_CHAINS - 1));
    alive = (pc < Cp) && (a.phase[pc] < PH_IDLE_K1);
  }
  if (!__syncthreads_or(alive)) return;

  // tile range of this split
  const int64_t per = (a.P.ntiles + a.nsplit - 1) / a.nsplit;
  const int64_t t0 = (int64_t)split * per;
  int64_t t1 = t0 + per; if (t1 > a.P.ntiles) t1 = a.P.ntiles;
  const int64_t nt = (t1 > t0) ? (t1 - t0) : 0;
  const uint32_t tile_bytes = (uint32_t)(TILE_D * sizeof(double));

  if (tid == 0) {
    for (int s = 0; s < K1_STAGES; s++) { mbar_init(&full[s], 1); released[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // beta tile: betas[c][j] = q[j][chain0 + c] (zero-padded features).  Logistic: the sign convention of the link
  // (hyper[1] = +-1, checked at model_create) is folded into beta, so phase 1 yields x = sign * eta directly; a sign
  // flip commutes with every rounding, so x is bit-identical to sign * (X beta)
  const double bsign = (FAM == MCMCGPU_FAM_LOGISTIC) ? a.hyper[1] : 1.0;
  for (int idx = tid; idx < K1_CHAINS * 8 * DK; idx += K1_THREADS) {
    const int j = idx / K1_CHAINS, c = idx % K1_CHAINS;
    betas[c * S + j] = (j < d && chain0 + c < Cp) ? bsign * a.q[(int64_t)j * Cp + chain0 + c] : 0.0;
  }
  if (PSM) {
    for (int idx = tid; idx < 3 * PS_N; idx += K1_THREADS) {
      const int row = idx / PS_N, i = idx - row * PS_N;
      ptab[idx] = probit_tab[row * PROBIT_NINT + PS_K0 + i];
    }
  }
  if (FAM == MCMCGPU_FAM_LOGISTIC) {
    if (tid < EXP_NTAB) etab[tid] = exp_tab_g[tid];
    if (tid < LOG_NINT) ltab[tid] = make_double2(log_tab_g[2 * tid], log_tab_g[2 * tid + 1]);
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < K1_STAGES && s < nt; s++) {
      mbar_expect_tx(&full[s], tile_bytes);
      bulk_g2s(tiles + (size_t)s * TILE_D, a.P.tiles + (t0 + s) * a.P.tile_doubles, tile_bytes, &full[s]);
    }
  }
  // per-warp flag: does any of this warp's chains need the log-likelihood value this wave?
  bool need_ll = true;
  if (a.need_ll) need_ll = __any_sync(0xffffffffu, mychain < Cp && a.need_ll[mychain] != 0);

  double G[DK][2];
#pragma unroll
  for (int jb = 0; jb < DK; jb++) { G[jb][0] = 0.0; G[jb][1] = 0.0; }
  double ll1 = 0.0, ll2 = 0.0;
  int nbad = 0;
  // linear family: sum of squared standardised residuals and the number of real rows this lane has seen; the constant
  // -(ln sqrt(2 pi) + log sd) per row is added once at the end
  double lin_s2 = 0.0;
  int lin_rows = 0;
  const double lin_inv_sd = (FAM == MCMCGPU_FAM_LINEAR) ? 1.0 / a.hyper[1] : 0.0;
  const double lin_inv_var = (FAM == MCMCGPU_FAM_LINEAR) ? 1.0 / (a.hyper[1] * a.hyper[1]) : 0.0;
  const double hy[4] = {a.hyper[0], a.hyper[1], a.hyper[2], a.hyper[3]};
  const int64_t N = a.P.N;

  // shared-memory offsets of this lane's fragment elements (doubles)
  //   phase 1 A fragment: beta[chain 8w+g][4ks + t]
  //   phase 1 B fragment: X[8n + pi(g)][4ks + t]
  //   phase 2 B fragment: X[8n + pi(2t+s)][8jb + g]
  const double* bfrag = betas + (warp * 8 + g) * S + t;
  const int ksn = (d + 3) >> 2;               // k-steps that hold real features (<= KS)
  const int off1 = pi8(g) * S + t;
  const int row2a = pi8(2 * t), row2b = pi8(2 * t + 1);
  const int off2a = row2a * S + g, off2b = row2b * S + g;

  for (int64_t it = 0; it < nt; it++) {
    const int slot = (int)(it % K1_STAGES);
    const uint32_t par = (uint32_t)((it / K1_STAGES) & 1);
    mbar_wait(&full[slot], par);
    const double* X = tiles + (size_t)slot * TILE_D;
    const double* ys = X + K1_ROWS * S;
    const int64_t rowbase = (t0 + it) * K1_ROWS;

#pragma unroll 1
    for (int rg = 0; rg < NG; rg++) {
      const double* Xg = X + rg * (8 * NR) * S;
      // ---- phase 1: eta for 8*NR rows x 8 chains ----
      double acc[NR][2];
#pragma unroll
      for (int n = 0; n < NR; n++) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
      // (measured and dropped: splitting the feature range into two halves for 2*NR independent accumulator chains:
      //  43.8 instead of 35.4 ms per wave)
#pragma unroll 4
      for (int ks = 0; ks < ksn; ks++) {
        const double av = bfrag[4 * ks];
#pragma unroll
        for (int n = 0; n < NR; n++) {
          double b = Xg[off1 + n * 8 * S + 4 * ks];
          dmma(acc[n][0], acc[n][1], av, b);
        }
      }
      // ---- epilogue: link function on this lane's 2*NR (row, chain) elements ----
      bool done = false;
      if (FAM == MCMCGPU_FAM_LINEAR && a.debug == 0) {
        // examples/linear_regression.jl:16-17: resid = Y - X*vars; resid ~ Normal(0, sd).  Four FP64 instructions per
        // element (every math instruction of the link adds to the DMMA time, see the probit link): the two divisions by sd
        // and sd^2 are multiplications by reciprocals formed once per thread (<= 2 ulp from the quotients; the sums of this
        // kernel are compared with the reference within a tolerance by construction), the log-likelihood is carried as
        // sum z^2 plus a row count, and padded rows (y = 0, X = 0) contribute exact zeros to both sums.
#pragma unroll
        for (int n = 0; n < NR; n++)
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const int lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
            const double resid = ys[lr] - acc[n][s];
            const double z = resid * lin_inv_sd;
            lin_s2 = fma(z, z, lin_s2);
            acc[n][s] = resid * lin_inv_var;                 // MCMCDerivRules.jl:57 through resid = Y - X*vars
          }
        if (rowbase + K1_ROWS <= N) lin_rows += 2 * NR;
        else {
#pragma unroll
          for (int n = 0; n < NR; n++)
#pragma unroll
            for (int s = 0; s < 2; s++) lin_rows += ((rowbase + rg * 8 * NR + 8 * n + (s ? row2b : row2a)) < N) ? 1 : 0;
        }
        done = true;
      }
      if (FAM == MCMCGPU_FAM_LOGISTIC && a.debug == 0) {
        // all 2*NR elements stage by stage, so their dependency chains interleave.  The same arithmetic is used whether
        // or not the log-likelihood is wanted (need_ll is a per-WARP flag: a chain's numbers must not depend on the
        // phase of its neighbours, or sharding / stepwise execution would change the draws).  acc holds x = sign * eta
        // (sign folded into beta).  15 FP64 instructions per element on a gradient wave: exp 10, 1 + e, reciprocal 3,
        // e * p; the response test, the support test and the sign of r are integer work on the bit patterns.
        // (Integer work is kept to a minimum as well -- every math instruction of the link adds to the DMMA time: one running
        // maximum of |x| serves the range test and tells whether any element can be out of support at all; the padded rows of
        // the last tile are masked on a path of their own.)
        unsigned amax = 0u;
#pragma unroll
        for (int n = 0; n < NR; n++)
#pragma unroll
          for (int s = 0; s < 2; s++) amax = max(amax, (unsigned)__double2hiint(acc[n][s]) & 0x7fffffffu);
        const bool fast = amax < 0x4085E000u;                                         // |x| < 700 and not NaN
        if (fast) {
          double ev[2 * NR], pv[2 * NR];
#pragma unroll
          for (int n = 0; n < NR; n++)
#pragma unroll
            for (int s = 0; s < 2; s++) ev[2 * n + s] = exp_tab64(acc[n][s], etab);
#pragma unroll
          for (int i = 0; i < 2 * NR; i++) pv[i] = rcp_cubic(1.0 + ev[i]);
          const int sbit = __double2hiint(hy[1]) & 0x80000000;      // sign bit of the link's sign convention
          unsigned y1mask = 0u;
          // support: log(p) is finite for |x| < 700; log(1 - p) is finite unless 1 + e rounds to 1, i.e.
          // x <= ln 2^-53 = -36.736800569677101 (padded rows have x = 0): only looked for when some |x| is that large
          if (amax >= 0x40425E4Fu) {
#pragma unroll
            for (int n = 0; n < NR; n++)
#pragma unroll
              for (int s = 0; s < 2; s++) {
                const int lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
                const bool y1 = (__double_as_longlong(ys[lr]) << 1) != 0;
                nbad += (!y1 && (unsigned long long)__double_as_longlong(acc[n][s]) >= 0xC0425E4F7B2737FAull) ? 1 : 0;
              }
          }
#pragma unroll
          for (int n = 0; n < NR; n++)
#pragma unroll
            for (int s = 0; s < 2; s++) {
              const int i = 2 * n + s, lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
              const long long yb = __double_as_longlong(ys[lr]);
              const bool y1 = (yb << 1) != 0;                        // y != 0.0 (NaN counts as a success, as != does)
              y1mask |= y1 ? (1u << i) : 0u;
              // r = y ? (-sign) (e p) : sign p  (the reference's AD chain in closed form, see link<>): magnitude, then sign bit
              const double ep = ev[i] * pv[i];
              const double mag = y1 ? ep : pv[i];
              const int flip = (y1 ? (int)0x80000000 : 0) ^ sbit;
              acc[n][s] = __hiloint2double(__double2hiint(mag) ^ flip, __double2loint(mag));
            }
          if (need_ll) {                                                        // warp-uniform; kept out of the straight-line code above
            auto loglik = [&](auto mask_tag) {
              constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll
              for (int n = 0; n < NR; n++)
#pragma unroll
                for (int s = 0; s < 2; s++) {
                  const int i = 2 * n + s, lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
                  const bool y1 = (y1mask >> i) & 1u;
                  const double arg = y1 ? pv[i] : 1.0 - pv[i];                  // Bernoulli: p1, or p0 = 1 - p1 by subtraction
                  const double lg = log_tab129(arg, ltab);
                  if (MASK) ll1 += ((rowbase + lr) < N) ? lg : 0.0;             // padded rows contribute an exact zero
                  else ll1 += lg;
                }
            };
            if (rowbase + K1_ROWS <= N) loglik(std::false_type{}); else loglik(std::true_type{});
          }
          done = true;
        }
      }
      if (FAM == MCMCGPU_FAM_PROBIT && a.debug == 0) {
        // binary responses with |eta| < 36.9 (the only case met in practice): log Phi(z) and phi(z)/Phi(z), z = +-eta, from
        // the table; log Phi is skipped when no chain of the warp needs the value.  Response tests, range tests and sign
        // flips are integer work on the bit patterns, and as little of it as possible (every math instruction of the link
        // adds to the DMMA time): one running maximum of |z| instead of two compares per element, y in {0, 1} checked once when the
        // design matrix is packed, the y == 0 flags shifted
        // into one register, ONE log-likelihood accumulator (the two dot products of probit_regression.jl:29 are added at the
        // end anyway), and the padded rows of the last tile masked on a warp-uniform path of their own.
        double zv[2 * NR];
        unsigned neg = 0u, amax = 0u;                 // neg bit (2 NR - 1 - i): y_i == 0, so z = -eta and r = -w
#pragma unroll
        for (int n = 0; n < NR; n++)
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const int i = 2 * n + s, lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
            const unsigned yh = (unsigned)__double2hiint(ys[lr]);                            // 0x3FF00000 or 0 (ybin, checked by k1_pack)
            const unsigned sg = ~(yh << 2) & 0x80000000u;                                    // y == 1: 0; y == 0: the sign bit
            neg = __funnelshift_l(sg, neg, 1);
            const unsigned zh = (unsigned)__double2hiint(acc[n][s]) ^ sg;                    // z = y ? eta : -eta
            zv[i] = __hiloint2double((int)zh, __double2loint(acc[n][s]));
            amax = max(amax, zh & 0x7fffffffu);
          }
        const bool fast = ybin && (amax < 0x40427333u);                                      // |z| < 36.9 and not NaN
        // the shared-memory copies only when every element of the warp is central, |z| < PSH (warp-uniform choice of the
        // code path; same numbers, same arithmetic: the choice never changes a result)
        const bool insm = PSM && __all_sync(0xffffffffu, fast && (amax < (PSH == 8 ? 0x40200000u : 0x40100000u)));
        if (fast) {
          auto stages = [&](auto sm_tag, auto mask_tag) {
            constexpr bool SM = decltype(sm_tag)::value, MASK = decltype(mask_tag)::value;
            constexpr int STR = SM ? PS_N : PROBIT_NINT;
            const unsigned long long* T = SM ? ptab - PS_K0 : probit_tab;
            if (a.need_grad && need_ll) {
              // value and gradient (MALA; the first wave of any run; the last leapfrog of a trajectory for the warps that hold
              // such a chain).  need_ll is a per-WARP flag and a chain's numbers must not depend on the phase of its
              // neighbours: w comes out of the same instruction sequence here and in the gradient-only branch below
              // (test_regression_chain_sharding_invariance, test_unsplit_likelihood_fuses_the_leapfrog compare bit for bit)
#pragma unroll
              for (int n = 0; n < NR; n++)
#pragma unroll
                for (int s = 0; s < 2; s++) {
                  const int i = 2 * n + s, lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
                  double l, w;
                  probit_eval<SM, STR, true, true>(T, zv[i], l, w);
                  acc[n][s] = __hiloint2double(__double2hiint(w) ^ (int)((neg << (32 - 2 * NR + i)) & 0x80000000u), __double2loint(w));   // r = y ? w : -w
                  if (MASK && (rowbase + lr) >= N) l = 0.0;
                  ll1 += l;
                }
            } else if (a.need_grad) {
#pragma unroll
              for (int n = 0; n < NR; n++)
#pragma unroll
                for (int s = 0; s < 2; s++) {
                  const int i = 2 * n + s;
                  double l, w;
                  probit_eval<SM, STR, false, true>(T, zv[i], l, w);
                  acc[n][s] = __hiloint2double(__double2hiint(w) ^ (int)((neg << (32 - 2 * NR + i)) & 0x80000000u), __double2loint(w));
                }
            } else if (need_ll) {
#pragma unroll
              for (int n = 0; n < NR; n++)
#pragma unroll
                for (int s = 0; s < 2; s++) {
                  const int i = 2 * n + s, lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);
                  double l, w;
                  probit_eval<SM, STR, true, false>(T, zv[i], l, w);
                  if (MASK && (rowbase + lr) >= N) l = 0.0;
                  ll1 += l;
                }
            }
          };
          if (rowbase + K1_ROWS <= N) {
            if (insm) stages(std::true_type{}, std::false_type{}); else stages(std::false_type{}, std::false_type{});
          } else {
            if (insm) stages(std::true_type{}, std::true_type{}); else stages(std::false_type{}, std::true_type{});
          }
          done = true;
        }
      }
      if (!done) {
#pragma unroll
        for (int n = 0; n < NR; n++) {
#pragma unroll
          for (int s = 0; s < 2; s++) {
            const int lr = rg * 8 * NR + 8 * n + (s ? row2b : row2a);   // local row of accumulator column 2t+s
            const double y = ys[lr];
            LinkOut o;
            if (a.debug == 1) { o.ll1 = 0.0; o.ll2 = 0.0; o.bad = false; o.r = acc[n][s] * y; }
            else if (FAM == MCMCGPU_FAM_LOGISTIC) {
              // acc = sign * eta (sign folded into beta): the link with sign 1, then d/d eta = sign * d/d x
              const double h1[4] = {hy[0], 1.0, hy[2], hy[3]};
              o = link<FAM>(acc[n][s], y, h1, need_ll);
              o.r *= hy[1];
            } else o = link<FAM>(acc[n][s], y, hy, need_ll);
            const bool valid = (rowbase + lr) < N;
            if (valid) { ll1 += o.ll1; ll2 += o.ll2; nbad += o.bad ? 1 : 0; }
            acc[n][s] = o.r;    // rows >= N have X == 0, so their r never reaches G
          }
        }
      }
      // ---- phase 2: G += r^T X ----
      if (a.need_grad && a.debug != 2) {
        // consecutive DMMAs go to different accumulators (DK independent chains)
#pragma unroll
        for (int n = 0; n < NR; n++) {
#pragma unroll
          for (int jb = 0; jb < DK; jb++) {
            double b0 = Xg[off2a + n * 8 * S + 8 * jb];
            dmma(G[jb][0], G[jb][1], acc[n][0], b0);
          }
#pragma unroll
          for (int jb = 0; jb < DK; jb++) {
            double b1 = Xg[off2b + n * 8 * S + 8 * jb];
            dmma(G[jb][0], G[jb][1], acc[n][1], b1);
          }
        }
      }
    }
    // release the slot: the LAST warp to finish with it issues the refill (tile it + STAGES), so no warp ever waits
    // for the slowest one just to start a copy
    __syncwarp();
    if (lane == 0) {
      const unsigned int prev = atom_add_acq_rel_cta(&released[slot], 1u);
      if (prev == K1_WARPS - 1) {
        // the counter reset is published by the release of mbarrier.arrive.expect_tx below (the next arrivals on this slot
        // come after an acquire-wait on that barrier); with no refill left the slot is never counted again
        released[slot] = 0;
        if (it + K1_STAGES < nt) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads (acquired above) before the async-proxy overwrite
          mbar_expect_tx(&full[slot], tile_bytes);
          bulk_g2s(tiles + (size_t)slot * TILE_D, a.P.tiles + (t0 + it + K1_STAGES) * a.P.tile_doubles, tile_bytes,
                   &full[slot]);
        }
      }
    }
  }

  if (FAM == MCMCGPU_FAM_LINEAR && a.debug == 0) {
    ll1 = -fma(0.5, lin_s2, (double)lin_rows * (MG_LN_SQRT_2PI + a.hyper[3]));     // hyper[3] = log(noise_sd)
    nbad = isfinite(lin_s2) ? 0 : 1;                   // a non-finite term anywhere makes the sum of squares non-finite
  }
  // ---- write this split's partial sums ----
  // quad reduction of the scalar sums (the 4 lanes of a quad hold different rows of the same chain)
  ll1 += __shfl_xor_sync(0xffffffffu, ll1, 1); ll1 += __shfl_xor_sync(0xffffffffu, ll1, 2);
  ll2 += __shfl_xor_sync(0xffffffffu, ll2, 1); ll2 += __shfl_xor_sync(0xffffffffu, ll2, 2);
  nbad += __shfl_xor_sync(0xffffffffu, nbad, 1); nbad += __shfl_xor_sync(0xffffffffu, nbad, 2);
  double* part = a.part + (int64_t)split * (d + 2) * Cp;
  if (mychain >= Cp) return;          // overhang of the last (wide) chain tile
  if (t == 0) {
    part[(int64_t)d * Cp + mychain] = ll1 + ll2;
    part[(int64_t)(d + 1) * Cp + mychain] = (double)nbad;
  }
  if (a.need_grad) {
    // fused interior leapfrog (K1Args::fuse_leap; one split, so G is the chain's complete X'r).  Written with explicit
    // round-to-nearest intrinsics: this translation unit allows multiply-add contraction and the update must round as
    // the transition kernel's does (HMC.jl:95-98: (0.5 g) eps, m += ., p += eps m).
    bool fused = false;
    int lp = 0, nlc = 0;
    if (a.fuse_leap) { lp = a.leap[mychain]; nlc = a.nleaps_cur[mychain]; fused = (a.phase[mychain] == PH_LEAP_K1) && (lp + 2 <= nlc); }
    if (a.fuse_leap) {     // the chain's counters too: on a wave of interior leapfrogs the transition kernel need not run at all
      const unsigned done = __ballot_sync(0xffffffffu, fused && t == 0);
      if (t == 0) a.k1_done[mychain] = fused ? 1 : 0;      // written for every evaluated chain, every wave: never stale
      if (fused && t == 0) {
        a.leap[mychain] = lp + 1;
        a.need_ll_rw[mychain] = (lp + 2 == nlc) ? 1 : 0;
      }
      if (lane == 0 && done) atomicAdd(a.n_evals, (unsigned long long)__popc(done));
    }
    if (fused) {
      const double eps = a.eps_cur[mychain];
      const bool linlog = (FAM == MCMCGPU_FAM_LINEAR || FAM == MCMCGPU_FAM_LOGISTIC);
      const bool oos = linlog && nbad > 0;                 // LLAcc: out of support => zero gradient (modelparser.jl:64-72)
      const double pvar = __dmul_rn(hy[0], hy[0]);
      const bool vpow2 = k1_var_is_pow2(pvar);
      const double pinv = vpow2 ? __ddiv_rn(1.0, pvar) : 0.0;
      const double* brow = betas + (warp * 8 + g) * S;
      // all momentum loads first: the stores below go to the same array as far as the compiler knows, so a load placed after
      // a store waits for it (one global round trip per element, ~16 us per CTA, when the loop is written element by element)
      double mo[DK][2];
#pragma unroll
      for (int jb = 0; jb < DK; jb++) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int j = 8 * jb + 2 * t + i;
          mo[jb][i] = (j < d) ? __ldcg(a.mom + (int64_t)j * Cp + mychain) : 0.0;
        }
      }
#pragma unroll
      for (int jb = 0; jb < DK; jb++) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const int j = 8 * jb + 2 * t + i;
          if (j < d) {
            const int64_t idx = (int64_t)j * Cp + mychain;
            const double qj = __dmul_rn(bsign, brow[j]);   // beta tile holds sign * q, sign = +-1: exact
            double gj;
            if (linlog) gj = oos ? 0.0 : __dadd_rn(G[jb][i], k1_div_by_var(__dsub_rn(0.0, qj), pvar, pinv, vpow2));
            else gj = __dsub_rn(G[jb][i], k1_div_by_var(qj, pvar, pinv, vpow2));
            const double hg = __dmul_rn(__dmul_rn(0.5, gj), eps);
            double m = mo[jb][i];
            m = __dadd_rn(m, hg);                          // end of this leapfrog      HMC.jl:98
            m = __dadd_rn(m, hg);                          // start of the next one     HMC.jl:95
            const double p = __dadd_rn(qj, __dmul_rn(eps, m));   //                     HMC.jl:96
            __stcg(a.mom + idx, m);
            __stcg(a.q_rw + idx, p);
          }
        }
      }
    } else {
#pragma unroll
      for (int jb = 0; jb < DK; jb++) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          int j = 8 * jb + 2 * t + i;   // C fragment: row g (chain), column 2t+i (feature)
          if (j < d) part[(int64_t)j * Cp + mychain] = G[jb][i];
        }
      }
    }
  }
}

// ---- packing kernel: column-major X -> tile images -------------------------------------------
__global__ void k1_pack_kernel(double* tiles, const double* X, const double* y, int64_t N, int64_t d, int S,
                               int64_t tile_doubles, int64_t ntiles, unsigned int* ynb) {
  // one block per tile; threads stride over (row, col)
  const int64_t tl = blockIdx.x;
  double* T = tiles + tl * tile_doubles;
  const int64_t row0 = tl * K1_ROWS;
  for (int idx = threadIdx.x; idx < K1_ROWS * S; idx += blockDim.x) {
    int j = idx / K1_ROWS, r = idx % K1_ROWS;    // r fastest: coalesced reads along a column of X
    int64_t row = row0 + r;
    double v = (j < d && row < N) ? X[(int64_t)j * N + row] : 0.0;
    T[r * S + j] = v;
  }
  for (int r = threadIdx.x; r < K1_ROWS; r += blockDim.x) {
    int64_t row = row0 + r;
    const double yv = (row < N) ? y[row] : 0.0;
    T[K1_ROWS * S + r] = yv;
    const long long yb = __double_as_longlong(yv);
    if (yb != 0ll && yb != 0x3FF0000000000000ll) atomicOr(ynb, 1u);     // the probit link tests this word once instead of every y
  }
}

// supported feature-block counts: every DK up to 13 (d <= 104, 128 registers, two CTAs per SM), then 14, 15, 16, 20, 25
// (d <= 200, one CTA per SM with up to 255 registers for the 2*DK gradient accumulators per thread: two CTAs of DK >= 14
// would need 2 x 119 KB of shared memory).  14 and 15 exist so that 104 < d <= 120 does not execute the dead phase-2 DMMAs
// of DK = 16 (d = 105: 14 % of them).
static int k1_round_dk(int64_t d) {
  int dk = (int)((d + 7) / 8);
  if (dk <= K1_MAX_DK_2CTA) return dk;
  if (dk <= 16) return dk;
  if (dk <= 20) return 20;
  if (dk <= 25) return 25;
  return -1;
}
bool k1_supported(int64_t d) { return d >= 1 && k1_round_dk(d) > 0; }

cudaError_t k1_pack(K1Pack& P, const double* dX, const double* dy, int64_t N, int64_t d, cudaStream_t st) {
  P.N = N; P.d = d;
  P.DK = k1_round_dk(d);
  P.S = 8 * P.DK + 4;
  P.ntiles = (N + K1_ROWS - 1) / K1_ROWS;
  P.tile_doubles = (int64_t)K1_ROWS * P.S + K1_ROWS;
  cudaError_t e = cudaMalloc(&P.tiles, sizeof(double) * (size_t)(P.ntiles * P.tile_doubles + 2));
  if (e != cudaSuccess) return e;
  unsigned int* ynb = reinterpret_cast<unsigned int*>(P.tiles + P.ntiles * P.tile_doubles);
  P.ynb = ynb;
  e = cudaMemsetAsync(ynb, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return e;
  k1_pack_kernel<<<(unsigned)P.ntiles, 256, 0, st>>>(P.tiles, dX, dy, N, d, P.S, P.tile_doubles, P.ntiles, ynb);
  return cudaGetLastError();
}
void k1_free(K1Pack& P) { if (P.tiles) cudaFree(P.tiles); P.tiles = nullptr; }

int k1_choose_splits(const K1Pack& P, int64_t Cp, int device, int family) {
  // Row splits: enough CTAs to fill the machine, and a CTA count whose last wave is nearly full
  // (two resident CTAs per SM).  Every split keeps >= 8 tiles so the prologue stays amortised.
  const int64_t ctiles = Cp / K1_CHAINS;
  int sms = 148;
  if (device < 0) cudaGetDevice(&device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int64_t small_ctas = (family == MCMCGPU_FAM_PROBIT) ? K1_PROBIT_SMALL_CTAS : K1_OTHER_SMALL_CTAS;
  const int64_t slots = (P.DK <= K1_MAX_DK_3CTA ? small_ctas : (P.DK <= K1_MAX_DK_2CTA ? 2LL : 1LL)) * sms;
  int64_t maxs = P.ntiles / 8;
  if (maxs < 1) maxs = 1;
  if (maxs > 512) maxs = 512;
  int best = 1; double beste = -1.0;
  for (int64_t s = 1; s <= maxs; s++) {
    const double waves = (double)(ctiles * s) / (double)slots;
    double eff = waves / std::ceil(waves);
    if (waves < 1.0) eff = waves;              // not enough CTAs to fill the machine once
    if (eff > beste + 0.02) { beste = eff; best = (int)s; }   // prefer fewer splits unless clearly better
  }
  // small problems: when even the best choice leaves SMs idle, latency matters more than amortising the prologue --
  // spread the rows over as many CTAs as there are free slots (down to one tile per CTA)
  if (ctiles * best < slots) {
    int64_t s = slots / ctiles;
    if (s > P.ntiles) s = P.ntiles;
    if (s > 512) s = 512;
    if (s > best) best = (int)s;
  }
  return best;
}

template <int FAM, int DK, int WARPS, int STAGES, int CTAS>
static cudaError_t launch_shape(const K1Args& a, cudaStream_t st) {
  constexpr int NR = K1_NR;
  constexpr size_t smem = k1_smem_base(DK, WARPS, STAGES) +
                          ((FAM == MCMCGPU_FAM_PROBIT && k1_probit_half(DK, WARPS, STAGES, CTAS) > 0) ? ps_bytes(k1_probit_half(DK, WARPS, STAGES, CTAS)) : 0);
  static_assert(smem <= 232448, "one CTA must fit the 227 KB of dynamic shared memory");
  static bool attr_done[64] = {false};      // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k1_kernel<FAM, DK, NR, WARPS, STAGES, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  dim3 grid((unsigned)((a.Cp + 8 * WARPS - 1) / (8 * WARPS)), (unsigned)a.nsplit);
  k1_kernel<FAM, DK, NR, WARPS, STAGES, CTAS><<<grid, 32 * WARPS, smem, st>>>(a);
  return cudaGetLastError();
}

template <int FAM, int DK>
static cudaError_t launch_fd(const K1Args& a, cudaStream_t st) {
  // 8 warps x 64 chains, 2-deep ring; 3 / 2 / 1 resident CTAs by register and shared-memory budget.  (Measured and dropped:
  // one 16-warp x 128-chain CTA per SM with a 4-deep ring for 32 < d <= 104 -- same warps per SM, half the X-tile traffic,
  // three tiles of slack between the fastest and the slowest warp -- 35.0 vs 35.1 ms per wave: the ring is not the limiter.)
  constexpr bool SMALL = (DK <= K1_MAX_DK_3CTA);
  constexpr bool PSMALL = SMALL && (FAM == MCMCGPU_FAM_PROBIT);
  constexpr int CTAS = PSMALL ? K1_PROBIT_SMALL_CTAS : (SMALL ? K1_OTHER_SMALL_CTAS : ((DK <= K1_MAX_DK_2CTA) ? 2 : 1));
  return launch_shape<FAM, DK, K1_WARPS, PSMALL ? K1_PROBIT_SMALL_STAGES : (SMALL ? K1_OTHER_SMALL_STAGES : K1_STAGES), CTAS>(a, st);
}

template <int FAM>
static cudaError_t launch_f(const K1Args& a, cudaStream_t st) {
  switch (a.P.DK) {
    case 3: return launch_fd<FAM, 3>(a, st);
#ifndef K1_ONLY_DK3        // experiments: -DK1_ONLY_DK3 builds the d <= 24 kernels only (seconds instead of minutes per variant)
    case 1: return launch_fd<FAM, 1>(a, st);
    case 2: return launch_fd<FAM, 2>(a, st);
    case 4: return launch_fd<FAM, 4>(a, st);
    case 5: return launch_fd<FAM, 5>(a, st);
    case 6: return launch_fd<FAM, 6>(a, st);
    case 7: return launch_fd<FAM, 7>(a, st);
    case 8: return launch_fd<FAM, 8>(a, st);
    case 9: return launch_fd<FAM, 9>(a, st);
    case 10: return launch_fd<FAM, 10>(a, st);
    case 11: return launch_fd<FAM, 11>(a, st);
    case 12: return launch_fd<FAM, 12>(a, st);
    case 13: return launch_fd<FAM, 13>(a, st);
    case 14: return launch_fd<FAM, 14>(a, st);
    case 15: return launch_fd<FAM, 15>(a, st);
    case 16: return launch_fd<FAM, 16>(a, st);
    case 20: return launch_fd<FAM, 20>(a, st);
    case 25: return launch_fd<FAM, 25>(a, st);
#endif
  }
  return cudaErrorInvalidValue;
}

cudaError_t k1_launch(const K1Args& a, cudaStream_t st) {
  switch (a.family) {
    case MCMCGPU_FAM_LINEAR: return launch_f<MCMCGPU_FAM_LINEAR>(a, st);
    case MCMCGPU_FAM_LOGISTIC: return launch_f<MCMCGPU_FAM_LOGISTIC>(a, st);
    case MCMCGPU_FAM_PROBIT: return launch_f<MCMCGPU_FAM_PROBIT>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace mg
