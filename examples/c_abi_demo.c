/* c_abi_demo.c -- the C ABI of libmcmcgpu.so used from plain C, the way a Julia `ccall` (or any FFI) would:
 * logistic regression (examples/logistic_regression.jl shape, synthetic data), HMC(2, 0.1), 8 chains, 1000 steps.
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -Lmcmc.jl_b200 -lmcmcgpu -Wl,-rpath,$PWD/mcmc.jl_b200 -lm -o c_abi_demo
 * With the argument "layout" it only prints the struct layouts (no GPU needed); tests compare them with the bindings. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mcmcgpu.h"

static double urand(unsigned long long* s) { *s = *s * 6364136223846793005ULL + 1442695040888963407ULL; return ((*s >> 11) + 0.5) / 9007199254740992.0; }
static double nrand(unsigned long long* s) { return sqrt(-2.0 * log(urand(s))) * cos(6.283185307179586 * urand(s)); }

#define CHECK(x) do { int rc__ = (x); if (rc__ != MCMCGPU_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc__, mcmcgpu_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "layout") == 0) {
    printf("sampler_cfg %zu kind %zu nleaps %zu scale %zu rate %zu len %zu shrinkage %zu t0 %zu step %zu max_leaps %zu tuner_on %zu adapt_step %zu max_step %zu target_path %zu target_rate %zu\n",
           sizeof(mcmcgpu_sampler_cfg), offsetof(mcmcgpu_sampler_cfg, kind), offsetof(mcmcgpu_sampler_cfg, nleaps), offsetof(mcmcgpu_sampler_cfg, scale),
           offsetof(mcmcgpu_sampler_cfg, rate), offsetof(mcmcgpu_sampler_cfg, len), offsetof(mcmcgpu_sampler_cfg, shrinkage), offsetof(mcmcgpu_sampler_cfg, t0),
           offsetof(mcmcgpu_sampler_cfg, step), offsetof(mcmcgpu_sampler_cfg, max_leaps), offsetof(mcmcgpu_sampler_cfg, tuner_on),
           offsetof(mcmcgpu_sampler_cfg, adapt_step), offsetof(mcmcgpu_sampler_cfg, max_step), offsetof(mcmcgpu_sampler_cfg, target_path),
           offsetof(mcmcgpu_sampler_cfg, target_rate));
    printf("runner_cfg %zu first %zu step %zu last %zu nchains %zu chain_offset %zu seed %zu init_per_chain %zu store_grad %zu store_logtarget %zu engine %zu store_rb %zu stream_stats %zu stream_batchlen %zu\n",
           sizeof(mcmcgpu_runner_cfg), offsetof(mcmcgpu_runner_cfg, first), offsetof(mcmcgpu_runner_cfg, step), offsetof(mcmcgpu_runner_cfg, last),
           offsetof(mcmcgpu_runner_cfg, nchains), offsetof(mcmcgpu_runner_cfg, chain_offset), offsetof(mcmcgpu_runner_cfg, seed),
           offsetof(mcmcgpu_runner_cfg, init_per_chain), offsetof(mcmcgpu_runner_cfg, store_grad), offsetof(mcmcgpu_runner_cfg, store_logtarget),
           offsetof(mcmcgpu_runner_cfg, engine), offsetof(mcmcgpu_runner_cfg, store_rb), offsetof(mcmcgpu_runner_cfg, stream_stats),
           offsetof(mcmcgpu_runner_cfg, stream_batchlen));
    printf("run_info %zu gpu_ms %zu n_grad_evals %zu n_waves %zu n_launches %zu eval_ms %zu comm_ms %zu\n", sizeof(mcmcgpu_run_info),
           offsetof(mcmcgpu_run_info, gpu_ms), offsetof(mcmcgpu_run_info, n_grad_evals), offsetof(mcmcgpu_run_info, n_waves),
           offsetof(mcmcgpu_run_info, n_launches), offsetof(mcmcgpu_run_info, eval_ms), offsetof(mcmcgpu_run_info, comm_ms));
    return 0;
  }
  const int64_t N = 1000, d = 10, C = 8, first = 101, last = 1000, S = last - first + 1;
  unsigned long long seed = 1;
  double* X = malloc(sizeof(double) * N * d);     /* column-major N x d */
  double* y = malloc(sizeof(double) * N);
  double beta0[10];
  for (int j = 0; j < d; j++) beta0[j] = nrand(&seed);
  for (int64_t i = 0; i < N; i++) X[i] = 1.0;
  for (int64_t j = 1; j < d; j++) for (int64_t i = 0; i < N; i++) X[j * N + i] = nrand(&seed);
  for (int64_t i = 0; i < N; i++) {
    double eta = 0.0;
    for (int64_t j = 0; j < d; j++) eta += X[j * N + i] * beta0[j];
    y[i] = urand(&seed) < 1.0 / (1.0 + exp(-eta)) ? 1.0 : 0.0;
  }
  mcmcgpu_ctx* ctx; mcmcgpu_model* m;
  CHECK(mcmcgpu_init(-1, &ctx));
  double hyper[2] = {1.0, -1.0};
  CHECK(mcmcgpu_model_create(ctx, MCMCGPU_FAM_LOGISTIC, N, d, X, y, hyper, 2, 0, &m));
  mcmcgpu_sampler_cfg s; memset(&s, 0, sizeof(s));
  s.kind = MCMCGPU_HMC; s.nleaps = 2; s.scale = 0.1;                       /* HMC(2, 0.1), logistic_regression.jl:32 */
  mcmcgpu_runner_cfg r; memset(&r, 0, sizeof(r));
  r.first = first; r.step = 1; r.last = last; r.nchains = C; r.seed = 42;   /* SerialMC(101:1000) x 8 chains */
  double init[10] = {0};
  double* samples = malloc(sizeof(double) * d * S * C);
  double* grads = malloc(sizeof(double) * d * S * C);
  uint8_t* accept = malloc((size_t)(S * C));
  double* lt = malloc(sizeof(double) * S * C);
  mcmcgpu_run_info info;
  CHECK(mcmcgpu_run_chains(m, &s, &r, init, NULL, NULL, NULL, samples, grads, accept, lt, &info));
  double acc = 0.0, err = 0.0;
  for (int64_t k = 0; k < S * C; k++) acc += accept[k];
  for (int64_t j = 0; j < d; j++) {
    double mean = 0.0;
    for (int64_t c = 0; c < C; c++) for (int64_t t = 0; t < S; t++) mean += samples[(c * S + t) * d + j];
    mean /= (double)(S * C);
    err = fmax(err, fabs(mean - beta0[j]));
  }
  double ess[80], var[80];
  CHECK(mcmcgpu_stats(ctx, samples, S, d, C, MCMCGPU_VAR_IMSE, -1, 0, NULL, NULL, var, ess, NULL));
  printf("C_ABI_DEMO ok: acceptance %.1f %%, max |posterior mean - beta0| %.3f, ESS[0] %.1f, gpu %.1f ms, %lld gradient evaluations, %lld launches\n",
         100.0 * acc / (double)(S * C), err, ess[0], info.gpu_ms, (long long)info.n_grad_evals, (long long)info.n_launches);
  mcmcgpu_model_destroy(m);
  mcmcgpu_destroy(ctx);
  return (acc > 0 && err < 1.0) ? 0 : 2;
}
