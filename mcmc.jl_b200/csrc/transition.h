#pragma once
#include "common.cuh"
namespace mg {

// per-chain phases of the wave engine
constexpr int PH_INIT = 0;   // q = init has been evaluated
constexpr int PH_RWM = 1;    // q = RWM proposal has been evaluated
constexpr int PH_MALA = 2;   // q = MALA proposal has been evaluated
constexpr int PH_LEAP = 3;   // q = position after a leapfrog position update has been evaluated
constexpr int PH_RAM_BEGIN = 4;  // RAM: the next proposal is made by ram_kernel (after the factor update)
constexpr int PH_PAUSE = 98;  // reached the step limit of this execute call; resumes at the next one
constexpr int PH_DONE = 99;

struct WaveArgs {
  ModelDev M;
  SamplerDev S;
  RunnerDev R;
  int32_t nsplit;              // partial buffers to sum (1 after an all-reduce)
  int32_t resume;              // 1: this launch only restarts paused chains (no evaluation is consumed)
  int32_t restore_da;          // 1: HMCDA adaptation state was set by the host (mcmcgpu_run_set_state)
  int64_t step0;               // chains start at step step0 + 1
  int64_t step_limit;          // chains pause before starting a step > step_limit
  int32_t fused_interior;      // 1: the likelihood kernel has already made the interior leapfrog updates (K1Args::fuse_leap)
  uint8_t* k1_done;            // [Cp] with fused_interior: 1 for the chains the likelihood kernel has fully advanced this wave
                               // (it writes the flag of every chain it evaluates, every wave)
  // evaluation in/out
  double* q;                   // [d][Cp]
  const double* part;          // [nsplit][d+2][Cp]
  // chain state
  double *cur_pars, *cur_grad, *cur_lt, *mom, *H0;
  int32_t *phase, *leap, *nleaps_cur;
  int64_t *istep, *kept;
  double *eps_cur;
  double *da_leapstep, *da_dual, *da_dualH;
  double *tn_step; int64_t *tn_nleaps, *tn_acc, *tn_prop;
  double* ram_S;               // RAM: [d*d][Cp] lower-triangular proposal factor (row-major index a*d+b), d <= RAM_WAVE_MAX_D
  double* ram_Sb;              // RAM, d > RAM_WAVE_MAX_D: [C][d*d] chain-major factor (one CTA per chain), else null
  double* ram_scratch;         // RAM, d > RAM_WAVE_MAX_D: [C][3][d*d] work matrices
  double* ram_al;              // RAM: min(1, exp(ratio)) of the step just decided
  uint8_t* ram_pending;        // RAM: factor update pending for that step
  uint8_t* need_ll;
  int32_t* status;
  int32_t* remaining;
  unsigned long long* n_evals;
  // inputs
  const double* init; const double* scale; const double* inj_normals; const double* inj_uniforms;
  // outputs
  double* samples; double* grads; uint8_t* accept; double* logtarget; double* eps; int32_t* nleaps;
  double* final_eps;
  double* rb;                  // [S][d][Cp] Rao-Blackwell sums (store_rb) or null
  double* rb_acc;              // [d][Cp] running sum_k w_k pars_k of the current trajectory
  double* init_lt;             // [Cp] log-target at the initial point, or null
  const int64_t* chain_ids;    // [Cp] global chain id of every chain (Philox key), or null = chain_offset + position
};

cudaError_t launch_transition(const WaveArgs& W, cudaStream_t st);
constexpr int RAM_WAVE_MAX_D = 16;    // one thread per chain, factor in local memory; larger d: one CTA per chain (ram_big_kernel)
constexpr int RAM_BIG_MAX_D = 128;
cudaError_t launch_ram(const WaveArgs& W, bool init, cudaStream_t st);   // RAM.jl:50-60,73-78 for the wave engine
// closed-form families through the wave engine: writes part[0][(d+2)][Cp] directly (final lt / grad)
cudaError_t launch_eval_closed(const ModelDev& M, const double* q, double* part, int64_t C, int64_t Cp,
                               const int32_t* phase, const int32_t* remaining, cudaStream_t st);
// finalize an evaluation outside a run (mcmcgpu_logtarget_grad): part -> lt[Cp], grad[d][Cp]
cudaError_t launch_finalize(const ModelDev& M, const double* q, const double* part, int nsplit, int64_t C, int64_t Cp,
                            double* lt, double* grad, cudaStream_t st);
cudaError_t launch_reduce_splits(const double* part, int nsplit, int64_t rows, int64_t Cp, double* red, cudaStream_t st);
// layout changes between the host's chain-major arrays and the device's chain-minor arrays
cudaError_t transpose_to_chain_minor(const double* in /*[C][K]*/, double* out /*[K][Cp]*/, int64_t C, int64_t K, int64_t Cp, cudaStream_t st);
cudaError_t transpose_to_chain_major(const double* in /*[K][Cp]*/, double* out /*[nc][K]*/, int64_t c0, int64_t nc, int64_t K, int64_t Cp, cudaStream_t st);
cudaError_t transpose_to_chain_major_u8(const uint8_t* in, uint8_t* out, int64_t c0, int64_t nc, int64_t K, int64_t Cp, cudaStream_t st);
cudaError_t launch_philox_dump(uint64_t seed, int64_t chain_offset, int64_t C, int64_t d, int64_t last,
                               double* normals /*[C][last+1][d]*/, double* uniforms /*[C][last+1]*/, cudaStream_t st);
cudaError_t launch_fill(double* p, double v, int64_t n, cudaStream_t st);

}  // namespace mg
