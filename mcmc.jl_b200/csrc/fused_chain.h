#pragma once
#include "common.cuh"
namespace mg {
constexpr int FUSED_THREADS = 64;   // 65 536 chains = 1024 CTAs = 6.9 per SM: 64-thread CTAs balance the SMs to 1 % (128: 16 vs 12 warps per SM)
constexpr int FUSED_MAX_D = 8;
constexpr int FUSED_STREAM_ROWS = 5;   // mean, var_iid, var_bm, ess(bm), actime(bm)
struct FusedArgs {
  ModelDev M;
  SamplerDev S;
  RunnerDev R;
  const double* init;          // [d] or [d][Cp]
  const double* scale;         // [d]
  const double* inj_normals;   // [(last+1)][d][Cp] or null
  const double* inj_uniforms;  // [(last+1)][Cp] or null
  double* samples;             // [S][d][Cp]
  double* grads;               // [S][d][Cp] or null
  uint8_t* accept;             // [S][Cp]
  double* logtarget;           // [S][Cp] or null
  double* rb;                  // [S][d][Cp] or null: Rao-Blackwell sums (mean.jl:11-35)
  double* eps;                 // [S][Cp] or null
  int32_t* nleaps;             // [S][Cp] or null
  double* final_eps;           // [Cp] or null
  double* final_pars;          // [d][Cp] or null
  double* stream;              // [FUSED_STREAM_ROWS][d][Cp] streamed summaries (mcmcgpu_runner_cfg::stream_stats) or null;
                               // then samples / grads / logtarget may be null and no draw is stored
  double* stream_accept;       // [Cp] acceptance in percent (summary.jl:13), with `stream`
  int64_t stream_batchlen;
  int32_t* status;             // [Cp]
  unsigned long long* n_evals; // scalar
};
bool fused_supported(int family, int64_t d, int64_t N);
cudaError_t launch_fused(const FusedArgs& A, cudaStream_t st);
}  // namespace mg
