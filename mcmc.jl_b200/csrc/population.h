#pragma once
#include "common.cuh"
namespace mg {
constexpr int POP_MAX_TASKS = 64;
struct PopTasks {            // per-task model hyper-parameters and sampler settings (kernel argument, by value)
  int32_t nt, family, d;
  double hyper[POP_MAX_TASKS][2];
  int32_t kind[POP_MAX_TASKS], nleaps[POP_MAX_TASKS];
  double scale[POP_MAX_TASKS];
};
struct SeqArgs {
  PopTasks T;
  int64_t npart, Np;         // particles held by THIS rank, pitch
  int64_t gpart;             // particles over all ranks (== npart on one GPU)
  int32_t rank, nranks;
  double* sendbuf;           // [d+2][Np]: pars rows, logtarget, logW of this rank (multi-GPU only)
  const double* gathered;    // [nranks][d+2][Np] after ncclAllGather (multi-GPU only)
  int64_t iter, target;      // current iteration (1-based) and target (0-based)
  int64_t steps, burnin;
  uint64_t seed;
  double trigger;
  double *pars, *pars_tmp;   // [d][Np]
  double *logW, *logtarget, *lt_tmp, *W, *cp;   // [Np]
  const double *inj_normals, *inj_uniforms, *inj_res;   // host layouts, or null
  double* samples; double* weights;   // host layouts: d x S*npart, S*npart
  unsigned long long* nres;
  unsigned long long* nevals;
};
cudaError_t launch_seqmc_mutate(const SeqArgs& A, cudaStream_t st);
cudaError_t launch_seqmc_resample(const SeqArgs& A, cudaStream_t st);
cudaError_t launch_seqmc_store(const SeqArgs& A, cudaStream_t st);
cudaError_t launch_seqmc_pack(const SeqArgs& A, cudaStream_t st);
// ppars [d][Np], plt [Np], ll0 [Np]: the one-step result of the wave engine for the current (iteration, target)
cudaError_t launch_seqmc_apply(const SeqArgs& A, const double* ppars, const double* plt, const double* ll0, cudaStream_t st);

struct TempArgs {
  PopTasks T;
  int64_t nrep, steps, burnin, swap_period;
  uint64_t seed;
  const double* inits;       // d x nt (host layout)
  const double *inj_normals, *inj_uniforms, *inj_pick, *inj_swap;
  double* samples;           // d x S x nrep (host layout)
  int32_t* at;               // S x nrep or null
  int32_t* status;           // [nrep]
  unsigned long long* nevals;
};
cudaError_t launch_serialtemp(const TempArgs& A, cudaStream_t st);

// SerialTempMC over arbitrary models: per-replica state on the device, replica-minor ([d][Rp])
struct TempMArgs {
  int32_t nt, d;
  int64_t nrep, Rp, rep_offset;      // replicas of this context, pitch, global id of the first one (Philox key)
  int64_t steps, burnin;
  uint64_t seed;
  double *state, *pars, *ppars;      // [d][Rp]: sampler state, s.pars, s.ppars (SerialTempMC.jl:51-71)
  double *logtarget;                 // [Rp] s.logtarget
  int32_t *at, *sel;                 // [Rp] current task; task consumed at this step
  double *res_pp, *res_lt0;          // [d][Rp], [Rp]: the one-step results scattered back by replica
  const double *inj_pick, *inj_swap; // host layouts (nrep x (steps+1)) on the device, or null
  double* samples;                   // d x S x nrep (host layout)
  int32_t* at_out;                   // S x nrep or null
};
cudaError_t launch_temp_plan(const TempMArgs& A, int64_t i, bool swap_step, cudaStream_t st);
cudaError_t launch_temp_gather(const TempMArgs& A, const int64_t* idx, int64_t n, int64_t Cp, double* start, int64_t* chain_ids, cudaStream_t st);
cudaError_t launch_temp_scatter(const TempMArgs& A, const int64_t* idx, int64_t n, int64_t Cp, const double* ppars, const double* lt0, cudaStream_t st);
cudaError_t launch_temp_update(const TempMArgs& A, int64_t i, bool swap_step, cudaStream_t st);
bool pop_supported(int family, int64_t d);
}  // namespace mg
