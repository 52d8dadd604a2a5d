#pragma once
#include "common.cuh"
namespace mg {
// K3: src/stats over stored draws, one thread per (chain, parameter) series.
// samples: [S][d][Cp] chain-minor.  Outputs [d][Cp] (any may be null).
// `mean` is required (the second pass reads it); scratch: d * Cp doubles for MCMCGPU_VAR_BM, STATS_SCRATCH_PLANES * d * Cp + 2
// for IMSE / IPSE.  For IMSE / IPSE *n_unfinished (host) receives the number of series whose Geyer scan needs more than the
// first window (the stream is synchronised); launch_stats_more finishes them: with `gather` (n_unfinished * S doubles) through
// a compact time-contiguous copy, one warp per series; with gather == nullptr one thread per series on the strided draws.
constexpr int STATS_SCRATCH_PLANES = 7;
cudaError_t launch_stats(const double* samples, int64_t S, int64_t d, int64_t C, int64_t Cp, int vtype, int64_t maxlag,
                         int64_t batchlen, double* mean, double* var_iid, double* var, double* ess, double* actime,
                         double* scratch, unsigned int* n_unfinished, cudaStream_t st);
cudaError_t launch_stats_more(const double* samples, int64_t S, int64_t d, int64_t C, int64_t Cp, int vtype, int64_t maxlag,
                              const double* mean, double* var_iid, double* var, double* ess, double* actime, double* scratch,
                              unsigned int cnt, double* gather, cudaStream_t st);
// acceptance(c) in percent per chain (summary.jl:6-15): accept [S][Cp] -> rate [Cp]
cudaError_t launch_accept_rate(const uint8_t* accept, int64_t S, int64_t C, int64_t Cp, double* rate, cudaStream_t st);
}  // namespace mg
