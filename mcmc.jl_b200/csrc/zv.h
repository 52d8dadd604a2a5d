#pragma once
#include "common.cuh"
namespace mg {
// zero-variance control variates (src/stats/zv.jl:8-66): one CTA per chain.
// samples / grads: [S][d][Cp] chain-minor.  zv_out: [S][d][Cp] or null; a_out: [k*d][Cp] (index p*d+i).
// status[c] = 1 when the feature covariance is singular.
cudaError_t launch_zv(const double* samples, const double* grads, int64_t S, int64_t d, int64_t C, int64_t Cp, int order,
                      double* zv_out, double* a_out, int32_t* status, cudaStream_t st);
int64_t zv_features(int64_t d, int order);
bool zv_supported(int64_t d, int order);
}  // namespace mg
