// nccl_dyn.h -- NCCL entered through dlopen so that libmcmcgpu.so has no link-time dependency on it
// (inside a PyTorch process this resolves to the NCCL torch already loaded; in a Julia process to the
// system libnccl.so.2).  Only the handful of entry points the row-sharded path needs.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace mg {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t /*sendcount*/, int /*dtype*/, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_FLOAT64 = 8;  // ncclFloat64
constexpr int NCCL_SUM = 0;      // ncclSum
// returns nullptr (and fills err) when libnccl cannot be loaded
const NcclApi* nccl_api(const char** err);
}  // namespace mg
