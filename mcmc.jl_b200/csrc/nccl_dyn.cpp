#include "nccl_dyn.h"
#include <dlfcn.h>
#include <cstdlib>
namespace mg {
static NcclApi g_api;
static bool g_tried = false;
static const char* g_err = nullptr;
const NcclApi* nccl_api(const char** err) {
  if (!g_tried) {
    g_tried = true;
    const char* names[] = {getenv("MCMCGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      g_api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_api.handle) break;
    }
    if (!g_api.handle) { g_err = "libnccl.so.2 not found (set MCMCGPU_NCCL_LIB)"; }
    else {
      g_api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(g_api.handle, "ncclGetUniqueId");
      g_api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(g_api.handle, "ncclCommInitRank");
      g_api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(g_api.handle, "ncclAllReduce");
      g_api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(g_api.handle, "ncclAllGather");
      g_api.CommDestroy = (int (*)(NcclComm))dlsym(g_api.handle, "ncclCommDestroy");
      g_api.GetErrorString = (const char* (*)(int))dlsym(g_api.handle, "ncclGetErrorString");
      if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.AllReduce || !g_api.AllGather || !g_api.CommDestroy) g_err = "libnccl lacks required symbols";
    }
  }
  if (g_err) { if (err) *err = g_err; return nullptr; }
  return &g_api;
}
}  // namespace mg
