// Host-side launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include <mutex>
#include <unordered_map>

#include "../../include/nfk.h"

namespace nfk {

// Opt a kernel in to > 48 KB of dynamic shared memory exactly once (keyed by the kernel's address), so later launches
// — including launches made while a CUDA graph is being captured — issue no attribute call.
inline int ensure_dyn_smem(const void* kernel, int bytes) {
  if (bytes <= 48 * 1024) return NFK_OK;
  if (bytes > 227 * 1024) return NFK_ERR_SHAPE;
  static std::mutex mu;
  static std::unordered_map<const void*, bool> done;
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(kernel);
  if (it != done.end()) return NFK_OK;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return NFK_ERR_LAUNCH;
  done[kernel] = true;
  return NFK_OK;
}

// Launch with programmatic dependent launch allowed (see ptx.cuh: pdl_wait / pdl_launch_dependents). NFK_PDL=0 in the
// environment turns the attribute off (the kernels' griddepcontrol instructions are then no-ops).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NFK_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace nfk
