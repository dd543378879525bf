// Host-side launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

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

}  // namespace nfk
