// Microbenchmark: cycles per tcgen05.mma (cta_group::2, M = 256, bf16) for N in {128, 256}, A from shared memory or
// from tensor memory, one accumulator (dependent chain) or two alternating accumulators. Operand contents are
// whatever shared memory / TMEM hold (timing only).   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../nf_distillation_b200/csrc/ptx.cuh"
using namespace nfk;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// variant: N, a_tmem, nacc (accumulators used round-robin), rot (how many distinct 16 KB B slots are cycled)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
bench(int N, int a_tmem, int nacc, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc_pair(&slot, 512); tmem_relinquish_pair(); }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, N, false, false);
    const uint32_t base = smem_u32(smem);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a_s = base + (i & 7) * 16384 + ((i >> 3) & 3) * 32;          // 8 A panels
      const uint32_t b_s = base + 131072 + (i & 3) * 16384 + ((i >> 2) & 3) * 32;  // 4 B slots
      // accumulators: nacc == 1 -> cols [256, 256+N); nacc == 2 -> alternate two N-wide regions (needs 2N <= 256 if a_tmem)
      const uint32_t d = tb + 256 + (nacc == 2 ? (i & 1) * 128 : 0);
      if (a_tmem) mma_ts(d, tb + (i & 31) * 8, umma_desc_sw128(b_s, 16, 1024), idesc, i >= nacc ? 1u : 0u);
      else umma_f16_pair(d, umma_desc_sw128(a_s, 16, 1024), umma_desc_sw128(b_s, 16, 1024), idesc, i >= nacc ? 1u : 0u);
    }
    umma_commit_pair(&bar, 1);
    mbar_wait(&bar, 0);
    out[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 74 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  struct V { int N, a_tmem, nacc; const char* name; } vs[] = {
    {256, 0, 1, "SS N=256 one accumulator"}, {128, 0, 1, "SS N=128 one accumulator"}, {128, 0, 2, "SS N=128 two accumulators alternating"},
    {256, 1, 1, "TS N=256 one accumulator"}, {128, 1, 1, "TS N=128 one accumulator"}, {128, 1, 2, "TS N=128 two accumulators alternating"},
    {64, 1, 1, "TS N=64 one accumulator"}, {64, 0, 1, "SS N=64 one accumulator"}};
  for (int grid : {2, 148}) for (auto& v : vs) {
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<grid, 128, 200 * 1024>>>(v.N, v.a_tmem, v.nacc, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
    }
    long long h[74]; cudaMemcpy(h, d, (grid / 2) * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < grid / 2; ++i) s += h[i];
    printf("grid %3d  %-42s %.1f cycles per MMA (256xNx16)\n", grid, v.name, s / (grid / 2) / iters);
  }
  return 0;
}
