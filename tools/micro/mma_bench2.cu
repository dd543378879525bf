// Microbenchmark 2: what slows tcgen05.mma (cta_group::2, 256x256x16, SS) down when the rest of the fused kernel's
// traffic runs beside it?  Side traffic (each optional): epilogue-like warps streaming tcgen05.ld over 256 TMEM columns,
// warps writing 16-byte chunks into shared memory, one thread streaming bulk copies global -> shared memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../nf_distillation_b200/csrc/ptx.cuh"
using namespace nfk;

struct Out { long long mma_cycles, ld_cycles, ld_count, st_count, tma_count; long long pad[3]; };

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// warps: 0 = alloc + bulk-copy producer, 1 = MMA issuer, 2..9 = tcgen05.ld warps, 10..13 = st.shared warps
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(448, 1)
bench(int N, int iters, int ld_warps, int st_warps, int tma_on, int ld_same, const uint8_t* gsrc, Out* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, tbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&tbar[i], 1); done = 0; fence_mbar_init(); }
  if (warp == 0) { tmem_alloc_pair(&slot, 512); tmem_relinquish_pair(); }
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = slot;
  Out* o = out + blockIdx.x;
  if (warp == 1) {
    if (lane == 0) {
      if (rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(256, N, false, false);
        const uint32_t base = smem_u32(smem);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          const uint32_t a_s = base + (i & 7) * 16384 + ((i >> 3) & 3) * 32;
          const uint32_t b_s = base + 131072 + (i & 1) * 16384 + ((i >> 1) & 3) * 32;
          umma_f16_pair(tb + 256, umma_desc_sw128(a_s, 16, 1024), umma_desc_sw128(b_s, 16, 1024), idesc, i ? 1u : 0u);
        }
        umma_commit_pair(&bar, 3);
        mbar_wait(&bar, 0);
        o->mma_cycles = clock64() - t0;
      } else {
        mbar_wait(&bar, 0);
      }
      done = 1;
    }
  } else if (warp == 0) {
    if (lane == 0 && tma_on) {   // 16 KB bulk copies into slots at 160 KB.. (2 slots), back to back
      long long n = 0; uint32_t ph[2] = {0, 0};
      while (!done) {
        for (int s = 0; s < 2; ++s) {
          mbar_expect_tx(&tbar[s], 16384);
          bulk_load(smem + 163840 + s * 16384, gsrc + ((n * 2 + s) & 63) * 16384, 16384, &tbar[s]);
        }
        for (int s = 0; s < 2; ++s) { mbar_wait(&tbar[s], ph[s]); ph[s] ^= 1; }
        n += 2;
      }
      o->tma_count = n;
    }
  } else if (warp < 10) {
    if (warp - 2 < ld_warps) {
      const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
      const uint32_t colbase = ld_same ? 256u : 0u;
      long long n = 0; uint32_t acc = 0;
      const long long t0 = clock64();
      while (!done) {
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          uint32_t R[8][16];
#pragma unroll
          for (int i = 0; i < 8; ++i) tmem_ld16(tb + lane_off + colbase + ((warp - 2) >> 2) * 128 + 16 * i, R[i]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) acc ^= R[i][lane & 15];
        }
        n += 16;
      }
      if (lane == 0) { o[0].pad[0] = acc; if (warp == 2) { o->ld_cycles = clock64() - t0; o->ld_count = n; } }
    }
  } else {
    if (warp - 10 < st_warps) {
      uint8_t* dst = smem + 131072 + 32768 + (warp - 10) * 4096 + lane * 128;   // private 4 KB panel at 160 KB+..(no overlap w/ B)
      dst = smem + 196608 - 16384 + (warp - 10) * 4096 + lane * 128;
      long long n = 0;
      const uint32_t sw = lane & 7;
      while (!done) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(c) ^ sw) << 4)) = make_uint4(n, c, lane, warp);
        n += 8;
      }
      if (lane == 0 && warp == 10) o->st_count = n;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tb, 512);
}

int main() {
  Out* d; cudaMalloc(&d, 148 * sizeof(Out));
  uint8_t* src; cudaMalloc(&src, 64 * 16384); cudaMemset(src, 0x3c, 64 * 16384);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 8192;
  struct V { int N, ld, st, tma, same; const char* name; } vs[] = {
    {256, 0, 0, 0, 0, "MMA alone"},
    {256, 4, 0, 0, 0, "+ 4 warps tcgen05.ld (other columns)"},
    {256, 8, 0, 0, 0, "+ 8 warps tcgen05.ld (other columns)"},
    {256, 8, 0, 0, 1, "+ 8 warps tcgen05.ld (accumulator columns)"},
    {256, 0, 4, 0, 0, "+ 4 warps st.shared.v4"},
    {256, 0, 0, 1, 0, "+ bulk copies global->smem"},
    {256, 8, 4, 1, 0, "+ all three"},
    {128, 0, 0, 0, 0, "N=128 MMA alone"},
    {128, 8, 4, 1, 0, "N=128 + all three"},
  };
  for (auto& v : vs) {
    cudaMemset(d, 0, 148 * sizeof(Out));
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<148, 448, 200 * 1024>>>(v.N, iters, v.ld, v.st, v.tma, v.same, src, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", v.name, cudaGetErrorString(e)); return 1; }
    }
    Out h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double mc = 0, ldb = 0, stb = 0, tmab = 0;
    for (int i = 0; i < 148; i += 2) {
      mc += h[i].mma_cycles;
      const double cyc = h[i].mma_cycles;
      ldb += h[i].ld_count * 32.0 * 64 * v.ld / cyc;     // bytes per clock per SM: each x16 ld = 32 lanes x 64 B per warp
      stb += h[i].st_count * 512.0 * v.st / cyc;
      tmab += h[i].tma_count * 16384.0 / cyc;
    }
    printf("%-46s %6.1f cycles/MMA | per SM: tcgen05.ld %6.1f B/clk, st.shared %6.1f B/clk, bulk %6.1f B/clk\n", v.name,
           mc / 74 / iters, ldb / 74, stb / 74, tmab / 74);
  }
  return 0;
}
