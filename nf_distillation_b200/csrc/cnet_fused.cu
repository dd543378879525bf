// Fused conv#1 -> conv#2 of the coupling network (forward): for a block of 256 pixels (one CTA pair, 128 rows per CTA)
//
//   h1 = relu(col1 * B1^T + b1)      conv3x3 as im2col GEMM          (reference models/flows.py:27-28)
//   h2 = relu(h1   * B2^T + b2)      conv1x1                         (:29-30)
//
// run back to back on the tensor cores WITHOUT h1 leaving the SM: the epilogue warps write the bf16 activations
// straight into 128B-swizzled K-major shared-memory panels that conv#2's tcgen05.mma reads as its A operand, so
// conv#2 streams only its weights (from L2). h2 goes out through TMA stores; the training path additionally stores
// h1 and the 1-bit ReLU masks. (The zero-init conv3x3 stays a separate GEMM: with P's accumulator in TMEM the conv
// quarters would have to be N = 128 MMAs, which measured 143 cycles instead of 64 — shared-memory operand bandwidth.)
//
//   smem   h1: 8 panels x 16 KB (128 rows x 64 ch) | ring: 4 x 16 KB operand slots | 8 x 4 KB store staging
//   TMEM   2 x 256 columns: ping-pong accumulator over the four 256-channel half-GEMMs of a tile
//   warps  0: TMA producer, 1: MMA issuer (leader CTA) + TMEM alloc, 2-9: epilogue (2 per TMEM lane quadrant)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"
#include "launch_util.h"
#include "ptx.cuh"

namespace nfk {

constexpr int CF_THREADS = 320;
constexpr int CF_HID = 512;
constexpr int CF_SLOT = 16384;      // ring slot: one 128-row x 128-byte operand tile
constexpr int CF_SLOTS = 4;
constexpr int CF_PANEL = 16384;     // 128 rows x 64 bf16

struct CnetArgs {
  int M;        // pixels
  int kb1;      // K1p / 64
  const float* bias1;
  const float* bias2;
  uint32_t* mask1;   // optional 1-bit ReLU masks, word-major [16][ldmask]
  uint32_t* mask2;
  long long ldmask;
  int store_h1;
  long long* prof;   // diagnostics: [grid][8] cycle counters of the MMA issuer
  // conv#2 k-blocks (64 input channels each) that are not structurally zero for output half 0 / 1; 8 = all. MADE's
  // degree-sorted hidden mask is block lower triangular: the low-degree half of the outputs never sees the high-degree
  // inputs, so those B2 tiles are neither loaded nor multiplied (they are exact zeros: the result is bit-identical).
  int kb2_end[2];
};

struct CnetSmem {
  static constexpr int h1 = 0;
  static constexpr int ring = h1 + 8 * CF_PANEL;
  static constexpr int stage = ring + CF_SLOTS * CF_SLOT;
  static constexpr int bars = stage + 8 * 4096;
  static constexpr int total = bars + 256;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CF_THREADS, 1)
cnet_fwd_fused_kernel(const __grid_constant__ CUtensorMap tmCol, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmH1,
                      const __grid_constant__ CUtensorMap tmH2, const CnetArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  if (smem_u32(smem) & 1023u) __trap();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = (g.M + 255) / 256;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CnetSmem::bars);   // [4] operand slot filled (leader)
  uint64_t* empty = full + CF_SLOTS;                                      // [4] operand slot consumed (local)
  uint64_t* acc_full = empty + CF_SLOTS;                                  // [2]
  uint64_t* acc_empty = acc_full + 2;                                     // [2] (leader, 16 arrivals)
  uint64_t* h1_full = acc_empty + 2;                                      // [8] per 64-channel h1 panel (leader, 8 arrivals)
  uint64_t* h1_empty = h1_full + 8;                                       // conv#2 of the tile has finished reading h1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h1_empty + 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmCol); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmH2);
    for (int s = 0; s < CF_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 16); }
    for (int p = 0; p < 8; ++p) mbar_init(&h1_full[p], 8);
    mbar_init(h1_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the prologue above overlapped the previous kernel's tail; global memory only from here on

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs; bytes land on the leader's barrier)
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* tm, int c0, int c1) {
        mbar_wait(&empty[s], ph ^ 1);
        if (rank == 0) mbar_expect_tx(&full[s], 2 * CF_SLOT);
        tma_load_2d_pair(smem + CnetSmem::ring + s * CF_SLOT, tm, &full[s], c0, c1);
        if (++s == CF_SLOTS) { s = 0; ph ^= 1; }
      };
      for (int t = pair; t < num_tiles; t += num_pairs) {
        const int row0 = t * 256 + static_cast<int>(rank) * 128;
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < g.kb1; ++kb) {
            load(&tmCol, kb * 64, row0);
            load(&tmB1, kb * 64, h * 256 + static_cast<int>(rank) * 128);
          }
        for (int h = 0; h < 2; ++h)
          for (int kb = 0; kb < g.kb2_end[h]; ++kb) load(&tmB2, kb * 64, h * 256 + static_cast<int>(rank) * 128);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA, one thread)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(256, 256, false, false);
      const uint32_t ring_addr = smem_u32(smem + CnetSmem::ring);
      const uint32_t h1_addr = smem_u32(smem + CnetSmem::h1);
      int s = 0; uint32_t ph = 0;
      uint32_t nacc = 0;          // accumulator uses so far (stage = nacc & 1, phase = (nacc >> 1) & 1)
      uint32_t tile_ph = 0;
      long long w_op = 0, w_acc = 0, w_h1 = 0;
      const long long t_begin = g.prof ? clock64() : 0;
      auto take = [&]() {
        const long long c0 = g.prof ? clock64() : 0;
        mbar_wait(&full[s], ph);
        if (g.prof) w_op += clock64() - c0;
        tc_fence_after();
        return ring_addr + s * CF_SLOT;
      };
      auto advance = [&]() { if (++s == CF_SLOTS) { s = 0; ph ^= 1; } };
      auto release = [&]() { umma_commit_pair(&empty[s], 3); advance(); };
      auto acc_begin = [&]() {
        const uint32_t st = nacc & 1;
        const long long c0 = g.prof ? clock64() : 0;
        mbar_wait(&acc_empty[st], ((nacc >> 1) & 1) ^ 1);
        if (g.prof) w_acc += clock64() - c0;
        tc_fence_after();
        return tmem_base + st * 256;
      };
      auto acc_end = [&]() { umma_commit_pair(&acc_full[nacc & 1], 3); ++nacc; };
      for (int t = pair; t < num_tiles; t += num_pairs) {
        // ---- conv#1: two 256-channel halves; each k-step consumes two slots (im2col tile, B1 tile)
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = acc_begin();
          for (int kb = 0; kb < g.kb1; ++kb) {
            const uint32_t a = take();
            const int sa = s;
            advance();
            const uint32_t b = take();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_pair(d, umma_desc_sw128(a + k * 32, 16, 1024), umma_desc_sw128(b + k * 32, 16, 1024), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair(&empty[sa], 3);
            release();
          }
          acc_end();
        }
        // ---- conv#2: A = the h1 panels the epilogue warps just wrote (both CTAs), B2 streamed
        //      k-block kb only needs h1 panel kb, so conv#2 starts as soon as the first panels have been written
        int panels_seen = 0;      // h1 panels of this tile already waited for
        for (int h = 0; h < 2; ++h) {
          const uint32_t d = acc_begin();
          for (int kb = 0; kb < g.kb2_end[h]; ++kb) {
            if (kb >= panels_seen) {
              const long long c0 = g.prof ? clock64() : 0;
              mbar_wait(&h1_full[kb], tile_ph);
              if (g.prof) w_h1 += clock64() - c0;
              tc_fence_after();
              panels_seen = kb + 1;
            }
            const uint32_t b = take();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_pair(d, umma_desc_sw128(h1_addr + kb * CF_PANEL + k * 32, 16, 1024),
                            umma_desc_sw128(b + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            release();
          }
          acc_end();
        }
        umma_commit_pair(h1_empty, 3);      // h1 is free once every MMA issued so far has retired
        tile_ph ^= 1;
      }
      if (g.prof) {
        long long* o = g.prof + blockIdx.x * 8;
        o[0] = clock64() - t_begin; o[1] = w_op; o[2] = w_acc; o[3] = w_h1;
      }
    }
  } else {
    // ===================================================== epilogue warps
    const int qd = warp & 3;               // TMEM lane quadrant
    const int hf = (warp - 2) >> 2;        // which 128-column half of a 256-column accumulator
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint8_t* stage = smem + CnetSmem::stage + (warp - 2) * 4096;
    uint32_t nacc = 0, tile_ph = 0;
    // 16 accumulator columns (already in registers) -> bias + ReLU -> bf16 -> two 16-byte chunks (cc0, cc0 + 1) of
    // this thread's 128-byte row in a swizzled panel; returns the 16 ReLU mask bits
    auto epi16 = [&](const uint32_t (&r)[16], const float* bias16, uint8_t* dst, int cc0, bool want_bits) -> uint32_t {
      // packed arithmetic: bias add as add.f32x2, round to bf16x2, ReLU on the packed pair (max.bf16x2 with +0 equals
      // rounding the fp32 ReLU: rounding is monotonic and keeps the sign) -- 24 issue slots per 16 columns instead of 40
      uint32_t p[8];
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias16 + j));
        const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                     make_float2(b.x, b.y));
        const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])),
                                     make_float2(b.z, b.w));
        const __nv_bfloat162 h0 = __hmax2(__float22bfloat162_rn(s0), zero2);
        const __nv_bfloat162 h1 = __hmax2(__float22bfloat162_rn(s1), zero2);
        p[j / 2] = *reinterpret_cast<const uint32_t*>(&h0);
        p[j / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
      }
      uint32_t bits = 0;
      if (want_bits) {   // value > 0  <=>  the (non-negative) bf16 is not +0
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          bits |= ((p[k] & 0xFFFFu) ? 1u : 0u) << (2 * k);
          bits |= ((p[k] >> 16) ? 1u : 0u) << (2 * k + 1);
        }
      }
      *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(cc0) ^ sw) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(cc0 + 1) ^ sw) << 4)) =
          make_uint4(p[4], p[5], p[6], p[7]);
      return bits;
    };
    // 64 columns held in R[4 pz .. 4 pz + 3] -> one swizzled panel slice (32 rows x 128 B) at `dst_rows`, plus the two
    // 1-bit ReLU mask words of those columns
    auto panel64 = [&](const uint32_t (&r0)[16], const uint32_t (&r1)[16], const uint32_t (&r2)[16],
                       const uint32_t (&r3)[16], const float* bias, uint8_t* dst_rows, uint32_t* mask, long long row,
                       int col0) {
      uint8_t* dst = dst_rows + lane * 128;
      const bool wb = mask != nullptr;
      const uint32_t b0 = epi16(r0, bias, dst, 0, wb), b1 = epi16(r1, bias + 16, dst, 2, wb);
      const uint32_t b2 = epi16(r2, bias + 32, dst, 4, wb), b3 = epi16(r3, bias + 48, dst, 6, wb);
      if (wb && row < g.M) {
        mask[static_cast<long long>(col0 >> 5) * g.ldmask + row] = b0 | (b1 << 16);
        mask[static_cast<long long>((col0 + 32) >> 5) * g.ldmask + row] = b2 | (b3 << 16);
      }
    };
    // this warp's 128 accumulator columns -> registers, all loads in flight at once; the accumulator stage can be
    // handed back to the MMA issuer as soon as they have landed, before any of the epilogue math
    auto load128 = [&](uint32_t tm, uint32_t (&R)[8][16]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tmem_ld16(tm + 16 * i, R[i]);
      tmem_ld_wait();
    };
    for (int t = pair; t < num_tiles; t += num_pairs) {
      const int row0 = t * 256 + static_cast<int>(rank) * 128 + qd * 32;
      const long long row = row0 + lane;
      // h1 of the previous tile must be dead (its conv#2 MMAs retired); its optional h1 stores must have been read
      mbar_wait(h1_empty, tile_ph ^ 1);
      if (g.store_h1) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
      // ---- conv#1 halves -> h1 panels. This warp takes panels 4h + hf and 4h + 2 + hf, so after the first round the
      //      two lowest panels of the half exist and conv#2 can start on them (its k-block kb reads panel kb)
      for (int h = 0; h < 2; ++h) {
        const uint32_t st = nacc & 1;
        mbar_wait(&acc_full[st], (nacc >> 1) & 1);
        tc_fence_after();
        const uint32_t tm = tmem_base + st * 256 + lane_off;
        uint32_t R[8][16];
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_ld16(tm + hf * 64 + 16 * i, R[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_ld16(tm + (hf + 2) * 64 + 16 * i, R[4 + i]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&acc_empty[st], 0);
#pragma unroll
        for (int pz = 0; pz < 2; ++pz) {
          const int panel = 4 * h + 2 * pz + hf;
          uint8_t* dst = smem + CnetSmem::h1 + panel * CF_PANEL + qd * 32 * 128;
          if (pz == 0) panel64(R[0], R[1], R[2], R[3], g.bias1 + panel * 64, dst, g.mask1, row, panel * 64);
          else panel64(R[4], R[5], R[6], R[7], g.bias1 + panel * 64, dst, g.mask1, row, panel * 64);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(&h1_full[panel], 0);
            if (g.store_h1 && row0 < g.M) tma_store_2d(dst, &tmH1, panel * 64, row0);
          }
        }
        if (g.store_h1 && lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        ++nacc;
      }
      // ---- conv#2 halves -> staging panel -> TMA store of h2
      for (int h = 0; h < 2; ++h) {
        const uint32_t st = nacc & 1;
        mbar_wait(&acc_full[st], (nacc >> 1) & 1);
        tc_fence_after();
        const uint32_t tm = tmem_base + st * 256 + hf * 128 + lane_off;
        uint32_t R[8][16];
        load128(tm, R);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&acc_empty[st], 0);   // the MMAs of the next half may overwrite it now
#pragma unroll
        for (int pz = 0; pz < 2; ++pz) {
          const int col0 = h * 256 + hf * 128 + pz * 64;
          // the staging panel is reused: the previous bulk store must have finished reading it. (With store_h1 the
          // same wait also covers the h1 stores, which only READ the h1 panels that conv#2 is reading anyway.)
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          if (pz == 0) panel64(R[0], R[1], R[2], R[3], g.bias2 + col0, stage, g.mask2, row, col0);
          else panel64(R[4], R[5], R[6], R[7], g.bias2 + col0, stage, g.mask2, row, col0);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && row0 < g.M) {
            tma_store_2d(stage, &tmH2, col0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        ++nacc;
      }
      tile_ph ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int cf_tmap(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows,
                   bool f32) {
  static EncodeTiledFn2 enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return NFK_ERR_DRIVER;
    enc = reinterpret_cast<EncodeTiledFn2>(p);
  }
  const uint64_t es = f32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * es) % 16) return NFK_ERR_ALIGN;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
             dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? NFK_OK
             : NFK_ERR_DRIVER;
}

}  // namespace nfk

using namespace nfk;

static long long* g_cnet_prof = nullptr;
extern "C" int nfk_cnet_set_prof(void* buf) {
  g_cnet_prof = static_cast<long long*>(buf);
  return NFK_OK;
}

extern "C" int nfk_cnet_fwd_fused(const void* col, int K1p, const void* B1, const void* B2, const float* bias1,
                                  const float* bias2, void* h1, void* h2, void* mask1, void* mask2,
                                  long long ldmask, int M, int hid, void* stream) {
  return nfk_cnet_fwd_fused_ranged(col, K1p, B1, B2, bias1, bias2, h1, h2, mask1, mask2, ldmask, M, hid, 8, stream);
}

extern "C" int nfk_cnet_fwd_fused_ranged(const void* col, int K1p, const void* B1, const void* B2, const float* bias1,
                                         const float* bias2, void* h1, void* h2, void* mask1, void* mask2,
                                         long long ldmask, int M, int hid, int kb2_end_half0, void* stream) {
  if (M <= 0 || hid != CF_HID || K1p % 64 || K1p < 64 || K1p > 512) return NFK_ERR_SHAPE;
  if (kb2_end_half0 < 1 || kb2_end_half0 > 8) return NFK_ERR_ARG;
  if (!col || !B1 || !B2 || !bias1 || !bias2 || !h2) return NFK_ERR_ARG;
  if ((mask1 || mask2) && ldmask < M) return NFK_ERR_ARG;
  CnetArgs g{M, K1p / 64, bias1, bias2, static_cast<uint32_t*>(mask1), static_cast<uint32_t*>(mask2), ldmask,
             h1 ? 1 : 0, g_cnet_prof, {kb2_end_half0, 8}};
  CUtensorMap tmCol, tmB1, tmB2, tmH1, tmH2;
  int rc;
  if ((rc = cf_tmap(&tmCol, col, K1p, M, K1p, 128, false))) return rc;
  if ((rc = cf_tmap(&tmB1, B1, K1p, hid, K1p, 128, false))) return rc;
  if ((rc = cf_tmap(&tmB2, B2, hid, hid, hid, 128, false))) return rc;
  if ((rc = cf_tmap(&tmH2, h2, hid, M, hid, 32, false))) return rc;
  if (h1) { if ((rc = cf_tmap(&tmH1, h1, hid, M, hid, 32, false))) return rc; }
  else tmH1 = tmH2;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(cnet_fwd_fused_kernel), CnetSmem::total))) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (M + 255) / 256;
  const int pairs = tiles < sms / 2 ? tiles : sms / 2;
  const cudaError_t le = launch_pdl(cnet_fwd_fused_kernel, dim3(2 * pairs), dim3(CF_THREADS), CnetSmem::total,
                                    static_cast<cudaStream_t>(stream), tmCol, tmB1, tmB2, tmH1, tmH2, g);
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? NFK_OK : NFK_ERR_LAUNCH;
}
