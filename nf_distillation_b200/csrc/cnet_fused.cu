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
#include <cstdlib>

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
  // MODE 1 (backward chain): mask1 / mask2 are the INPUT ReLU masks applied to the outputs of GEMM 1 / GEMM 2, and the
  // column sums of those masked outputs (the bias gradients) are accumulated here
  float* colsum1;
  float* colsum2;
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

// MODE 0: forward   h1 = relu(col B1^T + b1), h2 = relu(h1 B2^T + b2)                      (+ optional h1 / mask outputs)
// MODE 1: backward  d2 = mask_a .* (dhcol B3T^T),  d1 = mask_b .* (d2 B2T^T)                (the two dgrads of the chain:
//         Conv2dZeros input gradient -> ReLU mask of h2 -> conv1x1 input gradient -> ReLU mask of h1), both written
//         out as bf16 (the weight-gradient GEMMs read them) with their column sums = the two bias gradients. Same tiles,
//         same pipeline: d2 stays in the shared-memory panels as the A operand of the second GEMM.
template <int MODE>
__device__ __forceinline__ void cnet_fused_body(const CUtensorMap& tmCol, const CUtensorMap& tmB1,
                                                const CUtensorMap& tmB2, const CUtensorMap& tmH1,
                                                const CUtensorMap& tmH2, const CnetArgs& g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_launch_dependents();
  if (smem_u32(smem) & 1023u) __trap();
  const int warp = static_cast<int>(uniform_u32(threadIdx.x >> 5)), lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x & 1;      // == %cluster_ctarank for (2,1,1) clusters on a 1-D grid, provably uniform
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
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  pdl_wait();      // the prologue above overlapped the previous kernel's tail; global memory only from here on

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs; bytes land on the leader's barrier)
    {   // whole warp, uniform control flow; single-thread instructions elected inside the asm (ptx.cuh)
      int s = 0; uint32_t ph = 0;
      auto load = [&](const CUtensorMap* tm, int c0, int c1) {
        mbar_wait_warp(&empty[s], ph ^ 1);
        if (rank == 0) mbar_expect_tx_elect(&full[s], 2 * CF_SLOT);
        tma_load_2d_pair_elect(smem + CnetSmem::ring + s * CF_SLOT, tm, &full[s], c0, c1);
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
    // ===================================================== MMA issuer (leader CTA; whole warp, uniform control flow)
    if (rank == 0) {
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
        mbar_wait_warp(&full[s], ph);
        if (g.prof) w_op += clock64() - c0;
        tc_fence_after();
        return ring_addr + s * CF_SLOT;
      };
      auto advance = [&]() { if (++s == CF_SLOTS) { s = 0; ph ^= 1; } };
      auto release = [&]() { umma_commit_pair_elect(&empty[s], 3); advance(); };
      auto acc_begin = [&]() {
        const uint32_t st = nacc & 1;
        const long long c0 = g.prof ? clock64() : 0;
        mbar_wait_warp(&acc_empty[st], ((nacc >> 1) & 1) ^ 1);
        if (g.prof) w_acc += clock64() - c0;
        tc_fence_after();
        return tmem_base + st * 256;
      };
      auto acc_end = [&]() { umma_commit_pair_elect(&acc_full[nacc & 1], 3); ++nacc; };
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
              umma_f16_pair_elect(d, umma_desc_sw128(a + k * 32, 16, 1024), umma_desc_sw128(b + k * 32, 16, 1024), idesc,
                            (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit_pair_elect(&empty[sa], 3);
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
              mbar_wait_warp(&h1_full[kb], tile_ph);
              if (g.prof) w_h1 += clock64() - c0;
              tc_fence_after();
              panels_seen = kb + 1;
            }
            const uint32_t b = take();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_pair_elect(d, umma_desc_sw128(h1_addr + kb * CF_PANEL + k * 32, 16, 1024),
                            umma_desc_sw128(b + k * 32, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            release();
          }
          acc_end();
        }
        umma_commit_pair_elect(h1_empty, 3);      // h1 is free once every MMA issued so far has retired
        tile_ph ^= 1;
      }
      if (g.prof && lane == 0) {
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
    // ---- MODE 1 pieces: masked copy of 16 accumulator columns, a 64-column panel, and the column sums of a panel
    auto epi16m = [&](const uint32_t (&r)[16], uint32_t bits16, uint8_t* dst, int cc0) {
      uint32_t p[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float lo = ((bits16 >> (2 * k)) & 1u) ? __uint_as_float(r[2 * k]) : 0.f;
        const float hi = ((bits16 >> (2 * k + 1)) & 1u) ? __uint_as_float(r[2 * k + 1]) : 0.f;
        p[k] = pack2(lo, hi);
      }
      *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(cc0) ^ sw) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<uint4*>(dst + ((static_cast<uint32_t>(cc0 + 1) ^ sw) << 4)) =
          make_uint4(p[4], p[5], p[6], p[7]);
    };
    auto panel64m = [&](const uint32_t (&r0)[16], const uint32_t (&r1)[16], const uint32_t (&r2)[16],
                        const uint32_t (&r3)[16], uint32_t w0, uint32_t w1, uint8_t* dst_rows) {
      uint8_t* dst = dst_rows + lane * 128;
      epi16m(r0, w0 & 0xFFFFu, dst, 0); epi16m(r1, w0 >> 16, dst, 2);
      epi16m(r2, w1 & 0xFFFFu, dst, 4); epi16m(r3, w1 >> 16, dst, 6);
    };
    // column sums of this warp's 32 x 64 panel slice, straight from the staged bf16 rows: lane owns the column pair
    // (lane >> 2) * 8 + (lane & 3) * 2 (+1); at row r the 32 lanes read the whole 128-byte row (conflict-free)
    auto panel_colsum = [&](const uint8_t* rows, float& s0, float& s1) {
      const uint8_t* pb = rows + (lane & 3) * 4;
      const uint32_t chunk = static_cast<uint32_t>(lane >> 2);
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; r += 2) {
        const uint32_t w0 = *reinterpret_cast<const uint32_t*>(pb + r * 128 + ((chunk ^ (r & 7)) << 4));
        const uint32_t w1 = *reinterpret_cast<const uint32_t*>(pb + (r + 1) * 128 + ((chunk ^ ((r + 1) & 7)) << 4));
        a0 += __uint_as_float(w0 << 16); a1 += __uint_as_float(w0 & 0xFFFF0000u);
        b0 += __uint_as_float(w1 << 16); b1 += __uint_as_float(w1 & 0xFFFF0000u);
      }
      s0 += a0 + b0; s1 += a1 + b1;
    };
    float cs1[4][2] = {}, cs2[4][2] = {};    // [2 h + pz][column of the pair]: bias-gradient partial sums of this CTA
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
      uint32_t mw1[8], mw2[8];     // MODE 1: this thread's ReLU-mask words of the tile (fetched before the waits)
      if constexpr (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {            // i = 2 h + pz
          const int c1 = (4 * (i >> 1) + 2 * (i & 1) + hf) * 64;            // GEMM 1: panel columns
          const int c2 = (i >> 1) * 256 + hf * 128 + (i & 1) * 64;           // GEMM 2: staged columns
          const bool ok = row < g.M;
          mw1[2 * i] = ok ? __ldg(g.mask1 + static_cast<long long>(c1 >> 5) * g.ldmask + row) : 0u;
          mw1[2 * i + 1] = ok ? __ldg(g.mask1 + static_cast<long long>((c1 + 32) >> 5) * g.ldmask + row) : 0u;
          mw2[2 * i] = ok ? __ldg(g.mask2 + static_cast<long long>(c2 >> 5) * g.ldmask + row) : 0u;
          mw2[2 * i + 1] = ok ? __ldg(g.mask2 + static_cast<long long>((c2 + 32) >> 5) * g.ldmask + row) : 0u;
        }
      }
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
          if constexpr (MODE == 0) {
            if (pz == 0) panel64(R[0], R[1], R[2], R[3], g.bias1 + panel * 64, dst, g.mask1, row, panel * 64);
            else panel64(R[4], R[5], R[6], R[7], g.bias1 + panel * 64, dst, g.mask1, row, panel * 64);
          } else {
            // (h is a run-time loop variable: select with ternaries so the word arrays stay in registers)
            if (pz == 0) panel64m(R[0], R[1], R[2], R[3], h ? mw1[4] : mw1[0], h ? mw1[5] : mw1[1], dst);
            else panel64m(R[4], R[5], R[6], R[7], h ? mw1[6] : mw1[2], h ? mw1[7] : mw1[3], dst);
          }
          fence_proxy_async();
          __syncwarp();
          if constexpr (MODE == 1) {
            float s0 = 0.f, s1 = 0.f;
            panel_colsum(dst, s0, s1);
            if (h == 0) { cs1[pz][0] += s0; cs1[pz][1] += s1; } else { cs1[2 + pz][0] += s0; cs1[2 + pz][1] += s1; }
          }
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
          if constexpr (MODE == 0) {
            if (pz == 0) panel64(R[0], R[1], R[2], R[3], g.bias2 + col0, stage, g.mask2, row, col0);
            else panel64(R[4], R[5], R[6], R[7], g.bias2 + col0, stage, g.mask2, row, col0);
          } else {
            if (pz == 0) panel64m(R[0], R[1], R[2], R[3], h ? mw2[4] : mw2[0], h ? mw2[5] : mw2[1], stage);
            else panel64m(R[4], R[5], R[6], R[7], h ? mw2[6] : mw2[2], h ? mw2[7] : mw2[3], stage);
          }
          fence_proxy_async();
          __syncwarp();
          if constexpr (MODE == 1) {
            float s0 = 0.f, s1 = 0.f;
            panel_colsum(stage, s0, s1);
            if (h == 0) { cs2[pz][0] += s0; cs2[pz][1] += s1; } else { cs2[2 + pz][0] += s0; cs2[2 + pz][1] += s1; }
          }
          if (lane == 0 && row0 < g.M) {
            tma_store_2d(stage, &tmH2, col0, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        ++nacc;
      }
      tile_ph ^= 1;
    }
    if constexpr (MODE == 1) {   // one atomic per column pair, accumulator and CTA
      const int cpair = (lane >> 2) * 8 + (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c1 = (4 * (i >> 1) + 2 * (i & 1) + hf) * 64 + cpair;
        const int c2 = (i >> 1) * 256 + hf * 128 + (i & 1) * 64 + cpair;
        if (g.colsum1) { atomicAdd(g.colsum1 + c1, cs1[i][0]); atomicAdd(g.colsum1 + c1 + 1, cs1[i][1]); }
        if (g.colsum2) { atomicAdd(g.colsum2 + c2, cs2[i][0]); atomicAdd(g.colsum2 + c2 + 1, cs2[i][1]); }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CF_THREADS, 1)
cnet_fwd_fused_kernel(const __grid_constant__ CUtensorMap tmCol, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmH1,
                      const __grid_constant__ CUtensorMap tmH2, const CnetArgs g) {
  cnet_fused_body<0>(tmCol, tmB1, tmB2, tmH1, tmH2, g);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CF_THREADS, 1)
cnet_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmCol, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmH1,
                      const __grid_constant__ CUtensorMap tmH2, const CnetArgs g) {
  cnet_fused_body<1>(tmCol, tmB1, tmB2, tmH1, tmH2, g);
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

// cnet_ts.cu: the same two GEMMs with h1 resident in tensor memory (A operand of conv#2 read from TMEM). Default path;
// NFK_CNET_TS=0 in the environment selects the shared-memory-panel kernels of this file instead.
int cnet_ts_fwd(const void* col, int K1p, const void* B1, const void* B2, const float* bias1, const float* bias2,
                void* h1, void* h2, void* mask1, void* mask2, long long ldmask, int M, int kb2_end_half0,
                long long* prof, void* stream);
int cnet_ts_bwd(const void* dhcol, int K3p, const void* B3T, const void* B2T, const void* mask_h2,
                const void* mask_h1, long long ldmask, void* dpre2, void* dpre1, float* dbias2, float* dbias1, int M,
                void* stream);
static bool cnet_use_ts() {
  static const bool v = [] { const char* e = getenv("NFK_CNET_TS"); return !(e && e[0] == '0'); }();
  return v;
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
  if (cnet_use_ts())
    return cnet_ts_fwd(col, K1p, B1, B2, bias1, bias2, h1, h2, mask1, mask2, ldmask, M, kb2_end_half0, g_cnet_prof,
                       stream);
  CnetArgs g{M, K1p / 64, bias1, bias2, static_cast<uint32_t*>(mask1), static_cast<uint32_t*>(mask2), ldmask,
             h1 ? 1 : 0, g_cnet_prof, {kb2_end_half0, 8}, nullptr, nullptr};
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

extern "C" int nfk_cnet_bwd_fused(const void* dhcol, int K3p, const void* B3T, const void* B2T, const void* mask_h2,
                                  const void* mask_h1, long long ldmask, void* dpre2, void* dpre1, float* dbias2,
                                  float* dbias1, int M, int hid, void* stream) {
  if (M <= 0 || hid != CF_HID || K3p % 64 || K3p < 64 || K3p > 512) return NFK_ERR_SHAPE;
  if (!dhcol || !B3T || !B2T || !mask_h2 || !mask_h1 || !dpre2 || !dpre1) return NFK_ERR_ARG;
  if (ldmask < M) return NFK_ERR_ARG;
  if (cnet_use_ts())
    return cnet_ts_bwd(dhcol, K3p, B3T, B2T, mask_h2, mask_h1, ldmask, dpre2, dpre1, dbias2, dbias1, M, stream);
  CnetArgs g{M, K3p / 64, nullptr, nullptr, static_cast<uint32_t*>(const_cast<void*>(mask_h2)),
             static_cast<uint32_t*>(const_cast<void*>(mask_h1)), ldmask, 1, nullptr, {8, 8}, dbias2, dbias1};
  CUtensorMap tmA, tmB1, tmB2, tmD2, tmD1;
  int rc;
  if ((rc = cf_tmap(&tmA, dhcol, K3p, M, K3p, 128, false))) return rc;
  if ((rc = cf_tmap(&tmB1, B3T, K3p, hid, K3p, 128, false))) return rc;
  if ((rc = cf_tmap(&tmB2, B2T, hid, hid, hid, 128, false))) return rc;
  if ((rc = cf_tmap(&tmD2, dpre2, hid, M, hid, 32, false))) return rc;
  if ((rc = cf_tmap(&tmD1, dpre1, hid, M, hid, 32, false))) return rc;
  if ((rc = ensure_dyn_smem(reinterpret_cast<const void*>(cnet_bwd_fused_kernel), CnetSmem::total))) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (M + 255) / 256;
  const int pairs = tiles < sms / 2 ? tiles : sms / 2;
  const cudaError_t le = launch_pdl(cnet_bwd_fused_kernel, dim3(2 * pairs), dim3(CF_THREADS), CnetSmem::total,
                                    static_cast<cudaStream_t>(stream), tmA, tmB1, tmB2, tmD2, tmD1, g);
  return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? NFK_OK : NFK_ERR_LAUNCH;
}
