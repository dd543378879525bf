// MADE autoregressive inverse with the activations resident in shared memory (north_star kernel (4)).
//
// The reference repository ships no MAF/MADE code, so this follows Papamakarios et al. 2017 eq. 3:
//   x_i = u_i * exp(alpha_i(x_<i)) + mu_i(x_<i),  i = 0 .. D-1,   log|det| += sum_i alpha_i.
// Instead of D full passes of the three masked linears (D x one forward's FLOPs, activations through HBM / L2), the
// pre-activation of every hidden unit is finalised exactly ONCE: hidden units are sorted by degree, so after x_{d-1}
// is known the layer-1 units of degree d are final, then the layer-2 units of degree d (they read layer-1 units of
// degree <= d only), then (mu_d, alpha_d) (layer-2 units of degree <= d). Total work ~ one forward pass.
//
// A CTA owns warps x MT*16 samples for the whole D-step recursion. Each WARP keeps x (bf16), h1 and h2 (bf16, the
// same roundings the forward GEMM epilogues apply) of its own samples in its slice of shared memory; every product is
// a warp-level mma.sync.m16n8k16 (bf16 operands, fp32 accumulation) with the activations as the A operand and the
// masked bf16 weights as the B operand, both through ldmatrix. tcgen05 does not fit this recursion: a step touches
// 8-16 output columns of a 16-row tile, far below its 64 x 8 x 16 minimum shape with a TMEM round trip per step.
// The recursion is a fixed stream of JOBS (step d: layer-1 tile pairs of degree d, layer-2 tile pairs, the
// (mu_d, alpha_d) row pair), the same for every sample tile, so the host builds the job table once
// (nfk_made_inverse_jobs). A producer warp streams each job's weight rows into a 3-stage shared-memory ring with
// cp.async.bulk (TMA, one bulk copy per row, completion on an mbarrier); the consumer warps wait on the stage's
// "full" barrier, multiply, and release it through its "empty" barrier — no block-wide barrier in the recursion,
// warps drift apart freely. History (B = 65 536, D = 63, H = 512): every warp streaming its B fragments straight
// from L2 (with 213 KB of shared memory carved out there is no L1 left): 1.62 ms; cp.async ring filled by all
// threads + __syncthreads per job + the job stream derived on the fly by every thread: 1.12 ms, issue-bound on that
// bookkeeping (ncu: 1 000 warp instructions per warp and step, 60 % of them index arithmetic).
// Units are processed in aligned 8-column tiles; a tile that straddles two degrees is evaluated at both steps (the
// not-yet-final columns hold finite scratch values that only ever meet masked-zero weights before being overwritten).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"
#include "ptx.cuh"

namespace nfk {

constexpr int MI_STAGES = 3;

struct MiArgs {
  const float* u_in;            // [B, D] layer output order (flipped if flip)
  const __nv_bfloat16* B1;      // [H, Dp]   masked, k contiguous
  const __nv_bfloat16* B2;      // [H, H]
  const __nv_bfloat16* B3;      // [N3p, H]  rows: mu_0..mu_{D-1}, alpha_0..alpha_{D-1}
  const float *b1, *b2, *b3;    // [H], [H], [>= 2D]
  const int4* jobs;             // [njobs] {phase, row0 (phase 2: d), k-chunks of 16, second tile present}
  int njobs;
  float* x;                     // [B, D]
  const float* ld_in;           // [B] or null
  float* ld_out;                // [B] or null
  int B, D, H, Dp, flip;
};

__device__ __forceinline__ void mi_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}

__device__ __forceinline__ void mi_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// one row of weights, global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void mi_bulk_row(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t mi_pack_relu(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(fmaxf(a, 0.f), fmaxf(b, 0.f));
  return *reinterpret_cast<const uint32_t*>(&v);
}

// blockDim = (consumer warps + 1) * 32; the last warp is the weight producer
template <int MT>
__global__ void __launch_bounds__(288) made_inverse_resident_kernel(const MiArgs p) {
  extern __shared__ __align__(128) unsigned char mi_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, cwarps = (blockDim.x >> 5) - 1;
  const int D = p.D, H = p.H, Dp = p.Dp;
  const int ldx = Dp + 8, ldh = H + 8;                       // bf16 elements; +8 keeps ldmatrix rows on distinct banks
  const int ldb_bytes = ((H > Dp ? H : Dp) + 8) * 2;         // weight ring row stride
  const int stage_bytes = 16 * ldb_bytes;
  constexpr int R = MT * 16;
  const int per_warp = R * (ldx + 2 * ldh) * 2;              // bytes
  unsigned char* ring = mi_smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(mi_smem + MI_STAGES * stage_bytes);
  uint64_t* empty = full + MI_STAGES;
  unsigned char* act = mi_smem + MI_STAGES * stage_bytes + 128;
  const uint32_t ring_s = smem_u32(ring);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MI_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], cwarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int ctiles = (p.B + R * cwarps - 1) / (R * cwarps);   // CTA tiles of cwarps * R samples
  const int my_tiles = (ctiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int njobs = p.njobs;

  if (warp == cwarps) {
    // ---------------- producer: the same job stream once per CTA tile, MI_STAGES jobs ahead of the slowest consumer
    int stage = 0, par = 0;
    for (int tile = 0; tile < my_tiles; ++tile) {
      int4 jd = __ldg(p.jobs);
      for (int j = 0; j < njobs; ++j) {
        const int4 cur = jd;
        if (j + 1 < njobs) jd = __ldg(p.jobs + j + 1);
        if (tile > 0 || j >= MI_STAGES) mbar_wait(&empty[stage], par ^ 1);
        const int phase = cur.x, kch = cur.z;
        const int rows = phase == 2 ? 2 : (cur.w ? 16 : 8);
        const uint32_t row_bytes = kch * 32;
        if (kch == 0) {
          if (lane == 0) mbar_arrive(&full[stage]);
        } else {
          if (lane == 0) mbar_expect_tx(&full[stage], rows * row_bytes);
          __syncwarp();
          if (lane < rows) {
            const __nv_bfloat16* src;
            if (phase == 0) src = p.B1 + static_cast<size_t>(cur.y + lane) * Dp;
            else if (phase == 1) src = p.B2 + static_cast<size_t>(cur.y + lane) * H;
            else src = p.B3 + static_cast<size_t>(cur.y + lane * D) * H;     // rows d (mu) and D + d (alpha)
            mi_bulk_row(ring_s + stage * stage_bytes + lane * ldb_bytes, src, row_bytes, &full[stage]);
          }
        }
        if (++stage == MI_STAGES) { stage = 0; par ^= 1; }
      }
    }
    return;
  }

  // ---------------- consumers: warp w owns samples [base, base + R) of the CTA tile
  const int g = lane >> 2, t = lane & 3;
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(act + static_cast<size_t>(warp) * per_warp);
  __nv_bfloat16* h1 = xb + R * ldx;
  __nv_bfloat16* h2 = h1 + R * ldh;
  const int lrow = lane & 15, lcol = (lane >> 4) * 8;
  const uint32_t xb_lane = smem_u32(xb + lrow * ldx + lcol);
  const uint32_t h1_lane = smem_u32(h1 + lrow * ldh + lcol);
  const uint32_t h2_lane = smem_u32(h2 + lrow * ldh + lcol);
  // B-operand ldmatrix address of this lane inside a stage: matrices = (tile 0, k 0-7), (tile 0, k 8-15), (tile 1, ..)
  const uint32_t b_lane = ((lane & 7) + ((lane >> 4) << 3)) * ldb_bytes + ((lane >> 3) & 1) * 16;

  int stage = 0, par = 0;
  for (int tile = 0; tile < my_tiles; ++tile) {
    const long long base = ((static_cast<long long>(tile) * gridDim.x + blockIdx.x) * cwarps + warp) * R;
    float ldacc[MT][2], unext[MT][2];
    {
      // clear this warp's slice: scratch columns must be finite (they meet masked-zero weights)
      uint4* z = reinterpret_cast<uint4*>(xb);
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      for (int i = lane; i < per_warp / 16; i += 32) z[i] = zero;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const long long b = base + mt * 16 + g + 8 * hh;
          ldacc[mt][hh] = 0.f;
          unext[mt][hh] = (t == 0 && b < p.B) ? __ldg(p.u_in + b * D + (p.flip ? D - 1 : 0)) : 0.f;
        }
      __syncwarp();
    }
    int4 jd = __ldg(p.jobs);
    for (int j = 0; j < njobs; ++j) {
      const int4 cur = jd;
      if (j + 1 < njobs) jd = __ldg(p.jobs + j + 1);         // next descriptor requested a job ahead
      const int phase = cur.x, kch = cur.z;
      const uint32_t bs = ring_s + stage * stage_bytes + b_lane;

      if (phase < 2) {
        // out[:, row0 .. row0 + 16) = bf16(relu(in[:, 0 .. 16*kch) . W^T + bias))
        const bool l1 = phase == 0, two = cur.w != 0;
        const uint32_t in_lane = l1 ? xb_lane : h1_lane;
        const int lda_bytes = (l1 ? ldx : ldh) * 2;
        const float* bias = (l1 ? p.b1 : p.b2) + cur.y + 2 * t;
        const float bv00 = __ldg(bias), bv01 = __ldg(bias + 1);
        const float bv10 = two ? __ldg(bias + 8) : 0.f, bv11 = two ? __ldg(bias + 9) : 0.f;
        float acc[MT][2][2][4];   // [m-tile][n-tile][even / odd k-chunk chain]
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[mt][n][c][e] = 0.f;
        mbar_wait(&full[stage], par);
        int kc = 0;
#pragma unroll 2
        for (; kc + 1 < kch; kc += 2) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t b[4];
            mi_ldsm_x4(bs + (kc + c) * 32, b);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t a[4];
              mi_ldsm_x4(in_lane + mt * 16 * lda_bytes + (kc + c) * 32, a);
              mi_mma(acc[mt][0][c], a, b[0], b[1]);
              mi_mma(acc[mt][1][c], a, b[2], b[3]);
            }
          }
        }
        if (kc < kch) {
          uint32_t b[4];
          mi_ldsm_x4(bs + kc * 32, b);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t a[4];
            mi_ldsm_x4(in_lane + mt * 16 * lda_bytes + kc * 32, a);
            mi_mma(acc[mt][0][0], a, b[0], b[1]);
            mi_mma(acc[mt][1][0], a, b[2], b[3]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);             // weights consumed: the stage may be refilled
        __nv_bfloat16* out = (l1 ? h1 : h2) + cur.y + 2 * t;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          *reinterpret_cast<uint32_t*>(out + (mt * 16 + g) * ldh) = mi_pack_relu(
              acc[mt][0][0][0] + acc[mt][0][1][0] + bv00, acc[mt][0][0][1] + acc[mt][0][1][1] + bv01);
          *reinterpret_cast<uint32_t*>(out + (mt * 16 + g + 8) * ldh) = mi_pack_relu(
              acc[mt][0][0][2] + acc[mt][0][1][2] + bv00, acc[mt][0][0][3] + acc[mt][0][1][3] + bv01);
          if (two) {
            *reinterpret_cast<uint32_t*>(out + (mt * 16 + g) * ldh + 8) = mi_pack_relu(
                acc[mt][1][0][0] + acc[mt][1][1][0] + bv10, acc[mt][1][0][1] + acc[mt][1][1][1] + bv11);
            *reinterpret_cast<uint32_t*>(out + (mt * 16 + g + 8) * ldh + 8) = mi_pack_relu(
                acc[mt][1][0][2] + acc[mt][1][1][2] + bv10, acc[mt][1][0][3] + acc[mt][1][1][3] + bv11);
          }
        }
        __syncwarp();
      } else {
        // (mu_d, alpha_d) from the layer-2 units of degree <= d; stage row 0 = mu weights, row 1 = alpha weights
        const int d = cur.y;
        float uv[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uv[mt][hh] = unext[mt][hh];
            const long long b = base + mt * 16 + g + 8 * hh;   // next step's u, requested a whole step ahead
            unext[mt][hh] =
                (t == 0 && b < p.B && d + 1 < D) ? __ldg(p.u_in + b * D + (p.flip ? D - 2 - d : d + 1)) : 0.f;
          }
        const float bm = __ldg(p.b3 + d), ba = __ldg(p.b3 + D + d);
        float acc[MT][4][4];      // four independent k-chunk chains
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][c][e] = 0.f;
        mbar_wait(&full[stage], par);
        int kc = 0;
        for (; kc + 3 < kch; kc += 4) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t b[4];
            mi_ldsm_x4(bs + (kc + c) * 32, b);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t a[4];
              mi_ldsm_x4(h2_lane + mt * 16 * ldh * 2 + (kc + c) * 32, a);
              mi_mma(acc[mt][c], a, b[0], b[1]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (kc + c < kch) {
            uint32_t b[4];
            mi_ldsm_x4(bs + (kc + c) * 32, b);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t a[4];
              mi_ldsm_x4(h2_lane + mt * 16 * ldh * 2 + (kc + c) * 32, a);
              mi_mma(acc[mt][c], a, b[0], b[1]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (t == 0) {   // lanes t == 0 hold columns 0 (mu) and 1 (alpha) of samples g and g + 8
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float mu = (acc[mt][0][2 * hh] + acc[mt][1][2 * hh]) + (acc[mt][2][2 * hh] + acc[mt][3][2 * hh]) + bm;
              const float al = (acc[mt][0][2 * hh + 1] + acc[mt][1][2 * hh + 1]) +
                               (acc[mt][2][2 * hh + 1] + acc[mt][3][2 * hh + 1]) + ba;
              const float xv = uv[mt][hh] * expf(al) + mu;
              const int s = mt * 16 + g + 8 * hh;
              if (base + s < p.B) p.x[(base + s) * D + d] = xv;
              xb[s * ldx + d] = __float2bfloat16_rn(xv);
              ldacc[mt][hh] += al;
            }
        }
        __syncwarp();
      }
      if (++stage == MI_STAGES) { stage = 0; par ^= 1; }
    }
    if (t == 0 && p.ld_out) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const long long b = base + mt * 16 + g + 8 * hh;
          if (b < p.B) p.ld_out[b] = (p.ld_in ? p.ld_in[b] : 0.f) + ldacc[mt][hh];
        }
    }
  }
}

static inline int mi_per_warp_bytes(int mt, int H, int Dp) { return mt * 16 * ((Dp + 8) + 2 * (H + 8)) * 2; }
static inline int mi_fixed_bytes(int H, int Dp) { return MI_STAGES * 16 * ((H > Dp ? H : Dp) + 8) * 2 + 128; }

}  // namespace nfk

using namespace nfk;

static constexpr int MI_SMEM_MAX = 227 * 1024;

extern "C" int nfk_made_inverse_resident_supported(int D, int H, int Dp) {
  if (D <= 0 || H <= 0 || H % 64 || Dp % 64 || Dp < D) return 0;
  return mi_fixed_bytes(H, Dp) + mi_per_warp_bytes(1, H, Dp) <= MI_SMEM_MAX ? 1 : 0;
}

// Host-side: the job stream of one sample tile from the degree counts (cnt[d] = units with degree <= d, d = 0..D).
extern "C" int nfk_made_inverse_jobs(const int* cnt1, const int* cnt2, int D, int* jobs, int cap) {
  if (!cnt1 || !cnt2 || D <= 0 || cap < 0 || (cap > 0 && !jobs)) return NFK_ERR_ARG;
  int n = 0;
  auto put = [&](int phase, int row0, int kch, int two) {
    if (n < cap) { jobs[4 * n] = phase; jobs[4 * n + 1] = row0; jobs[4 * n + 2] = kch; jobs[4 * n + 3] = two; }
    ++n;
  };
  for (int d = 0; d < D; ++d) {
    const int c1p = d ? cnt1[d - 1] : 0, c1 = cnt1[d], c2p = d ? cnt2[d - 1] : 0, c2 = cnt2[d];
    if (c1 < c1p || c2 < c2p || c1p < 0 || c2p < 0) return NFK_ERR_ARG;
    if (d > 0) {
      if (c1 > c1p)   // layer-1 units of degree d: inputs x_0 .. x_{d-1}
        for (int nt = c1p >> 3, hi = (c1 + 7) >> 3; nt < hi; nt += 2) put(0, nt * 8, (d + 15) >> 4, nt + 1 < hi);
      if (c2 > c2p)   // layer-2 units of degree d: layer-1 units of degree <= d
        for (int nt = c2p >> 3, hi = (c2 + 7) >> 3; nt < hi; nt += 2) put(1, nt * 8, (c1 + 15) >> 4, nt + 1 < hi);
    }
    put(2, d, (c2 + 15) >> 4, 0);   // (mu_d, alpha_d): layer-2 units of degree <= d
  }
  return n;
}

template <int MT>
static int mi_launch(const MiArgs& p, cudaStream_t st) {
  const int per_warp = mi_per_warp_bytes(MT, p.H, p.Dp), fixed = mi_fixed_bytes(p.H, p.Dp);
  int warps = (MI_SMEM_MAX - fixed) / per_warp;
  if (warps > 8) warps = 8;
  if (warps < 1) return NFK_ERR_SHAPE;
  const int wtiles = (p.B + MT * 16 - 1) / (MT * 16);
  if (warps > wtiles) warps = wtiles;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = fixed + warps * per_warp;
  if (cudaFuncSetAttribute(made_inverse_resident_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
      cudaSuccess)
    return NFK_ERR_LAUNCH;
  int grid = (wtiles + warps - 1) / warps;
  if (grid > sms) grid = sms;
  made_inverse_resident_kernel<MT><<<grid, (warps + 1) * 32, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_made_inverse_resident(const float* u_in, const void* B1, const void* B2, const void* B3,
                                         const float* b1, const float* b2, const float* b3, const int* jobs,
                                         int njobs, float* x, const float* ld_in, float* ld_out, int B, int D, int H,
                                         int Dp, int flip, int mtiles, void* stream) {
  if (B <= 0 || njobs <= 0 || !nfk_made_inverse_resident_supported(D, H, Dp) || mtiles < 0 || mtiles > 2)
    return NFK_ERR_SHAPE;
  if (!u_in || !B1 || !B2 || !B3 || !b1 || !b2 || !b3 || !jobs || !x) return NFK_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(jobs) & 15) return NFK_ERR_ARG;
  MiArgs p;
  p.u_in = u_in;
  p.B1 = static_cast<const __nv_bfloat16*>(B1);
  p.B2 = static_cast<const __nv_bfloat16*>(B2);
  p.B3 = static_cast<const __nv_bfloat16*>(B3);
  p.b1 = b1; p.b2 = b2; p.b3 = b3;
  p.jobs = reinterpret_cast<const int4*>(jobs); p.njobs = njobs;
  p.x = x; p.ld_in = ld_in; p.ld_out = ld_out;
  p.B = B; p.D = D; p.H = H; p.Dp = Dp; p.flip = flip;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // one 16-sample tile per warp leaves room for the most warps (latency hiding); two halve the B-operand reads
  int mt = mtiles == 0 ? 1 : mtiles;
  if (mt == 2 && mi_fixed_bytes(H, Dp) + mi_per_warp_bytes(2, H, Dp) > MI_SMEM_MAX) mt = 1;
  return mt == 2 ? mi_launch<2>(p, st) : mi_launch<1>(p, st);
}
