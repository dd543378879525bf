// tcgen05 / TMEM / TMA GEMM tiles for the coupling network (reference: models/flows.py:25-34 get_block_2d,
// models/layers.py:190-260 Conv2d / Conv2dZeros). Activations are pixel-major ("NHWC") bf16 matrices
// [pixels, channels]; every convolution of the coupling net (and its dgrad / wgrad) is one of two GEMM forms:
//
//   NT:  out[M, N]   = A[M, K] * B[N, K]^T          (forward convs, dgrads; both operands K-major)
//   TN:  out[Mo, No] += sum_k A[k, Mo] * B[k, No]   (wgrads; both operands MN-major, split-K over pixels)
//
// One CTA = one 128 x BN accumulator tile living in TMEM. Warp 0 streams 128-byte-swizzled operand tiles with
// TMA into a multi-stage smem ring, one thread of warp 1 issues tcgen05.mma, warps 2-5 drain TMEM through the
// fused epilogue (bias+ReLU->bf16, ReLU-mask->bf16 + column sums, fp32 store, or split-K red.add).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"
#include "launch_util.h"
#include "ptx.cuh"

namespace nfk {

constexpr int BM = 128;           // accumulator rows per CTA (= TMEM lanes)
constexpr int BK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;        // K per tcgen05.mma for 16-bit operands
constexpr int GEMM_THREADS = 192; // warp0 TMA, warp1 MMA(+TMEM alloc), warps 2-5 epilogue

struct GemmArgs {
  int M, N, K;         // NT: out[M,N]; TN: out[M=Mo, N=No], K = pixels
  int BN;              // accumulator tile width (multiple of 16, <= 256)
  int stages;          // smem ring depth
  int n_tiles;         // tiles along N
  int kb_per_split;    // TN: k-blocks per split
  void* out;
  long long ldo;
  const float* bias;   // EPI_BIAS_RELU_BF16 / EPI_F32 (optional)
  const __nv_bfloat16* aux;  // EPI_MASK_BF16: post-ReLU activation of the layer being differentiated
  long long ldaux;
  float* colsum;       // EPI_MASK_BF16: per-column sum of the masked gradient (bias gradient)
  // MADE masked linears: k-blocks [kb_begin, kb_end) that are not structurally zero for each n-tile (kb_end 0 = all)
  short kb_begin[16];
  short kb_end[16];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ uint32_t tmem_cols_pow2(int n) {
  uint32_t c = 32;
  while (c < static_cast<uint32_t>(n)) c <<= 1;
  return c;
}

// ------------------------------------------------------------------------------------------------ NT kernel
// Persistent: grid = min(tiles, SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ... (n-tile fastest, so CTAs
// running side by side share the A row block through L2). Two TMEM accumulator stages let the epilogue of tile i
// overlap the MMAs of tile i+1.
template <int EPI>
__device__ __forceinline__ void nt_epilogue_16(const GemmArgs& g, const uint32_t (&r)[16], long long row, bool row_ok,
                                               int col, int cl, const float* bias_s, int lane, uint4 m0, uint4 m1) {
  if constexpr (EPI == NFK_EPI_F32) {
    if (row_ok) {
      float* o = static_cast<float*>(g.out) + row * g.ldo + col;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 v;
        v.x = __uint_as_float(r[j + 0]); v.y = __uint_as_float(r[j + 1]);
        v.z = __uint_as_float(r[j + 2]); v.w = __uint_as_float(r[j + 3]);
        if (g.bias) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + cl + j);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        *reinterpret_cast<float4*>(o + j) = v;
      }
    }
  } else if constexpr (EPI == NFK_EPI_BIAS_RELU_BF16) {
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + cl + j);
      p[(j >> 1)] = pack_bf16x2(fmaxf(__uint_as_float(r[j]) + b.x, 0.f), fmaxf(__uint_as_float(r[j + 1]) + b.y, 0.f));
      p[(j >> 1) + 1] =
          pack_bf16x2(fmaxf(__uint_as_float(r[j + 2]) + b.z, 0.f), fmaxf(__uint_as_float(r[j + 3]) + b.w, 0.f));
    }
    if (row_ok) {
      uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(g.out) + row * g.ldo + col);
      o[0] = make_uint4(p[0], p[1], p[2], p[3]);
      o[1] = make_uint4(p[4], p[5], p[6], p[7]);
    }
  } else {  // NFK_EPI_MASK_BF16: ReLU backward mask + bias-gradient column sums
    float v[16];
    const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // post-ReLU activations are >= 0, so "active" <=> the bf16 bit pattern is non-zero (and not -0).
      const uint32_t h = (j & 1) ? (mw[j >> 1] >> 16) : (mw[j >> 1] & 0xFFFFu);
      const bool on = row_ok && h != 0u && h != 0x8000u;
      v[j] = on ? __uint_as_float(r[j]) : 0.f;
    }
    if (row_ok) {
      uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(g.out) + row * g.ldo + col);
      o[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      o[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
    }
    if (g.colsum) {
      // 32 lanes x 16 columns -> lane j (< 16) ends with the sum of column j over the warp's 32 rows.
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
          const float send = up ? v[j] : v[j + off];
          const float keep = up ? v[j + off] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      if (lane < 16) atomicAdd(g.colsum + col + lane, v[0]);
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = BM * 128;
  const int b_bytes = g.BN * 128;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + g.stages * stage_bytes);
  uint64_t* empty = full + g.stages;
  uint64_t* tmem_full = empty + g.stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][256] (one per accumulator stage)

  const int num_kb = g.K / BK;
  const int m_tiles = (g.M + BM - 1) / BM;
  const int num_tiles = m_tiles * g.n_tiles;
  const uint32_t ncols = tmem_cols_pow2(2 * g.BN);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ncols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_tile = t % g.n_tiles, m_tile = t / g.n_tiles;
        const int kb0 = g.kb_begin[n_tile & 15], kb1 = g.kb_end[n_tile & 15] ? g.kb_end[n_tile & 15] : num_kb;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          mbar_expect_tx(&full[s], stage_bytes);
          tma_load_2d(sa, &tmA, &full[s], kb * BK, m_tile * BM);
          tma_load_2d(sa + a_bytes, &tmB, &full[s], kb * BK, n_tile * g.BN);
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, g.BN, false, false);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * g.BN;
        const int n_tile = t % g.n_tiles;
        const int kb0 = g.kb_begin[n_tile & 15], kb1 = g.kb_end[n_tile & 15] ? g.kb_end[n_tile & 15] : num_kb;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = umma_desc_sw128(a_addr + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bd = umma_desc_sw128(b_addr + k * (UMMA_K * 2), 16, 1024);
            umma_f16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - 64;  // 0..127
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int n_tile = t % g.n_tiles, m_tile = t / g.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      float* bs = bias_s + acc * 256;
      if (EPI != NFK_EPI_MASK_BF16 && g.bias) {
        // stage this tile's bias slice; the named barrier also orders it against the previous use of this stage
        for (int c = et; c < g.BN; c += 128) {
          const int col = n_tile * g.BN + c;
          bs[c] = col < g.N ? __ldg(g.bias + col) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const long long row = static_cast<long long>(m_tile) * BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      uint4 ma[4], mb[4];
      auto load_mask = [&](int c, uint4 (&mk)[4]) {
        const int col = n_tile * g.BN + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mk[j] = make_uint4(0, 0, 0, 0);
          if (row_ok && c + 8 * j < g.BN && col + 8 * j < g.N)
            mk[j] = __ldg(reinterpret_cast<const uint4*>(g.aux + row * g.ldaux + col + 8 * j));
        }
      };
      if constexpr (EPI == NFK_EPI_MASK_BF16) {
        load_mask(0, ma);
        load_mask(32, mb);
      }
      mbar_wait(&tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * g.BN;
      auto do32 = [&](int c, const uint4 (&mk)[4]) {
        uint32_t r0[16], r1[16];
        const bool two = c + 16 < g.BN;
        tmem_ld16(taddr + c, r0);
        if (two) tmem_ld16(taddr + c + 16, r1);
        tmem_ld_wait();
        const int col = n_tile * g.BN + c;
        if (col < g.N) nt_epilogue_16<EPI>(g, r0, row, row_ok, col, c, bs, lane, mk[0], mk[1]);
        if (two && col + 16 < g.N) nt_epilogue_16<EPI>(g, r1, row, row_ok, col + 16, c + 16, bs, lane, mk[2], mk[3]);
      };
      if constexpr (EPI == NFK_EPI_MASK_BF16) {
        // the ReLU mask comes from HBM with a row stride of ldaux: keep two 32-column groups of it in flight
        // (issued before the accumulator is even complete) so the epilogue never waits on a cold load
        for (int c = 0; c < g.BN; c += 64) {
          do32(c, ma);
          if (c + 64 < g.BN) load_mask(c + 64, ma);
          if (c + 32 < g.BN) {
            do32(c + 32, mb);
            if (c + 96 < g.BN) load_mask(c + 96, mb);
          }
        }
      } else {
        const uint4 none[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0),
                               make_uint4(0, 0, 0, 0)};
        for (int c = 0; c < g.BN; c += 32) do32(c, none);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------ TN kernel
// out[Mo, No] (fp32, pre-zeroed) += sum over this CTA's pixel range of A[k, m] * B[k, n].
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BOX = 64 * 128;  // one TMA box: 64 pixel rows x 64 channels (128 B)
  const int a_bytes = 2 * BOX;
  const int nb_boxes = g.BN / 64;
  const int stage_bytes = a_bytes + nb_boxes * BOX;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + g.stages * stage_bytes);
  uint64_t* empty = full + g.stages;
  uint64_t* tmem_full = empty + g.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int n_tile = blockIdx.x % g.n_tiles;
  const int m_tile = blockIdx.x / g.n_tiles;
  const int total_kb = (g.K + BK - 1) / BK;
  const int kb0 = blockIdx.y * g.kb_per_split;
  const int kb1 = min(kb0 + g.kb_per_split, total_kb);
  const int num_kb = kb1 - kb0;
  if (num_kb <= 0) return;
  const uint32_t ncols = tmem_cols_pow2(g.BN);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ncols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * stage_bytes;
        mbar_expect_tx(&full[s], stage_bytes);
        tma_load_2d(sa, &tmA, &full[s], m_tile * BM, kb * BK);
        tma_load_2d(sa + BOX, &tmA, &full[s], m_tile * BM + 64, kb * BK);
        for (int j = 0; j < nb_boxes; ++j)
          tma_load_2d(sa + a_bytes + j * BOX, &tmB, &full[s], n_tile * g.BN + j * 64, kb * BK);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, g.BN, true, true);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // 16 pixel rows per MMA = two 8-row swizzle groups (SBO = 1024 B); 64-channel blocks are BOX apart (LBO).
          const uint64_t ad = umma_desc_sw128(a_addr + k * (UMMA_K * 128), BOX, 1024);
          const uint64_t bd = umma_desc_sw128(b_addr + k * (UMMA_K * 128), BOX, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const long long row = static_cast<long long>(m_tile) * BM + q * 32 + lane;
    const bool row_ok = row < g.M;
    for (int c = 0; c < g.BN; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, r);
      tmem_ld_wait();
      const int col = n_tile * g.BN + c;
      if (col >= g.N) break;
      if (row_ok) {
        float* o = static_cast<float*>(g.out) + row * g.ldo + col;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(__uint_as_float(r[j])),
                       "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                       : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = 64 columns x box_rows rows.
static int make_tmap_bf16(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld,
                          uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return NFK_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16) return NFK_ERR_ALIGN;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? NFK_OK : NFK_ERR_DRIVER;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int pick_bn(int N, int cap) {
  // Widest tile (multiple of 16, <= cap) that wastes the fewest padded columns.
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= cap) return n16;
  int best = cap, best_waste = 1 << 30;
  for (int bn = cap; bn >= 128; bn -= 16) {
    const int tiles = (n16 + bn - 1) / bn;
    const int waste = tiles * bn - n16;
    if (waste < best_waste) { best = bn; best_waste = waste; }
  }
  return best;
}

template <typename K>
static int set_smem(K kernel, int bytes) {
  return ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes);
}

}  // namespace nfk

using namespace nfk;

static int gemm_nt_launch(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int epi,
                          void* out, long long ldo, const float* bias, const void* aux, long long ldaux,
                          float* colsum, int bn_force, const int* kb_begin, const int* kb_end, void* stream);

extern "C" int nfk_gemm_nt_bf16(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                                int epi, void* out, long long ldo, const float* bias, const void* aux,
                                long long ldaux, float* colsum, void* stream) {
  return gemm_nt_launch(A, lda, B, ldb, M, N, K, epi, out, ldo, bias, aux, ldaux, colsum, 0, nullptr, nullptr,
                        stream);
}

extern "C" int nfk_gemm_nt_bf16_ranged(const void* A, long long lda, const void* B, long long ldb, int M, int N,
                                       int K, int epi, void* out, long long ldo, const float* bias, const void* aux,
                                       long long ldaux, float* colsum, int bn, const int* kb_begin,
                                       const int* kb_end, void* stream) {
  if (bn <= 0 || bn % 16 || bn > 256 || (N + bn - 1) / bn > 16 || !kb_begin || !kb_end) return NFK_ERR_ARG;
  return gemm_nt_launch(A, lda, B, ldb, M, N, K, epi, out, ldo, bias, aux, ldaux, colsum, bn, kb_begin, kb_end,
                        stream);
}

static int gemm_nt_launch(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int epi,
                          void* out, long long ldo, const float* bias, const void* aux, long long ldaux,
                          float* colsum, int bn_force, const int* kb_begin, const int* kb_end, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return NFK_ERR_SHAPE;
  if (K % BK || N % 16 || ldo % 8) return NFK_ERR_SHAPE;
  if (epi == NFK_EPI_BIAS_RELU_BF16 && !bias) return NFK_ERR_ARG;
  if (epi == NFK_EPI_MASK_BF16 && (!aux || ldaux % 8)) return NFK_ERR_ARG;
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K;
  g.BN = pick_bn(N, 256);
  // few row tiles (small images x small batch): narrower accumulator tiles so the grid still covers the SMs
  const int m_tiles = (M + BM - 1) / BM;
  while (m_tiles * ((N + g.BN - 1) / g.BN) < 148 && g.BN >= 128 && g.BN % 32 == 0) g.BN /= 2;
  if (bn_force) g.BN = bn_force;
  g.n_tiles = (N + g.BN - 1) / g.BN;
  if (kb_begin)
    for (int i = 0; i < g.n_tiles && i < 16; ++i) {
      if (kb_begin[i] < 0 || kb_end[i] > K / BK || kb_begin[i] >= kb_end[i]) return NFK_ERR_ARG;
      g.kb_begin[i] = static_cast<short>(kb_begin[i]);
      g.kb_end[i] = static_cast<short>(kb_end[i]);
    }
  const int stage_bytes = BM * 128 + g.BN * 128;
  g.stages = min(8, (194 * 1024) / stage_bytes);
  g.out = out; g.ldo = ldo; g.bias = bias;
  g.aux = static_cast<const __nv_bfloat16*>(aux); g.ldaux = ldaux; g.colsum = colsum;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_bf16(&tmA, A, K, M, lda, BM)) != NFK_OK) return rc;
  if ((rc = make_tmap_bf16(&tmB, B, K, N, ldb, g.BN)) != NFK_OK) return rc;
  const int smem = g.stages * stage_bytes + 1024 + 256 + 2 * 256 * 4;
  const int tiles = m_tiles * g.n_tiles;
  const dim3 grid(tiles < num_sms() ? tiles : num_sms());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (epi) {
    case NFK_EPI_F32:
      if ((rc = set_smem(gemm_nt_kernel<NFK_EPI_F32>, smem))) return rc;
      gemm_nt_kernel<NFK_EPI_F32><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, g);
      break;
    case NFK_EPI_BIAS_RELU_BF16:
      if ((rc = set_smem(gemm_nt_kernel<NFK_EPI_BIAS_RELU_BF16>, smem))) return rc;
      gemm_nt_kernel<NFK_EPI_BIAS_RELU_BF16><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, g);
      break;
    case NFK_EPI_MASK_BF16:
      if ((rc = set_smem(gemm_nt_kernel<NFK_EPI_MASK_BF16>, smem))) return rc;
      gemm_nt_kernel<NFK_EPI_MASK_BF16><<<grid, GEMM_THREADS, smem, st>>>(tmA, tmB, g);
      break;
    default:
      return NFK_ERR_ARG;
  }
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_gemm_tn_bf16(const void* A, long long lda, const void* B, long long ldb, int Mo, int No, int Kpix,
                                float* out, long long ldo, int sm_count, void* stream) {
  if (Mo <= 0 || No <= 0 || Kpix <= 0) return NFK_ERR_SHAPE;
  if (No % 64 || ldo % 4 || Mo % 8) return NFK_ERR_SHAPE;
  GemmArgs g{};
  g.M = Mo; g.N = No; g.K = Kpix;
  g.BN = No <= 256 ? No : (No % 256 == 0 ? 256 : (No % 192 == 0 ? 192 : (No % 128 == 0 ? 128 : 64)));
  g.n_tiles = (No + g.BN - 1) / g.BN;
  const int stage_bytes = 2 * 64 * 128 + (g.BN / 64) * 64 * 128;
  g.stages = min(8, (196 * 1024) / stage_bytes);
  g.out = out; g.ldo = ldo;
  const int tiles = ((Mo + BM - 1) / BM) * g.n_tiles;
  const int total_kb = (Kpix + BK - 1) / BK;
  if (sm_count <= 0) sm_count = 148;
  int splits = max(1, min(total_kb, (sm_count + tiles - 1) / tiles));
  g.kb_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + g.kb_per_split - 1) / g.kb_per_split;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_bf16(&tmA, A, Mo, Kpix, lda, 64)) != NFK_OK) return rc;
  if ((rc = make_tmap_bf16(&tmB, B, No, Kpix, ldb, 64)) != NFK_OK) return rc;
  const int smem = g.stages * stage_bytes + 1024 + 256;
  if ((rc = set_smem(gemm_tn_kernel, smem))) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  gemm_tn_kernel<<<dim3(tiles, splits), GEMM_THREADS, smem, st>>>(tmA, tmB, g);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
