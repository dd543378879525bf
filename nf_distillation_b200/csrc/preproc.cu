// First kernel of the 2-D flow: image pre-processing + dequantisation noise + the first SqueezeLayer in ONE pass
// (reference: data/src/utils.py:7-18 preprocess, models/utils.py:26-41 uniform_binning_correction,
//  models/layers.py:32-44 squeeze2d — three full passes over the batch in the reference, one here).
//
//   v            = src (fp32, already in [-0.5, 0.5))                       when the batch arrives as floats, or
//                  floor(u8 / 2^(8 - n_bits)) / 2^n_bits - 0.5              when it arrives as raw uint8 pixels
//   x_out        = v + noise                      (fp32 [B, C, H, W]; may alias src: the reference adds the noise to
//                                                  the caller's batch in place and the teacher then noises it again)
//   sq_out       = space-to-depth of x_out        (fp32 [B, 4C, H/2, W/2], channel c*4 + (y&1)*2 + (x&1))
// The noise itself stays a torch RNG draw (same Philox stream as the reference's torch.zeros_like(x).uniform_()).
// HBM-bound: per element 4 (or 1) B read + 4 B noise + 8 B written; a thread moves four consecutive pixels of a row
// with 128-bit loads and two 64-bit squeezed stores.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"

namespace nfk {

__global__ void dequant_squeeze_kernel(const float* __restrict__ src_f, const uint8_t* __restrict__ src_u8, int shift,
                                       float inv_bins, const float* __restrict__ noise, float* x_out,
                                       float* __restrict__ sq_out, long long quads, int C, int H, int W) {
  const long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (q >= quads) return;
  const int wq = W >> 2;
  const int xq = static_cast<int>(q % wq);
  const long long r = q / wq;                   // (b * C + c) * H + y
  const int y = static_cast<int>(r % H);
  const long long bc = r / H;
  const long long e0 = r * W + 4 * xq;
  float4 v;
  if (src_u8) {
    const uchar4 u = *reinterpret_cast<const uchar4*>(src_u8 + e0);
    v = make_float4((u.x >> shift) * inv_bins - 0.5f, (u.y >> shift) * inv_bins - 0.5f,
                    (u.z >> shift) * inv_bins - 0.5f, (u.w >> shift) * inv_bins - 0.5f);
  } else {
    v = *reinterpret_cast<const float4*>(src_f + e0);
  }
  if (noise) {
    const float4 n = __ldg(reinterpret_cast<const float4*>(noise + e0));
    v.x += n.x; v.y += n.y; v.z += n.z; v.w += n.w;
  }
  if (x_out) *reinterpret_cast<float4*>(x_out + e0) = v;
  if (sq_out) {
    const int c = static_cast<int>(bc % C);
    const long long b = bc / C;
    const int H2 = H >> 1, W2 = W >> 1;
    // channel c*4 + (y&1)*2 + fw, position (y/2, x/2): pixels 0,2 of the quad -> fw = 0, pixels 1,3 -> fw = 1
    float* o = sq_out + (((b * 4 * C + c * 4 + (y & 1) * 2) * H2 + (y >> 1)) * W2 + 2 * xq);
    *reinterpret_cast<float2*>(o) = make_float2(v.x, v.z);
    *reinterpret_cast<float2*>(o + static_cast<long long>(H2) * W2) = make_float2(v.y, v.w);
  }
}

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_dequant_squeeze(const void* src, int src_is_u8, int n_bits, const float* noise, float* x_out,
                                   float* sq_out, int B, int C, int H, int W, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (W & 3) || (H & 1) || n_bits < 1 || n_bits > 8) return NFK_ERR_SHAPE;
  if (!src || (!x_out && !sq_out)) return NFK_ERR_ARG;
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(noise) |
                       reinterpret_cast<uintptr_t>(x_out);
  if ((al & (src_is_u8 ? 3 : 15)) || (reinterpret_cast<uintptr_t>(noise) & 15) ||
      (reinterpret_cast<uintptr_t>(x_out) & 15) || (reinterpret_cast<uintptr_t>(sq_out) & 7))
    return NFK_ERR_ALIGN;
  const long long quads = static_cast<long long>(B) * C * H * (W >> 2);
  const unsigned grid = static_cast<unsigned>((quads + 255) / 256);
  dequant_squeeze_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src_is_u8 ? nullptr : static_cast<const float*>(src), src_is_u8 ? static_cast<const uint8_t*>(src) : nullptr,
      8 - n_bits, 1.f / static_cast<float>(1 << n_bits), noise, x_out, sq_out, quads, C, H, W);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
