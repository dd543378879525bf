// Split2d (reference: models/layers.py:293-313), the Gaussian prior / bits-per-dim objective
// (models/layers.py:10-23, models/kd_flows.py:134-150) and the multi-level latent MSE of the KD loss
// (pl_module.py:266-282). fp32, HBM-bound reductions with one per-sample result.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>

#include "../../include/nfk.h"
#include "launch_util.h"

namespace nfk {

constexpr int ST = 256;
constexpr float kLog2Pi = 1.8378770664093453f;

struct SGeo {
  int B, C, HW, W, H, ipc, pixt;
  // SqueezeLayer (models/layers.py:32-44,316-327) folded into the split: the level's output z1 is ALSO written in
  // space-to-depth layout [B, 4*C/2, H/2, W/2] (forward), and the gradient of that squeezed tensor is read through the
  // same index map (backward) — the next level's input costs no separate permute/copy pass.
  float* z1_sq;
  const float* g_z1_sq;
};

// element (b, c, pixel p) of a [B, CH, H, W] map inside its 2x2 space-to-depth image [B, 4*CH, H/2, W/2]:
// channel c*4 + (y&1)*2 + (x&1), position (y/2, x/2)
__device__ __forceinline__ long long squeezed_index(int b, int c, int p, int CH, int H, int W) {
  const int y = p / W, x = p - y * W;
  return ((static_cast<long long>(b) * 4 * CH + c * 4 + (y & 1) * 2 + (x & 1)) * (H >> 1) + (y >> 1)) * (W >> 1) +
         (x >> 1);
}

// smem: ws[C*CH*9] | bsc[2*C] (bias, exp(3 logs)) | z1s[CH*ldp] | z2s[CH*ldp] | ls[CH*ldp]
__global__ void __launch_bounds__(ST)
split2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ logs, float* __restrict__ z1_out, const float* __restrict__ eps,
                   float temperature, float* __restrict__ out_full, float* __restrict__ ld, SGeo g, int reverse) {
  extern __shared__ float sm[];
  const int C = g.C, CH = C / 2, ldp = g.pixt + 1;
  float* ws = sm;
  float* bsc = ws + C * CH * 9;
  float* z1s = bsc + 2 * C;
  float* z2s = z1s + CH * ldp;
  float* ls = z2s + CH * ldp;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * g.ipc;
  const int nimg = min(g.ipc, g.B - b0);
  const int npix = nimg * g.HW;
  const int cin_total = reverse ? CH : C;  // channels of the input tensor x

  for (int i = tid; i < C * CH * 9; i += ST) ws[i] = w[i];
  for (int i = tid; i < C; i += ST) { bsc[i] = bias[i]; bsc[C + i] = expf(3.f * logs[i]); }
  for (int i = tid; i < CH * npix; i += ST) {
    const int img = i / (CH * g.HW), r = i - img * CH * g.HW;
    const int c = r / g.HW, p = r - c * g.HW;
    const long long base = (static_cast<long long>(b0 + img) * cin_total) * g.HW;
    const float v1 = x[base + static_cast<long long>(c) * g.HW + p];
    z1s[c * ldp + img * g.HW + p] = v1;
    if (!reverse) {
      z2s[c * ldp + img * g.HW + p] = x[base + static_cast<long long>(CH + c) * g.HW + p];
      z1_out[(static_cast<long long>(b0 + img) * CH + c) * g.HW + p] = v1;
      if (g.z1_sq) g.z1_sq[squeezed_index(b0 + img, c, p, CH, g.H, g.W)] = v1;
    } else {
      out_full[(static_cast<long long>(b0 + img) * C + c) * g.HW + p] = v1;
      z2s[c * ldp + img * g.HW + p] =
          eps ? eps[(static_cast<long long>(b0 + img) * CH + c) * g.HW + p] : 0.f;
    }
  }
  __syncthreads();
  const int PPP = ST / CH;
  const int j = tid % CH, pl0 = tid / CH;
  if (pl0 < PPP) {
    for (int pl = pl0; pl < npix; pl += PPP) {
      const int img = pl / g.HW, rem = pl - img * g.HW;
      const int yy = rem / g.W, xx = rem - yy * g.W;
      float mean = 0.f, lsg = 0.f;
      for (int tap = 0; tap < 9; ++tap) {
        const int ny = yy + tap / 3 - 1, nx = xx + tap % 3 - 1;
        if (ny < 0 || ny >= g.H || nx < 0 || nx >= g.W) continue;
        const int q = img * g.HW + ny * g.W + nx;
        const float* wm = ws + (2 * j) * CH * 9 + tap;
        const float* wl = wm + CH * 9;
        for (int ci = 0; ci < CH; ++ci) {
          const float v = z1s[ci * ldp + q];
          mean = fmaf(wm[ci * 9], v, mean);
          lsg = fmaf(wl[ci * 9], v, lsg);
        }
      }
      mean = (mean + bsc[2 * j]) * bsc[C + 2 * j];
      lsg = (lsg + bsc[2 * j + 1]) * bsc[C + 2 * j + 1];
      const float z2 = z2s[j * ldp + pl];
      if (!reverse) {
        const float d = z2 - mean;
        ls[j * ldp + pl] = -0.5f * (2.f * lsg + d * d * expf(-2.f * lsg) + kLog2Pi);
      } else {
        z2s[j * ldp + pl] = mean + expf(lsg) * temperature * z2;  // z2 held eps
      }
    }
  }
  __syncthreads();
  if (!reverse) {
    const int warp = tid >> 5, lane = tid & 31;
    for (int img = warp; img < nimg; img += ST / 32) {
      float acc = 0.f;
      for (int i = lane; i < CH * g.HW; i += 32) {
        const int jj = i / g.HW, p = i - jj * g.HW;
        acc += ls[jj * ldp + img * g.HW + p];
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0 && ld) ld[b0 + img] += acc;
    }
  } else {
    for (int i = tid; i < CH * npix; i += ST) {
      const int img = i / (CH * g.HW), r = i - img * CH * g.HW;
      const int c = r / g.HW, p = r - c * g.HW;
      out_full[(static_cast<long long>(b0 + img) * C + CH + c) * g.HW + p] = z2s[c * ldp + img * g.HW + p];
    }
  }
}

// Persistent CTAs over groups of whole images; weight / bias / logs gradients are accumulated in shared memory across
// all groups of a CTA (each thread owns its weight entries: no atomics) and flushed with one global atomic per entry.
// smem: ws[C*CH*9] | bsc[2C] | z1s[CH*ldp] | dps[C*ldp] | accb[2C] | accw[C*CH*9]
__global__ void __launch_bounds__(ST)
split2d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ logs, const float* __restrict__ g_z1, const float* __restrict__ g_ld,
                   float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias,
                   float* __restrict__ dlogs, SGeo g) {
  extern __shared__ float sm[];
  const int C = g.C, CH = C / 2, ldp = g.pixt + 1;
  float* ws = sm;
  float* bsc = ws + C * CH * 9;
  float* z1s = bsc + 2 * C;
  float* dps = z1s + CH * ldp;
  float* accb = dps + C * ldp;
  float* accw = accb + 2 * C;
  const int tid = threadIdx.x;
  const int ngroups = (g.B + g.ipc - 1) / g.ipc;

  for (int i = tid; i < C * CH * 9; i += ST) { ws[i] = w[i]; accw[i] = 0.f; }
  for (int i = tid; i < C; i += ST) { bsc[i] = bias[i]; bsc[C + i] = expf(3.f * logs[i]); }
  for (int i = tid; i < 2 * C; i += ST) accb[i] = 0.f;

  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b0 = grp * g.ipc;
    const int nimg = min(g.ipc, g.B - b0);
    const int npix = nimg * g.HW;
    __syncthreads();
    for (int i = tid; i < CH * npix; i += ST) {
      const int img = i / (CH * g.HW), r = i - img * CH * g.HW;
      const int c = r / g.HW, p = r - c * g.HW;
      z1s[c * ldp + img * g.HW + p] = x[(static_cast<long long>(b0 + img) * C + c) * g.HW + p];
    }
    __syncthreads();
    const int PPP = ST / CH;
    const int j = tid % CH, pl0 = tid / CH;
    if (pl0 < PPP) {
      float a_bm = 0.f, a_bl = 0.f, a_lm = 0.f, a_ll = 0.f;
      for (int pl = pl0; pl < npix; pl += PPP) {
        const int img = pl / g.HW, rem = pl - img * g.HW;
        const int yy = rem / g.W, xx = rem - yy * g.W;
        float mean = 0.f, lsg = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
          const int ny = yy + tap / 3 - 1, nx = xx + tap % 3 - 1;
          if (ny < 0 || ny >= g.H || nx < 0 || nx >= g.W) continue;
          const int q = img * g.HW + ny * g.W + nx;
          const float* wm = ws + (2 * j) * CH * 9 + tap;
          const float* wl = wm + CH * 9;
          for (int ci = 0; ci < CH; ++ci) {
            const float v = z1s[ci * ldp + q];
            mean = fmaf(wm[ci * 9], v, mean);
            lsg = fmaf(wl[ci * 9], v, lsg);
          }
        }
        const float em = bsc[C + 2 * j], el = bsc[C + 2 * j + 1];
        mean = (mean + bsc[2 * j]) * em;
        lsg = (lsg + bsc[2 * j + 1]) * el;
        const long long zi = (static_cast<long long>(b0 + img) * C + CH + j) * g.HW + rem;
        const float z2 = x[zi];
        const float gl = g_ld[b0 + img];
        const float d = z2 - mean;
        const float r = d * expf(-2.f * lsg);
        const float dmean = gl * r;
        const float dlsg = gl * (d * r - 1.f);
        dx[zi] = -gl * r;
        const float dpm = dmean * em, dpl = dlsg * el;
        dps[(2 * j) * ldp + pl] = dpm;
        dps[(2 * j + 1) * ldp + pl] = dpl;
        a_bm += dpm; a_bl += dpl;
        a_lm += dmean * mean; a_ll += dlsg * lsg;
      }
      atomicAdd(&accb[2 * j], a_bm);
      atomicAdd(&accb[2 * j + 1], a_bl);
      atomicAdd(&accb[C + 2 * j], 3.f * a_lm);
      atomicAdd(&accb[C + 2 * j + 1], 3.f * a_ll);
    }
    __syncthreads();
    // dz1[ci, m] = g_z1 + sum_{co,tap} dpre[co, m - off(tap)] * w[co][ci][tap]
    if (pl0 < PPP) {
      const int ci = j;
      for (int pl = pl0; pl < npix; pl += PPP) {
        const int img = pl / g.HW, rem = pl - img * g.HW;
        const int yy = rem / g.W, xx = rem - yy * g.W;
        float a = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
          const int ny = yy - (tap / 3 - 1), nx = xx - (tap % 3 - 1);
          if (ny < 0 || ny >= g.H || nx < 0 || nx >= g.W) continue;
          const int q = img * g.HW + ny * g.W + nx;
          const float* wp = ws + ci * 9 + tap;
          for (int co = 0; co < C; ++co) a = fmaf(dps[co * ldp + q], wp[co * CH * 9], a);
        }
        const long long gi = (static_cast<long long>(b0 + img) * CH + ci) * g.HW + rem;
        dx[(static_cast<long long>(b0 + img) * C + ci) * g.HW + rem] =
            a + (g_z1 ? g_z1[gi] : 0.f) +
            (g.g_z1_sq ? g.g_z1_sq[squeezed_index(b0 + img, ci, rem, CH, g.H, g.W)] : 0.f);
      }
    }
    // dw[co][ci][tap] += sum_m dpre[co, m] * z1[ci, m + off(tap)]   (entry e is owned by thread e % ST)
    for (int e = tid; e < C * CH * 9; e += ST) {
      const int co = e / (CH * 9), r = e - co * CH * 9;
      const int ci = r / 9, tap = r - ci * 9;
      const int dyy = tap / 3 - 1, dxx = tap % 3 - 1;
      float a = 0.f;
      for (int img = 0; img < nimg; ++img) {
        const int y0 = max(0, -dyy), y1 = min(g.H, g.H - dyy);
        const int x0 = max(0, -dxx), x1 = min(g.W, g.W - dxx);
        for (int yy = y0; yy < y1; ++yy)
          for (int xx = x0; xx < x1; ++xx)
            a = fmaf(dps[co * ldp + img * g.HW + yy * g.W + xx],
                     z1s[ci * ldp + img * g.HW + (yy + dyy) * g.W + xx + dxx], a);
      }
      accw[e] += a;
    }
  }
  __syncthreads();
  for (int i = tid; i < C; i += ST) { atomicAdd(dbias + i, accb[i]); atomicAdd(dlogs + i, accb[C + i]); }
  for (int e = tid; e < C * CH * 9; e += ST) atomicAdd(dw + e, accw[e]);
}

// ---- Split2d for the channel counts the shipped configs use (C = 12, 24, 48, 96): thread = pixel -----------------
// The 3x3 conv C/2 -> C is computed per pixel with all C outputs in registers; weights sit in shared memory as
// wt[tap][ci][co] (co contiguous: broadcast LDS.128), z1 / dpre tiles feature-major (a warp's pixels are consecutive
// addresses). CTA = `ipc` whole images (pixt pixels), as in the generic kernels above.
template <int CH>
__device__ __forceinline__ void split_conv_pixel(const float* __restrict__ wt, const float* __restrict__ z1s, int ldp,
                                                 int img_base, int yy, int xx, int H, int W, float (&acc)[2 * CH]) {
  constexpr int C = 2 * CH;
#pragma unroll
  for (int o = 0; o < C; ++o) acc[o] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int ny = yy + tap / 3 - 1, nx = xx + tap % 3 - 1;
    if (ny < 0 || ny >= H || nx < 0 || nx >= W) continue;
    const float* zp = z1s + img_base + ny * W + nx;
    const float* wp = wt + tap * CH * C;
#pragma unroll 2
    for (int ci = 0; ci < CH; ++ci) {
      const float v = zp[ci * ldp];
      const float4* w4 = reinterpret_cast<const float4*>(wp + ci * C);
#pragma unroll
      for (int o4 = 0; o4 < C / 4; ++o4) {
        const float4 w = w4[o4];
        acc[4 * o4] = fmaf(w.x, v, acc[4 * o4]);
        acc[4 * o4 + 1] = fmaf(w.y, v, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(w.z, v, acc[4 * o4 + 2]);
        acc[4 * o4 + 3] = fmaf(w.w, v, acc[4 * o4 + 3]);
      }
    }
  }
}

// smem: wt[9*CH*C] | bsc[2C] | z1s[CH*ldp] | red[ST/32 + 1]
template <int CH>
__global__ void __launch_bounds__(ST)
split2d_fwd_px_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ logs, float* __restrict__ z1_out, const float* __restrict__ eps,
                      float temperature, float* __restrict__ out_full, float* __restrict__ ld, SGeo g, int reverse) {
  constexpr int C = 2 * CH;
  extern __shared__ __align__(16) float sm[];
  const int ldp = g.pixt + 1;
  float* wt = sm;
  float* bsc = wt + 9 * CH * C;
  float* z1s = bsc + 2 * C;
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * g.ipc;
  const int nimg = min(g.ipc, g.B - b0);
  const int npix = nimg * g.HW;
  const int cin_total = reverse ? CH : C;
  for (int i = tid; i < C * CH * 9; i += ST) {   // w[co][ci][tap] -> wt[tap][ci][co]
    const int co = i / (CH * 9), r = i - co * CH * 9, ci = r / 9, tap = r - ci * 9;
    wt[(tap * CH + ci) * C + co] = w[i];
  }
  for (int i = tid; i < C; i += ST) { bsc[i] = bias[i]; bsc[C + i] = expf(3.f * logs[i]); }
  for (int i = tid; i < CH * npix; i += ST) {   // pixel fastest: coalesced
    const int c = i / npix, pl = i - c * npix;
    const int img = pl / g.HW, p = pl - img * g.HW;
    const float v1 = x[(static_cast<long long>(b0 + img) * cin_total + c) * g.HW + p];
    z1s[c * ldp + pl] = v1;
    if (!reverse) {
      z1_out[(static_cast<long long>(b0 + img) * CH + c) * g.HW + p] = v1;
      if (g.z1_sq) g.z1_sq[squeezed_index(b0 + img, c, p, CH, g.H, g.W)] = v1;
    } else {
      out_full[(static_cast<long long>(b0 + img) * C + c) * g.HW + p] = v1;
    }
  }
  __syncthreads();
  for (int pl0 = 0; pl0 < npix; pl0 += ST) {
    const int pl = pl0 + tid;
    float lsum = 0.f;
    int img = 0;
    if (pl < npix) {
      img = pl / g.HW;
      const int rem = pl - img * g.HW;
      const int yy = rem / g.W, xx = rem - yy * g.W;
      float acc[C];
      split_conv_pixel<CH>(wt, z1s, ldp, img * g.HW, yy, xx, g.H, g.W, acc);
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float mean = (acc[2 * j] + bsc[2 * j]) * bsc[C + 2 * j];
        const float lsg = (acc[2 * j + 1] + bsc[2 * j + 1]) * bsc[C + 2 * j + 1];
        if (!reverse) {
          const float z2 = x[(static_cast<long long>(b0 + img) * C + CH + j) * g.HW + rem];
          const float d = z2 - mean;
          lsum += -0.5f * (2.f * lsg + d * d * expf(-2.f * lsg) + kLog2Pi);
        } else {
          const long long ei = (static_cast<long long>(b0 + img) * CH + j) * g.HW + rem;
          const float e = eps ? eps[ei] : 0.f;
          out_full[(static_cast<long long>(b0 + img) * C + CH + j) * g.HW + rem] = mean + expf(lsg) * temperature * e;
        }
      }
    }
    if (!reverse && ld) {
      // lanes = consecutive pixels: segments of min(32, HW) lanes belong to one image
      const int lane = tid & 31, seg = g.HW < 32 ? g.HW : 32;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v = __shfl_xor_sync(0xffffffffu, lsum, o);
        if (o < seg) lsum += v;
      }
      if ((lane & (seg - 1)) == 0 && pl < npix) atomicAdd(ld + b0 + img, lsum);
    }
  }
}

// Backward. smem: wt[9*CH*C] | wt2[9*C*CH] | bsc[2C] | z1s[CH*ldp] | dps[C*ldp] | qs[C*ldp] | accb[2C]
// Weight gradients: a thread owns 4 output channels x one (ci, tap) and keeps its sums in registers across all image
// groups of the CTA (one global atomic per entry at the end).
template <int CH>
__global__ void __launch_bounds__(ST)
split2d_bwd_px_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      const float* __restrict__ logs, const float* __restrict__ g_z1, const float* __restrict__ g_ld,
                      float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias,
                      float* __restrict__ dlogs, SGeo g) {
  constexpr int C = 2 * CH;
  constexpr int NT = (C / 4) * CH * 9;            // weight-gradient thread tiles
  constexpr int NPASS = (NT + ST - 1) / ST;
  extern __shared__ __align__(16) float sm[];
  const int ldp = g.pixt + 1;
  float* wt = sm;                       // [tap][ci][co]
  float* wt2 = wt + 9 * CH * C;         // [tap][co][ci]
  float* bsc = wt2 + 9 * C * CH;
  float* z1s = bsc + 2 * C;
  float* dps = z1s + CH * ldp;
  float* qs = dps + C * ldp;
  float* accb = qs + C * ldp;
  const int tid = threadIdx.x;
  const int ngroups = (g.B + g.ipc - 1) / g.ipc;
  for (int i = tid; i < C * CH * 9; i += ST) {
    const int co = i / (CH * 9), r = i - co * CH * 9, ci = r / 9, tap = r - ci * 9;
    const float v = w[i];
    wt[(tap * CH + ci) * C + co] = v;
    wt2[(tap * C + co) * CH + ci] = v;
  }
  for (int i = tid; i < C; i += ST) { bsc[i] = bias[i]; bsc[C + i] = expf(3.f * logs[i]); }
  for (int i = tid; i < 2 * C; i += ST) accb[i] = 0.f;
  float wacc[NPASS][4];
#pragma unroll
  for (int p = 0; p < NPASS; ++p) { wacc[p][0] = wacc[p][1] = wacc[p][2] = wacc[p][3] = 0.f; }

  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const int b0 = grp * g.ipc;
    const int nimg = min(g.ipc, g.B - b0);
    const int npix = nimg * g.HW;
    __syncthreads();
    for (int i = tid; i < CH * npix; i += ST) {
      const int c = i / npix, pl = i - c * npix;
      const int img = pl / g.HW, p = pl - img * g.HW;
      z1s[c * ldp + pl] = x[(static_cast<long long>(b0 + img) * C + c) * g.HW + p];
    }
    __syncthreads();
    // ---- A: per pixel, recompute (mean, logs), gradient wrt the conv outputs -> dps, dlogs integrand -> qs, dz2
    for (int pl = tid; pl < npix; pl += ST) {
      const int img = pl / g.HW, rem = pl - img * g.HW;
      const int yy = rem / g.W, xx = rem - yy * g.W;
      float acc[C];
      split_conv_pixel<CH>(wt, z1s, ldp, img * g.HW, yy, xx, g.H, g.W, acc);
      const float gl = g_ld[b0 + img];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float em = bsc[C + 2 * j], el = bsc[C + 2 * j + 1];
        const float mean = (acc[2 * j] + bsc[2 * j]) * em;
        const float lsg = (acc[2 * j + 1] + bsc[2 * j + 1]) * el;
        const long long zi = (static_cast<long long>(b0 + img) * C + CH + j) * g.HW + rem;
        const float d = x[zi] - mean;
        const float r = d * expf(-2.f * lsg);
        const float dmean = gl * r;
        const float dlsg = gl * (d * r - 1.f);
        dx[zi] = -gl * r;
        dps[(2 * j) * ldp + pl] = dmean * em;
        dps[(2 * j + 1) * ldp + pl] = dlsg * el;
        qs[(2 * j) * ldp + pl] = dmean * mean;
        qs[(2 * j + 1) * ldp + pl] = dlsg * lsg;
      }
    }
    __syncthreads();
    // ---- bias / logs gradients: one warp per output channel sums its dps / qs rows
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int co = warp; co < C; co += ST / 32) {
        float a = 0.f, q = 0.f;
        for (int p = lane; p < npix; p += 32) { a += dps[co * ldp + p]; q += qs[co * ldp + p]; }
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) { accb[co] += a; accb[C + co] += 3.f * q; }
      }
    }
    // ---- B: dz1[ci, m] = g_z1 + sum_{tap,co} dpre[co, m - off(tap)] * w[co][ci][tap]
    for (int pl = tid; pl < npix; pl += ST) {
      const int img = pl / g.HW, rem = pl - img * g.HW;
      const int yy = rem / g.W, xx = rem - yy * g.W;
      float a[CH];
#pragma unroll
      for (int ci = 0; ci < CH; ++ci) a[ci] = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int ny = yy - (tap / 3 - 1), nx = xx - (tap % 3 - 1);
        if (ny < 0 || ny >= g.H || nx < 0 || nx >= g.W) continue;
        const float* dp = dps + img * g.HW + ny * g.W + nx;
        const float* wp = wt2 + tap * C * CH;
#pragma unroll 2
        for (int co = 0; co < C; ++co) {
          const float v = dp[co * ldp];
          if (CH % 4 == 0) {
            const float4* w4 = reinterpret_cast<const float4*>(wp + co * CH);
#pragma unroll
            for (int c4 = 0; c4 < CH / 4; ++c4) {
              const float4 ww = w4[c4];
              a[4 * c4] = fmaf(ww.x, v, a[4 * c4]); a[4 * c4 + 1] = fmaf(ww.y, v, a[4 * c4 + 1]);
              a[4 * c4 + 2] = fmaf(ww.z, v, a[4 * c4 + 2]); a[4 * c4 + 3] = fmaf(ww.w, v, a[4 * c4 + 3]);
            }
          } else {
            const float2* w2 = reinterpret_cast<const float2*>(wp + co * CH);
#pragma unroll
            for (int c2 = 0; c2 < CH / 2; ++c2) {
              const float2 ww = w2[c2];
              a[2 * c2] = fmaf(ww.x, v, a[2 * c2]); a[2 * c2 + 1] = fmaf(ww.y, v, a[2 * c2 + 1]);
            }
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < CH; ++ci) {
        const long long gi = (static_cast<long long>(b0 + img) * CH + ci) * g.HW + rem;
        dx[(static_cast<long long>(b0 + img) * C + ci) * g.HW + rem] =
            a[ci] + (g_z1 ? g_z1[gi] : 0.f) +
            (g.g_z1_sq ? g.g_z1_sq[squeezed_index(b0 + img, ci, rem, CH, g.H, g.W)] : 0.f);
      }
    }
    // ---- C: dw[co][ci][tap] += sum_m dpre[co, m] * z1[ci, m + off(tap)]; thread tile = 4 co x one (ci, tap)
#pragma unroll
    for (int p = 0; p < NPASS; ++p) {
      const int t = tid + p * ST;
      if (t < NT) {
        const int cb = t / (CH * 9), r = t - cb * CH * 9;      // (ci, tap) fastest: lanes share the 4 dps rows
        const int ci = r / 9, tap = r - ci * 9;
        const int dyy = tap / 3 - 1, dxx = tap % 3 - 1;
        const int y0 = max(0, -dyy), y1 = min(g.H, g.H - dyy);
        const int x0 = max(0, -dxx), x1 = min(g.W, g.W - dxx);
        const float* d0 = dps + (4 * cb) * ldp;
        const float* zc = z1s + ci * ldp + dyy * g.W + dxx;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        for (int img = 0; img < nimg; ++img) {
          for (int yy = y0; yy < y1; ++yy) {
            const int rowb = img * g.HW + yy * g.W;
            for (int xx = x0; xx < x1; ++xx) {
              const float z = zc[rowb + xx];
              const float* dq = d0 + rowb + xx;
              s0 = fmaf(dq[0], z, s0);
              s1 = fmaf(dq[ldp], z, s1);
              s2 = fmaf(dq[2 * ldp], z, s2);
              s3 = fmaf(dq[3 * ldp], z, s3);
            }
          }
        }
        wacc[p][0] += s0; wacc[p][1] += s1; wacc[p][2] += s2; wacc[p][3] += s3;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < C; i += ST) { atomicAdd(dbias + i, accb[i]); atomicAdd(dlogs + i, accb[C + i]); }
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    const int t = tid + p * ST;
    if (t < NT) {
      const int cb = t / (CH * 9), r = t - cb * CH * 9;
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(dw + (4 * cb + k) * CH * 9 + r, wacc[p][k]);
    }
  }
}

// One warp per sample: out[b] = -(logdet[b] + sum_i logN(z_i; mean_i, exp(logs_i))) * scale
__global__ void prior_bpd_fwd_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                     const float* __restrict__ logs, const float* __restrict__ logdet, int B, int n,
                                     float scale, float* __restrict__ out) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* zp = z + static_cast<long long>(b) * n;
  float acc = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float l = logs[i], d = zp[i] - mean[i];
    acc += -0.5f * (2.f * l + d * d * expf(-2.f * l) + kLog2Pi);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[b] = -(logdet[b] + acc) * scale;
}

__global__ void prior_bpd_bwd_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                     const float* __restrict__ logs, const float* __restrict__ g_bpd, int B, int n,
                                     float scale, float* __restrict__ dz, float* __restrict__ dlogdet) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * n) return;
  const int b = static_cast<int>(i / n), k = static_cast<int>(i - static_cast<long long>(b) * n);
  const float gs = g_bpd[b] * scale;
  dz[i] = gs * (z[i] - mean[k]) * expf(-2.f * logs[k]);
  if (k == 0) dlogdet[b] = -gs;
}

// acc[b] += scale * sum_i (s - t)^2 ; GROUP threads per sample
template <int GROUP>
__global__ void kd_mse_fwd_kernel(const float* __restrict__ s, const float* __restrict__ t, int B, int n, float scale,
                                  float* __restrict__ acc) {
  __shared__ float red[ST / 32];
  const int per_cta = ST / GROUP;
  const int b = blockIdx.x * per_cta + threadIdx.x / GROUP;
  const int l = threadIdx.x % GROUP;
  float a = 0.f;
  if (b < B) {
    const float* sp = s + static_cast<long long>(b) * n;
    const float* tp = t + static_cast<long long>(b) * n;
    if ((n & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(sp);
      const float4* t4 = reinterpret_cast<const float4*>(tp);
      for (int i = l; i < n / 4; i += GROUP) {
        const float4 u = __ldg(s4 + i), v = __ldg(t4 + i);
        const float d0 = u.x - v.x, d1 = u.y - v.y, d2 = u.z - v.z, d3 = u.w - v.w;
        a += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
    } else {
      for (int i = l; i < n; i += GROUP) {
        const float d = sp[i] - tp[i];
        a += d * d;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (GROUP == 32) {
    if (l == 0 && b < B) acc[b] += a * scale;
  } else {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0 && b < B) {
      float tot = 0.f;
      for (int i = 0; i < ST / 32; ++i) tot += red[i];
      acc[b] += tot * scale;
    }
  }
}

__global__ void kd_mse_bwd_kernel(const float* __restrict__ s, const float* __restrict__ t,
                                  const float* __restrict__ g, long long total, int n, float scale2,
                                  float* __restrict__ ds, int accumulate) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const float v = scale2 * g[i / n] * (s[i] - t[i]);
  ds[i] = accumulate ? ds[i] + v : v;
}

static SGeo make_sgeo(int B, int C, int H, int W, int target) {
  SGeo g;
  g.z1_sq = nullptr; g.g_z1_sq = nullptr;
  g.B = B; g.C = C; g.H = H; g.W = W; g.HW = H * W;
  g.ipc = g.HW >= target ? 1 : target / g.HW;
  if (g.ipc > B) g.ipc = B;
  while (g.ipc > 1 && (B + g.ipc - 1) / g.ipc < 2 * 148) g.ipc >>= 1;
  g.pixt = g.ipc * g.HW;
  return g;
}

// Thread-per-pixel kernels for C in {12, 24, 48}; returns 1 when the shape is not covered (caller falls back).
template <int CH>
static int split_fwd_px_t(const float* x, const float* w, const float* bias, const float* logs, float* z1_out,
                          const float* eps, float temperature, float* out_full, float* ld, const SGeo& g, int reverse,
                          void* stream) {
  constexpr int C = 2 * CH;
  const int smem = (9 * CH * C + 2 * C + CH * (g.pixt + 1) + 8) * 4;
  if (smem > 200 * 1024) return 1;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(split2d_fwd_px_kernel<CH>), smem)) return rc;
  split2d_fwd_px_kernel<CH><<<(g.B + g.ipc - 1) / g.ipc, ST, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, logs, z1_out, eps, temperature, out_full, ld, g, reverse);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
static int split_fwd_px(const float* x, const float* w, const float* bias, const float* logs, float* z1_out,
                        const float* eps, float temperature, float* out_full, float* ld, const SGeo& g, int reverse,
                        void* stream) {
  if (const char* e = getenv("NFK_SPLIT_GENERIC"); e && e[0] == '1') return 1;
  switch (g.C) {
    case 12: return split_fwd_px_t<6>(x, w, bias, logs, z1_out, eps, temperature, out_full, ld, g, reverse, stream);
    case 24: return split_fwd_px_t<12>(x, w, bias, logs, z1_out, eps, temperature, out_full, ld, g, reverse, stream);
    case 48: return split_fwd_px_t<24>(x, w, bias, logs, z1_out, eps, temperature, out_full, ld, g, reverse, stream);
    default: return 1;
  }
}
template <int CH>
static int split_bwd_px_t(const float* x, const float* w, const float* bias, const float* logs, const float* g_z1,
                          const float* g_ld, float* dx, float* dw, float* dbias, float* dlogs, const SGeo& g,
                          void* stream) {
  constexpr int C = 2 * CH;
  const int smem = (2 * 9 * CH * C + 2 * C + (CH + 2 * C) * (g.pixt + 1) + 2 * C + 8) * 4;
  if (smem > 200 * 1024) return 1;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(split2d_bwd_px_kernel<CH>), smem)) return rc;
  const int groups = (g.B + g.ipc - 1) / g.ipc;
  split2d_bwd_px_kernel<CH><<<groups < 4 * 148 ? groups : 4 * 148, ST, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, logs, g_z1, g_ld, dx, dw, dbias, dlogs, g);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
static int split_bwd_px(const float* x, const float* w, const float* bias, const float* logs, const float* g_z1,
                        const float* g_ld, float* dx, float* dw, float* dbias, float* dlogs, const SGeo& g,
                        void* stream) {
  if (const char* e = getenv("NFK_SPLIT_GENERIC"); e && e[0] == '1') return 1;
  switch (g.C) {
    case 12: return split_bwd_px_t<6>(x, w, bias, logs, g_z1, g_ld, dx, dw, dbias, dlogs, g, stream);
    case 24: return split_bwd_px_t<12>(x, w, bias, logs, g_z1, g_ld, dx, dw, dbias, dlogs, g, stream);
    case 48: return split_bwd_px_t<24>(x, w, bias, logs, g_z1, g_ld, dx, dw, dbias, dlogs, g, stream);
    default: return 1;
  }
}

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_split2d_fwd(const float* x, const float* w, const float* bias, const float* logs, float* z1_out,
                               float* ld, int B, int C, int H, int W, void* stream) {
  return nfk_split2d_squeeze_fwd(x, w, bias, logs, z1_out, nullptr, ld, B, C, H, W, stream);
}

extern "C" int nfk_split2d_squeeze_fwd(const float* x, const float* w, const float* bias, const float* logs,
                                       float* z1_out, float* z1_sq_out, float* ld, int B, int C, int H, int W,
                                       void* stream) {
  if (B <= 0 || C <= 0 || C % 2 || C > 128 || H * W > 4096) return NFK_ERR_SHAPE;
  if (z1_sq_out && ((H | W) & 1)) return NFK_ERR_SHAPE;
  if (!x || !w || !bias || !logs || !z1_out) return NFK_ERR_ARG;
  SGeo g = make_sgeo(B, C, H, W, 256);
  g.z1_sq = z1_sq_out;
  if (int rc = split_fwd_px(x, w, bias, logs, z1_out, nullptr, 0.f, nullptr, ld, g, 0, stream); rc != 1) return rc;
  const int smem = (C * (C / 2) * 9 + 2 * C + 3 * (C / 2) * (g.pixt + 1)) * 4;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(split2d_fwd_kernel), smem)) return rc;
  split2d_fwd_kernel<<<(B + g.ipc - 1) / g.ipc, ST, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, logs, z1_out, nullptr, 0.f, nullptr, ld, g, 0);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_split2d_rev(const float* z1, const float* w, const float* bias, const float* logs,
                               const float* eps, float temperature, float* out, int B, int C, int H, int W,
                               void* stream) {
  if (B <= 0 || C <= 0 || C % 2 || C > 128 || H * W > 4096) return NFK_ERR_SHAPE;
  if (!z1 || !w || !bias || !logs || !out) return NFK_ERR_ARG;
  SGeo g = make_sgeo(B, C, H, W, 256);
  if (int rc = split_fwd_px(z1, w, bias, logs, nullptr, eps, temperature, out, nullptr, g, 1, stream); rc != 1) return rc;
  const int smem = (C * (C / 2) * 9 + 2 * C + 3 * (C / 2) * (g.pixt + 1)) * 4;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(split2d_fwd_kernel), smem)) return rc;
  split2d_fwd_kernel<<<(B + g.ipc - 1) / g.ipc, ST, smem, static_cast<cudaStream_t>(stream)>>>(
      z1, w, bias, logs, nullptr, eps, temperature, out, nullptr, g, 1);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_split2d_bwd(const float* x, const float* w, const float* bias, const float* logs,
                               const float* g_z1, const float* g_ld, float* dx, float* dw, float* dbias,
                               float* dlogs, int B, int C, int H, int W, void* stream) {
  return nfk_split2d_squeeze_bwd(x, w, bias, logs, g_z1, nullptr, g_ld, dx, dw, dbias, dlogs, B, C, H, W, stream);
}

extern "C" int nfk_split2d_squeeze_bwd(const float* x, const float* w, const float* bias, const float* logs,
                                       const float* g_z1, const float* g_z1_sq, const float* g_ld, float* dx,
                                       float* dw, float* dbias, float* dlogs, int B, int C, int H, int W,
                                       void* stream) {
  if (B <= 0 || C <= 0 || C % 2 || C > 128 || H * W > 4096) return NFK_ERR_SHAPE;
  if (g_z1_sq && ((H | W) & 1)) return NFK_ERR_SHAPE;
  if (!x || !w || !bias || !logs || !g_ld || !dx || !dw || !dbias || !dlogs) return NFK_ERR_ARG;
  SGeo g = make_sgeo(B, C, H, W, 128);
  g.g_z1_sq = g_z1_sq;
  if (int rc = split_bwd_px(x, w, bias, logs, g_z1, g_ld, dx, dw, dbias, dlogs, g, stream); rc != 1) return rc;
  const int smem = (2 * C * (C / 2) * 9 + 2 * C + (C / 2 + C) * (g.pixt + 1) + 2 * C) * 4;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(split2d_bwd_kernel), smem)) return rc;
  const int groups = (B + g.ipc - 1) / g.ipc;
  split2d_bwd_kernel<<<groups < 4 * 148 ? groups : 4 * 148, ST, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, logs, g_z1, g_ld, dx, dw, dbias, dlogs, g);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_prior_bpd_fwd(const float* z, const float* mean, const float* logs, const float* logdet, int B,
                                 int n, float scale, float* out, void* stream) {
  if (B <= 0 || n <= 0) return NFK_ERR_SHAPE;
  if (!z || !mean || !logs || !logdet || !out) return NFK_ERR_ARG;
  const long long threads = static_cast<long long>(B) * 32;
  prior_bpd_fwd_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, mean, logs, logdet, B, n, scale, out);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_prior_bpd_bwd(const float* z, const float* mean, const float* logs, const float* g_bpd, int B,
                                 int n, float scale, float* dz, float* dlogdet, void* stream) {
  if (B <= 0 || n <= 0) return NFK_ERR_SHAPE;
  if (!z || !mean || !logs || !g_bpd || !dz || !dlogdet) return NFK_ERR_ARG;
  const long long total = static_cast<long long>(B) * n;
  prior_bpd_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, mean, logs, g_bpd, B, n, scale, dz, dlogdet);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_kd_mse_fwd(const float* s, const float* t, int B, int n, float scale, float* acc, void* stream) {
  if (B <= 0 || n <= 0) return NFK_ERR_SHAPE;
  if (!s || !t || !acc) return NFK_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(t)) & 15) return NFK_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n >= 1024) kd_mse_fwd_kernel<256><<<B, ST, 0, st>>>(s, t, B, n, scale, acc);
  else kd_mse_fwd_kernel<32><<<(B + 7) / 8, ST, 0, st>>>(s, t, B, n, scale, acc);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_kd_mse_bwd(const float* s, const float* t, const float* g, int B, int n, float scale, float* ds,
                              int accumulate, void* stream) {
  if (B <= 0 || n <= 0) return NFK_ERR_SHAPE;
  if (!s || !t || !g || !ds) return NFK_ERR_ARG;
  const long long total = static_cast<long long>(B) * n;
  kd_mse_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, t, g, total, n, 2.f * scale, ds, accumulate);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
