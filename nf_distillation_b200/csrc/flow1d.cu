// 1-D (tabular) Glow FlowStep, fused end to end in one kernel per direction (reference: the is_1d branches of
// models/flows.py:37-52,142-202 and models/layers.py:76,117,410-411).
//
// A persistent CTA (one per SM) walks tiles of NT samples with every weight of the step resident in shared memory.
// Activations live in shared memory feature-major ([feature][NT+4]); each layer is a small fp32 GEMM over the tile,
// register-blocked 8 outputs x 4 samples per thread (3 LDS.128 per 32 FMAs):
//   y = W' x + b'            (ActNorm1d + x @ W folded by nfk_invconv_prep with transpose=1)
//   h = MLP(y1 [, cond])     Linear-ReLU x4, Linear-Tanh, Linear
//   y2 = (y2 + h[0::2]) * sigmoid(h[1::2] + 2),  logdet += sum log sigmoid          (or the inverse order / formulas)
// The backward kernel reloads the saved MLP activations, walks the chain in reverse with the same tile GEMM for the
// data gradients and an 8x4 register-blocked outer-product GEMM (reduction over the samples) for the weight gradients,
// which accumulate per CTA in shared memory and are added to the global gradient block once.
// fp32 SIMT on purpose: these layers are 16..64 wide and the path's parity bar is 1e-4 on log-det in fp32.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"
#include "launch_util.h"

namespace nfk {

constexpr int F1_LAYERS = 7;  // 0: fused affine, 1..6: coupling MLP

struct F1Dims {
  int D, D1, D2, Cc, hid;
  int nin[F1_LAYERS], nout[F1_LAYERS], ninp[F1_LAYERS], noutp[F1_LAYERS];
  int offWT[F1_LAYERS], offB[F1_LAYERS];  // forward block: WT_l [nin][noutp], bias_l [noutp]
  int offW[F1_LAYERS];                    // backward block: W_l [nout][ninp]
  int offG[F1_LAYERS], offGB[F1_LAYERS];  // gradient block: dW_l [nout][ninp], db_l [noutp]
  int total_fwd, total_bwd, total_grad;
  int n_act;                              // saved activations per sample: H1..H5 (5*hid) + O (2*D2)
};

static inline int up8(int x) { return (x + 7) / 8 * 8; }
__host__ __device__ static inline int up8d(int x) { return (x + 7) / 8 * 8; }

static F1Dims f1_dims(int D, int Cc, int hid) {
  F1Dims d{};
  d.D = D; d.D1 = D / 2; d.D2 = D - D / 2; d.Cc = Cc; d.hid = hid;
  const int nin[F1_LAYERS] = {D, d.D1 + Cc, hid, hid, hid, hid, hid};
  const int nout[F1_LAYERS] = {D, hid, hid, hid, hid, hid, 2 * d.D2};
  int of = 0, ob = 0, og = 0;
  for (int l = 0; l < F1_LAYERS; ++l) {
    d.nin[l] = nin[l]; d.nout[l] = nout[l]; d.ninp[l] = up8(nin[l]); d.noutp[l] = up8(nout[l]);
    d.offWT[l] = of; of += nin[l] * d.noutp[l];
    d.offB[l] = of; of += d.noutp[l];
    d.offW[l] = ob; ob += nout[l] * d.ninp[l];
    d.offG[l] = og; og += nout[l] * d.ninp[l];
    d.offGB[l] = og; og += d.noutp[l];
  }
  d.total_fwd = of; d.total_bwd = ob; d.total_grad = og;
  d.n_act = 5 * hid + 2 * d.D2;
  return d;
}

struct F1Weights {
  const float* Wf;     // [D][D] fused affine, out-layout
  const float* bf;     // [D]
  const float* w[6];   // nn.Linear weights [nout][nin]
  const float* b[6];
};

__global__ void flow1d_pack_kernel(F1Weights src, F1Dims d, float* __restrict__ PF, float* __restrict__ PB) {
  const int total = d.total_fwd + (PB ? d.total_bwd : 0);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    if (e < d.total_fwd) {
      int l = F1_LAYERS - 1;
      while (e < d.offWT[l]) --l;
      const int r = e - d.offWT[l];
      const float* W = l == 0 ? src.Wf : src.w[l - 1];
      const float* bsrc = l == 0 ? src.bf : src.b[l - 1];
      float v = 0.f;
      if (r < d.nin[l] * d.noutp[l]) {
        const int i = r / d.noutp[l], o = r % d.noutp[l];
        if (o < d.nout[l]) v = W[o * d.nin[l] + i];
      } else {
        const int o = r - d.nin[l] * d.noutp[l];
        if (o < d.nout[l]) v = bsrc[o];
      }
      PF[e] = v;
    } else {
      const int q = e - d.total_fwd;
      int l = F1_LAYERS - 1;
      while (q < d.offW[l]) --l;
      const int r = q - d.offW[l];
      const int o = r / d.ninp[l], i = r % d.ninp[l];
      const float* W = l == 0 ? src.Wf : src.w[l - 1];
      PB[q] = i < d.nin[l] ? W[o * d.nin[l] + i] : 0.f;
    }
  }
}

constexpr int F1_THREADS = 512;

// Every layer is a small GEMM over the sample tile: out[o][s] = epi(bias[o] + sum_i WT[i][o] * in[i][s]).
// The CTA shares the tile; a thread owns 8 outputs x 2 consecutive samples (16 accumulators): one LDS.64 of
// activations and two broadcast LDS.128 of weights feed 16 FMAs, and a 128-sample tile of a 32-wide layer still
// spreads over 256 threads. Tiles live feature-major in shared memory with row stride ld = NT + 4 floats.
// MODE: 0 none, 1 relu, 2 tanh (forward epilogues); 3 multiply by relu'(hm), 4 multiply by 1 - hm^2 (dgrad epilogues)
template <int MODE>
__device__ __forceinline__ void lin_tile(const float* __restrict__ in, int nin, float* __restrict__ out, int nout,
                                         const float* __restrict__ WT, const float* __restrict__ bias, int noutp,
                                         const float* __restrict__ hm, int ld, int NT) {
  const int nsg = NT >> 1;
  const int items = (noutp >> 3) * nsg;
  for (int it = threadIdx.x; it < items; it += F1_THREADS) {
    const int og = it / nsg, sg = it - og * nsg;
    float acc[8][2];
    if (bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + og * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + og * 8 + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j][0] = bb[j]; acc[j][1] = bb[j]; }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; }
    }
    const float* wp = WT + og * 8;
    const float* ip = in + 2 * sg;
#pragma unroll 8
    for (int i = 0; i < nin; ++i) {
      const float2 a = *reinterpret_cast<const float2*>(ip + i * ld);
      const float4 w0 = *reinterpret_cast<const float4*>(wp + i * noutp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + i * noutp + 4);
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j][0] = fmaf(ww[j], a.x, acc[j][0]);
        acc[j][1] = fmaf(ww[j], a.y, acc[j][1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = og * 8 + j;
      if (o < nout) {
        float2 v = make_float2(acc[j][0], acc[j][1]);
        if (MODE == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
        if (MODE == 2) { v.x = tanhf(v.x); v.y = tanhf(v.y); }
        if (MODE == 3 || MODE == 4) {
          const float2 h = *reinterpret_cast<const float2*>(hm + o * ld + 2 * sg);
          if (MODE == 3) {
            v.x = h.x > 0.f ? v.x : 0.f; v.y = h.y > 0.f ? v.y : 0.f;
          } else {
            v.x *= 1.f - h.x * h.x; v.y *= 1.f - h.y * h.y;
          }
        }
        *reinterpret_cast<float2*>(out + o * ld + 2 * sg) = v;
      }
    }
  }
}

// Weight gradient of one layer over the sample tile, accumulated into the CTA's shared gradient block:
//   gacc[o*ninp + i] += sum_s dout[o][s] * in[i][s],   gbias[o] += sum_s dout[o][s].
// A thread owns an 8 (o, interleaved) x 4 (i, interleaved) block of dW and walks its share of the samples four at a
// time (12 LDS.128 per 128 FMAs). Small layers split the samples over ks adjacent lanes, which meet in a shuffle
// reduction; every dW element has exactly one writer, so there are no atomics.
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ dout, int nout, const float* __restrict__ in,
                                           int nin, int ninp, float* __restrict__ gacc, float* __restrict__ gbias,
                                           int ld, int NT) {
  const int n_ot = (nout + 7) >> 3, n_it = (nin + 3) >> 2;
  const int ntiles = n_ot * n_it;
  const int nsg = NT >> 2;
  int ks = 1;
  while (ks < 32 && ntiles * ks * 2 <= F1_THREADS && (nsg % (ks * 2)) == 0 && nsg / (ks * 2) >= 2) ks <<= 1;
  const int kq = nsg / ks;
  const int total = ntiles * ks;
  const int lane = threadIdx.x & 31;
  for (int base = threadIdx.x - lane; base < total; base += F1_THREADS) {
    const int item = base + lane;
    const bool valid = item < total;
    const int t = valid ? item / ks : 0, kp = valid ? item - t * ks : 0;
    const int ot = t / n_it, itl = t - ot * n_it;
    int dp[8], ip[4];  // row offsets (rows past the end are clamped; their results are dropped below)
#pragma unroll
    for (int j = 0; j < 8; ++j) dp[j] = min(ot + j * n_ot, nout - 1) * ld;
#pragma unroll
    for (int m = 0; m < 4; ++m) ip[m] = min(itl + m * n_it, nin - 1) * ld;
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll 2
    for (int q = kp * kq; q < (kp + 1) * kq; ++q) {
      float4 a[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) a[m] = *reinterpret_cast<const float4*>(in + ip[m] + 4 * q);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 d = *reinterpret_cast<const float4*>(dout + dp[j] + 4 * q);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          acc[j][m] = fmaf(d.x, a[m].x, acc[j][m]);
          acc[j][m] = fmaf(d.y, a[m].y, acc[j][m]);
          acc[j][m] = fmaf(d.z, a[m].z, acc[j][m]);
          acc[j][m] = fmaf(d.w, a[m].w, acc[j][m]);
        }
      }
    }
    for (int sh = 1; sh < ks; sh <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[j][m] += __shfl_xor_sync(0xffffffffu, acc[j][m], sh);
    }
    if (valid && kp == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = ot + j * n_ot;
        if (o < nout) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int i = itl + m * n_it;
            if (i < nin) gacc[o * ninp + i] += acc[j][m];
          }
        }
      }
    }
  }
  // bias gradient: one warp per output row
  const int warp = threadIdx.x >> 5;
  for (int o = warp; o < nout; o += F1_THREADS / 32) {
    float a = 0.f;
    for (int s = lane; s < NT; s += 32) a += dout[o * ld + s];
    for (int sh = 16; sh > 0; sh >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sh);
    if (lane == 0) gbias[o] += a;
  }
}

__device__ __forceinline__ void f1_sigmoid(float t, float& s, float& ls) {
  const float e = expf(-fabsf(t));
  const float l1p = log1pf(e);
  if (t >= 0.f) { s = 1.f / (1.f + e); ls = -l1p; }
  else { s = e / (1.f + e); ls = t - l1p; }
}

// global [samples][n] row-major  ->  smem [n][ld] (samples past nvalid read as zero). A thread gathers the same
// feature of four consecutive samples (each warp load is one coalesced row segment) and stores one float4.
__device__ __forceinline__ void tile_load(const float* __restrict__ g, int n, float* __restrict__ sm, int ld,
                                          long long s0, int nvalid, int NT) {
  const int fx = threadIdx.x & 63, qy = threadIdx.x >> 6;
  const int nsg = NT >> 2;
  for (int q = qy; q < nsg; q += F1_THREADS / 64) {
    const float* gp = g + (s0 + 4 * q) * n;
    const int left = nvalid - 4 * q;
    for (int f = fx; f < n; f += 64) {
      float4 v;
      v.x = left > 0 ? gp[f] : 0.f;
      v.y = left > 1 ? gp[n + f] : 0.f;
      v.z = left > 2 ? gp[2 * n + f] : 0.f;
      v.w = left > 3 ? gp[3 * n + f] : 0.f;
      *reinterpret_cast<float4*>(sm + f * ld + 4 * q) = v;
    }
  }
}
__device__ __forceinline__ void tile_store(float* __restrict__ g, int n, const float* __restrict__ sm, int ld,
                                           long long s0, int nvalid, int NT) {
  const int fx = threadIdx.x & 63, qy = threadIdx.x >> 6;
  const int nsg = NT >> 2;
  for (int q = qy; q < nsg; q += F1_THREADS / 64) {
    float* gp = g + (s0 + 4 * q) * n;
    const int left = nvalid - 4 * q;
    for (int f = fx; f < n; f += 64) {
      const float4 v = *reinterpret_cast<const float4*>(sm + f * ld + 4 * q);
      if (left > 0) gp[f] = v.x;
      if (left > 1) gp[n + f] = v.y;
      if (left > 2) gp[2 * n + f] = v.z;
      if (left > 3) gp[3 * n + f] = v.w;
    }
  }
}

__host__ __device__ static inline int up4(int x) { return (x + 3) / 4 * 4; }

// smem (floats): PF[total_fwd] | X[DP*ld] | Y[DP*ld] | A0[(D1+Cc)*ld if Cc] | Ha[hid*ld] | Hb[hid*ld]   (inference)
//                PF[total_fwd] | X[DP*ld] | Y[DP*ld] | A0[...] | ACT[n_act*ld]                             (acts saved)
// with DP = max(D, 2*D2) rows and ld = NT + 4.
__global__ void __launch_bounds__(F1_THREADS, 1)
flow1d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ cond, const float* __restrict__ PF,
                  const float* __restrict__ sl, float* __restrict__ y, const float* __restrict__ ld_in,
                  float* __restrict__ ld_out, float* __restrict__ acts, F1Dims d, int B, int reverse, int NT) {
  extern __shared__ __align__(16) float sm[];
  const int ld = NT + 4;
  const int DP = max(d.D, 2 * d.D2);
  float* W = sm;
  float* X = W + up4(d.total_fwd);
  float* Y = X + DP * ld;
  float* A0 = Y + DP * ld;
  float* Hbase = A0 + (d.Cc ? (d.D1 + d.Cc) * ld : 0);
  for (int e = threadIdx.x; e < d.total_fwd; e += F1_THREADS) W[e] = PF[e];
  const float sl0 = sl ? sl[0] : 0.f;
  const int tiles = (B + NT - 1) / NT;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long s0 = static_cast<long long>(t) * NT;
    const int nvalid = min(NT, B - static_cast<int>(s0));
    __syncthreads();
    tile_load(x, d.D, X, ld, s0, nvalid, NT);
    if (d.Cc) tile_load(cond, d.Cc, A0 + d.D1 * ld, ld, s0, nvalid, NT);
    __syncthreads();
    float* Z = X;  // the tensor the coupling acts on
    if (!reverse) {
      lin_tile<0>(X, d.D, Y, d.D, W + d.offWT[0], W + d.offB[0], d.noutp[0], nullptr, ld, NT);
      Z = Y;
      __syncthreads();
    }
    const float* a0 = Z;
    if (d.Cc) {
      for (int e = threadIdx.x; e < d.D1 * NT; e += F1_THREADS) {
        const int i = e / NT, s = e - i * NT;
        A0[i * ld + s] = Z[i * ld + s];
      }
      a0 = A0;
      __syncthreads();
    }
    float* H[6];
    if (acts) {
      for (int l = 0; l < 6; ++l) H[l] = Hbase + l * d.hid * ld;
    } else {
      float* Ha = Hbase;
      float* Hb = Hbase + d.hid * ld;
      H[0] = Ha; H[1] = Hb; H[2] = Ha; H[3] = Hb; H[4] = Ha;
      H[5] = reverse ? Y : X;  // forward: X is dead after the affine; reverse: Y is not used yet
    }
    lin_tile<1>(a0, d.nin[1], H[0], d.hid, W + d.offWT[1], W + d.offB[1], d.noutp[1], nullptr, ld, NT);
    __syncthreads();
    lin_tile<1>(H[0], d.hid, H[1], d.hid, W + d.offWT[2], W + d.offB[2], d.noutp[2], nullptr, ld, NT);
    __syncthreads();
    lin_tile<1>(H[1], d.hid, H[2], d.hid, W + d.offWT[3], W + d.offB[3], d.noutp[3], nullptr, ld, NT);
    __syncthreads();
    lin_tile<1>(H[2], d.hid, H[3], d.hid, W + d.offWT[4], W + d.offB[4], d.noutp[4], nullptr, ld, NT);
    __syncthreads();
    lin_tile<2>(H[3], d.hid, H[4], d.hid, W + d.offWT[5], W + d.offB[5], d.noutp[5], nullptr, ld, NT);
    __syncthreads();
    lin_tile<0>(H[4], d.hid, H[5], 2 * d.D2, W + d.offWT[6], W + d.offB[6], d.noutp[6], nullptr, ld, NT);
    __syncthreads();
    // affine coupling + per-sample log-det: four threads per sample (j mod 4), combined by two shuffles
    {
      const float* O = H[5];
      const int s = threadIdx.x >> 2, half = threadIdx.x & 3;
      float ldacc = 0.f;
      if (s < NT) {
        for (int j = half; j < d.D2; j += 4) {
          const float sh = O[(2 * j) * ld + s], lg = O[(2 * j + 1) * ld + s];
          float sg, ls;
          f1_sigmoid(lg + 2.f, sg, ls);
          const float z2 = Z[(d.D1 + j) * ld + s];
          Z[(d.D1 + j) * ld + s] = reverse ? (z2 / sg - sh) : (z2 + sh) * sg;
          ldacc += ls;
        }
      }
      ldacc += __shfl_xor_sync(0xffffffffu, ldacc, 1);
      ldacc += __shfl_xor_sync(0xffffffffu, ldacc, 2);
      if (reverse) ldacc = -ldacc;
      if (ld_out && half == 0 && s < nvalid) ld_out[s0 + s] = ld_in[s0 + s] + sl0 + ldacc;
    }
    __syncthreads();
    float* OUT = Z;
    if (reverse) {
      // inverse affine last; Y is free (it only ever held the MLP output, consumed above)
      lin_tile<0>(X, d.D, Y, d.D, W + d.offWT[0], W + d.offB[0], d.noutp[0], nullptr, ld, NT);
      OUT = Y;
      __syncthreads();
    }
    tile_store(y, d.D, OUT, ld, s0, nvalid, NT);
    if (acts) tile_store(acts, d.n_act, Hbase, ld, s0, nvalid, NT);
  }
}

// Backward of one 1-D FlowStep (either direction). Reads the saved MLP activations, the incoming gradients and
// (forward direction) the step OUTPUT y_out, whose first D1 features are the MLP input and whose last D2 features are
// (y2 + shift) * scale, or (reverse direction) the step input; writes the input gradient and accumulates the
// parameter gradients into G (global, pre-zeroed).
// smem (floats): PB[total_bwd] | GA[total_grad] | XY[DR*ld] | GZ[DR*ld] | A0[(D1+Cc)*ld if Cc] | ACT[n_act*ld] |
//                SC[scr*ld]      DR = D rounded up to 8; scr = hid | D1+Cc (forward direction), DR (reverse)
__global__ void __launch_bounds__(F1_THREADS, 1)
flow1d_bwd_kernel(const float* __restrict__ x_in, const float* __restrict__ cond, const float* __restrict__ acts,
                  const float* __restrict__ PB, const float* __restrict__ y_out, const float* __restrict__ g_out,
                  const float* __restrict__ g_ld, float* __restrict__ dx, float* __restrict__ G, F1Dims d, int B,
                  int reverse, int scr, int NT) {
  extern __shared__ __align__(16) float sm[];
  const int ld = NT + 4;
  const int DR = up8d(d.D);
  float* W = sm;
  float* GA = W + up4(d.total_bwd);
  float* XY = GA + up4(d.total_grad);
  float* GZ = XY + DR * ld;
  float* A0 = GZ + DR * ld;
  float* ACT = A0 + (d.Cc ? (d.D1 + d.Cc) * ld : 0);
  float* SC = ACT + d.n_act * ld;
  for (int e = threadIdx.x; e < d.total_bwd; e += F1_THREADS) W[e] = PB[e];
  for (int e = threadIdx.x; e < d.total_grad; e += F1_THREADS) GA[e] = 0.f;
  const int tiles = (B + NT - 1) / NT;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long s0 = static_cast<long long>(t) * NT;
    const int nvalid = min(NT, B - static_cast<int>(s0));
    __syncthreads();
    tile_load(reverse ? x_in : y_out, d.D, XY, ld, s0, nvalid, NT);
    tile_load(g_out, d.D, GZ, ld, s0, nvalid, NT);
    tile_load(acts, d.n_act, ACT, ld, s0, nvalid, NT);
    if (d.Cc) tile_load(cond, d.Cc, A0 + d.D1 * ld, ld, s0, nvalid, NT);
    __syncthreads();
    float* H[5];
    for (int l = 0; l < 5; ++l) H[l] = ACT + l * d.hid * ld;
    float* O = ACT + 5 * d.hid * ld;
    float* gz = GZ;      // gradient wrt the coupling output
    float* scr_buf = SC;  // scratch rows for the MLP backward
    if (reverse) {
      // reverse step = coupling^-1 then x = Wi z' + bi. Rebuild z' in place of the input's second half, take the
      // affine's weight gradient against it and pull g_out through Wi.
      for (int e = threadIdx.x; e < d.D2 * NT; e += F1_THREADS) {
        const int j = e / NT, s = e - j * NT;
        float sg, ls;
        f1_sigmoid(O[(2 * j + 1) * ld + s] + 2.f, sg, ls);
        XY[(d.D1 + j) * ld + s] = XY[(d.D1 + j) * ld + s] / sg - O[(2 * j) * ld + s];
      }
      __syncthreads();
      wgrad_tile(GZ, d.D, XY, d.D, d.ninp[0], GA + d.offG[0], GA + d.offGB[0], ld, NT);
      lin_tile<0>(GZ, d.D, SC, d.D, W + d.offW[0], nullptr, d.ninp[0], nullptr, ld, NT);  // dL/dz'
      __syncthreads();
      gz = SC;
      scr_buf = GZ;
    }
    const float* a0 = XY;  // first D1 features: the MLP input in either direction
    if (d.Cc) {
      for (int e = threadIdx.x; e < d.D1 * NT; e += F1_THREADS) {
        const int i = e / NT, s = e - i * NT;
        A0[i * ld + s] = XY[i * ld + s];
      }
      a0 = A0;
    }
    // coupling backward: dO in place over O; gz[D1+j] becomes the gradient wrt the pre-coupling z2
    for (int e = threadIdx.x; e < d.D2 * NT; e += F1_THREADS) {
      const int j = e / NT, s = e - j * NT;
      const float gl = (g_ld && s < nvalid) ? g_ld[s0 + s] : 0.f;
      const float sh = O[(2 * j) * ld + s];
      float sg, ls;
      f1_sigmoid(O[(2 * j + 1) * ld + s] + 2.f, sg, ls);
      const float g2 = gz[(d.D1 + j) * ld + s];
      const float v = XY[(d.D1 + j) * ld + s];  // forward: (y2 + sh) * sg;  reverse: z' = z2 / sg - sh
      float dsh, dlg, dz2;
      if (!reverse) {
        dsh = g2 * sg;
        dlg = (g2 * v + gl) * (1.f - sg);
        dz2 = g2 * sg;
      } else {
        dsh = -g2;
        dlg = -(g2 * (v + sh) + gl) * (1.f - sg);
        dz2 = g2 / sg;
      }
      O[(2 * j) * ld + s] = dsh;
      O[(2 * j + 1) * ld + s] = dlg;
      gz[(d.D1 + j) * ld + s] = dz2;
    }
    __syncthreads();
    // MLP backward, layer 6 .. 1. Gradient slots: layer 6 and layer 1 -> scratch, layers 5..2 -> the (consumed)
    // activation slot of the layer above.
    const float* dcur = O;
    int ncur = 2 * d.D2;
    for (int l = 6; l >= 1; --l) {
      const float* in = (l == 1) ? a0 : H[l - 2];
      const int nin = d.nin[l];
      wgrad_tile(dcur, ncur, in, nin, d.ninp[l], GA + d.offG[l], GA + d.offGB[l], ld, NT);
      float* dn = (l == 6 || l == 1) ? scr_buf : H[l - 1];
      if (l == 6) lin_tile<4>(dcur, ncur, dn, nin, W + d.offW[l], nullptr, d.ninp[l], H[l - 2], ld, NT);
      else if (l >= 2) lin_tile<3>(dcur, ncur, dn, nin, W + d.offW[l], nullptr, d.ninp[l], H[l - 2], ld, NT);
      else lin_tile<0>(dcur, ncur, dn, d.D1, W + d.offW[l], nullptr, d.ninp[l], nullptr, ld, NT);
      __syncthreads();
      dcur = dn;
      ncur = nin;
    }
    for (int e = threadIdx.x; e < d.D1 * NT; e += F1_THREADS) {
      const int i = e / NT, s = e - i * NT;
      gz[i * ld + s] += dcur[i * ld + s];
    }
    __syncthreads();
    const float* DX = gz;
    if (!reverse) {
      float* Xs = O;  // dO is dead: reuse its 2*D2 >= D rows for the step input
      tile_load(x_in, d.D, Xs, ld, s0, nvalid, NT);
      __syncthreads();
      wgrad_tile(gz, d.D, Xs, d.D, d.ninp[0], GA + d.offG[0], GA + d.offGB[0], ld, NT);
      lin_tile<0>(gz, d.D, XY, d.D, W + d.offW[0], nullptr, d.ninp[0], nullptr, ld, NT);  // dx = W'^T dy
      __syncthreads();
      DX = XY;
    }
    tile_store(dx, d.D, DX, ld, s0, nvalid, NT);
  }
  __syncthreads();
  // staggered start so the CTAs do not all hit the same addresses at once
  const int start = static_cast<int>((static_cast<long long>(blockIdx.x) * 997) % d.total_grad);
  for (int k = threadIdx.x; k < d.total_grad; k += F1_THREADS) {
    int e = k + start;
    if (e >= d.total_grad) e -= d.total_grad;
    const float v = GA[e];
    if (v != 0.f) atomicAdd(G + e, v);
  }
}

// Stand-alone per-row affine y = W x + b on [B, D] (ActNorm1d / InvertibleConv1x1 called on their own).
__global__ void affine_rows_kernel(const float* __restrict__ x, const float* __restrict__ Wf,
                                   const float* __restrict__ bf, const float* __restrict__ sl, float* __restrict__ y,
                                   const float* __restrict__ ld_in, float* __restrict__ ld_out, int B, int D,
                                   float pixels) {
  extern __shared__ float sm[];
  float* W = sm;
  float* b = W + D * D;
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) W[e] = Wf[e];
  for (int e = threadIdx.x; e < D; e += blockDim.x) b[e] = bf[e];
  __syncthreads();
  const long long total = static_cast<long long>(B) * D;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = e / D;
    const int o = static_cast<int>(e - r * D);
    const float* xr = x + r * D;
    float a = b[o];
    for (int i = 0; i < D; ++i) a = fmaf(W[o * D + i], xr[i], a);
    y[e] = a;
    if (o == 0 && ld_out) ld_out[r] = ld_in[r] + sl[0] * pixels;
  }
}

// Largest sample tile NT (multiple of 32, so the bank pattern of ld = NT + 4 holds) whose working set fits.
static int f1_tile_fwd(const F1Dims& d, bool save, int* smem_out) {
  const int DP = d.D > 2 * d.D2 ? d.D : 2 * d.D2;
  for (int nt = 128; nt >= 32; nt -= 32) {
    const int ld = nt + 4;
    long long fl = up4(d.total_fwd) + 2LL * DP * ld + (d.Cc ? (d.D1 + d.Cc) * ld : 0) +
                   (save ? 1LL * d.n_act * ld : 2LL * d.hid * ld);
    if (fl * 4 <= 225 * 1024) { *smem_out = static_cast<int>(fl * 4); return nt; }
  }
  return 0;
}

static int f1_tile_bwd(const F1Dims& d, int scr, int* smem_out) {
  const int DR = up8(d.D);
  for (int nt = 128; nt >= 32; nt -= 32) {
    const int ld = nt + 4;
    long long fl = up4(d.total_bwd) + up4(d.total_grad) + 2LL * DR * ld + (d.Cc ? (d.D1 + d.Cc) * ld : 0) +
                   1LL * d.n_act * ld + 1LL * scr * ld;
    if (fl * 4 <= 225 * 1024) { *smem_out = static_cast<int>(fl * 4); return nt; }
  }
  return 0;
}

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_flow1d_sizes(int D, int Cc, int hid, int* total_fwd, int* total_bwd, int* total_grad, int* n_act,
                                int* offsets /* [7*4]: offG, offGB, ninp, noutp per layer */) {
  if (D < 2 || hid <= 0 || Cc < 0) return NFK_ERR_SHAPE;
  const F1Dims d = f1_dims(D, Cc, hid);
  if (total_fwd) *total_fwd = d.total_fwd;
  if (total_bwd) *total_bwd = d.total_bwd;
  if (total_grad) *total_grad = d.total_grad;
  if (n_act) *n_act = d.n_act;
  if (offsets)
    for (int l = 0; l < F1_LAYERS; ++l) {
      offsets[4 * l + 0] = d.offG[l]; offsets[4 * l + 1] = d.offGB[l];
      offsets[4 * l + 2] = d.ninp[l]; offsets[4 * l + 3] = d.noutp[l];
    }
  return NFK_OK;
}

extern "C" int nfk_flow1d_pack(const float* Wf, const float* bf, const float* const* w, const float* const* b, int D,
                               int Cc, int hid, float* PF, float* PB, void* stream) {
  if (D < 2 || hid <= 0 || Cc < 0) return NFK_ERR_SHAPE;
  if (!Wf || !bf || !w || !b || !PF) return NFK_ERR_ARG;
  const F1Dims d = f1_dims(D, Cc, hid);
  F1Weights src{};
  src.Wf = Wf; src.bf = bf;
  for (int l = 0; l < 6; ++l) { src.w[l] = w[l]; src.b[l] = b[l]; if (!w[l] || !b[l]) return NFK_ERR_ARG; }
  const int total = d.total_fwd + (PB ? d.total_bwd : 0);
  flow1d_pack_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, d, PF, PB);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_flow1d_fwd(const float* x, const float* cond, const float* PF, const float* sl, float* y,
                              const float* ld_in, float* ld_out, float* acts, int B, int D, int Cc, int hid,
                              int reverse, void* stream) {
  if (B <= 0 || D < 2 || hid <= 0 || Cc < 0) return NFK_ERR_SHAPE;
  if (!x || !PF || !y || (Cc && !cond) || (ld_out && !ld_in)) return NFK_ERR_ARG;
  const F1Dims d = f1_dims(D, Cc, hid);
  int smem = 0;
  const int nt = f1_tile_fwd(d, acts != nullptr, &smem);
  if (!nt) return NFK_ERR_SHAPE;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(flow1d_fwd_kernel), smem)) return rc;
  const int tiles = (B + nt - 1) / nt;
  const int grid = tiles < 148 ? tiles : 148;
  flow1d_fwd_kernel<<<grid, F1_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(x, cond, PF, sl, y, ld_in, ld_out,
                                                                                  acts, d, B, reverse, nt);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_flow1d_bwd(const float* x_in, const float* cond, const float* acts, const float* PB,
                              const float* y_out, const float* g_out, const float* g_ld, float* dx, float* G, int B,
                              int D, int Cc, int hid, int reverse, void* stream) {
  if (B <= 0 || D < 2 || hid <= 0 || Cc < 0) return NFK_ERR_SHAPE;
  if (!x_in || !acts || !PB || !g_out || !dx || !G || (Cc && !cond) || (!reverse && !y_out)) return NFK_ERR_ARG;
  const F1Dims d = f1_dims(D, Cc, hid);
  int scr = hid > d.nin[1] ? hid : d.nin[1];
  if (reverse) scr = up8(D);
  int smem = 0;
  const int nt = f1_tile_bwd(d, scr, &smem);
  if (!nt) return NFK_ERR_SHAPE;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(flow1d_bwd_kernel), smem)) return rc;
  const int tiles = (B + nt - 1) / nt;
  const int grid = tiles < 148 ? tiles : 148;
  flow1d_bwd_kernel<<<grid, F1_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(x_in, cond, acts, PB, y_out, g_out,
                                                                                  g_ld, dx, G, d, B, reverse, scr, nt);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_affine_rows(const float* x, const float* Wf, const float* bf, const float* sl, float* y,
                               const float* ld_in, float* ld_out, int B, int D, float pixels, void* stream) {
  if (B <= 0 || D <= 0 || D > 160) return NFK_ERR_SHAPE;
  if (!x || !Wf || !bf || !y || (ld_out && (!ld_in || !sl))) return NFK_ERR_ARG;
  const int smem = (D * D + D) * 4;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(affine_rows_kernel), smem)) return rc;
  const long long total = static_cast<long long>(B) * D;
  const int grid = static_cast<int>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  affine_rows_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, Wf, bf, sl, y, ld_in, ld_out, B, D,
                                                                           pixels);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
