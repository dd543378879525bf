// 1-D (tabular) Glow FlowStep, fused end to end in one kernel per direction (reference: the is_1d branches of
// models/flows.py:37-52,142-202 and models/layers.py:76,117,410-411).
//
// One thread owns one sample; a CTA walks tiles of NT samples with every weight of the step resident in shared
// memory. Activations live in shared memory feature-major ([feature][NT+1]) so a thread's own column is bank-conflict
// free, and each layer is a register-blocked (8 outputs) FMA loop fed by one broadcast LDS.128 of weights per 4 FMAs:
//   y = W' x + b'            (ActNorm1d + x @ W folded by nfk_invconv_prep with transpose=1)
//   h = MLP(y1 [, cond])     Linear-ReLU x4, Linear-Tanh, Linear
//   y2 = (y2 + h[0::2]) * sigmoid(h[1::2] + 2),  logdet += sum log sigmoid          (or the inverse order / formulas)
// The backward kernel reloads the saved MLP activations, walks the chain in reverse and reduces weight gradients
// per CTA in shared memory (one thread per (out,in) pair), then adds them to the global gradient block once.
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

// out[o][s] = act(bias[o] + sum_i WT[i][o] * in[i][s]) for this thread's sample s. ACT: 0 none, 1 relu, 2 tanh.
template <int ACT>
__device__ __forceinline__ void lin_fwd(const float* __restrict__ in, int nin, float* __restrict__ out, int nout,
                                        const float* __restrict__ WT, const float* __restrict__ bias, int noutp,
                                        int ld, int s) {
  for (int o0 = 0; o0 < nout; o0 += 8) {
    float acc[8];
    const float4 b0 = *reinterpret_cast<const float4*>(bias + o0);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + o0 + 4);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    const float* wp = WT + o0;
#pragma unroll 4
    for (int i = 0; i < nin; ++i) {
      const float a = in[i * ld + s];
      const float4 w0 = *reinterpret_cast<const float4*>(wp + i * noutp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + i * noutp + 4);
      acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
      acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
      acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
      acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (o0 + j < nout) {
        float v = acc[j];
        if (ACT == 1) v = fmaxf(v, 0.f);
        if (ACT == 2) v = tanhf(v);
        out[(o0 + j) * ld + s] = v;
      }
    }
  }
}

// din[i][s] = sum_o W[o][i] * dout[o][s]   (W in [nout][ninp] layout, i contiguous)
__device__ __forceinline__ void lin_bwd_in(const float* __restrict__ dout, int nout, float* __restrict__ din, int nin,
                                           const float* __restrict__ W, int ninp, int ld, int s) {
  for (int i0 = 0; i0 < nin; i0 += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* wp = W + i0;
#pragma unroll 4
    for (int o = 0; o < nout; ++o) {
      const float a = dout[o * ld + s];
      const float4 w0 = *reinterpret_cast<const float4*>(wp + o * ninp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + o * ninp + 4);
      acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
      acc[2] = fmaf(a, w0.z, acc[2]); acc[3] = fmaf(a, w0.w, acc[3]);
      acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
      acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i0 + j < nin) din[(i0 + j) * ld + s] = acc[j];
  }
}

// CTA-wide: gacc[o*ninp + i] += sum_s dout[o][s] * in[i][s];  gbias[o] += sum_s dout[o][s]
__device__ __forceinline__ void lin_bwd_w(const float* __restrict__ dout, int nout, const float* __restrict__ in,
                                          int nin, int ninp, float* __restrict__ gacc, float* __restrict__ gbias,
                                          int ld, int nvalid) {
  const int pairs = nout * nin;
  for (int p = threadIdx.x; p < pairs + nout; p += blockDim.x) {
    float a = 0.f;
    if (p < pairs) {
      const int o = p / nin, i = p - o * nin;
      const float* dp = dout + o * ld;
      const float* ip = in + i * ld;
      for (int s = 0; s < nvalid; ++s) a = fmaf(dp[s], ip[s], a);
      gacc[o * ninp + i] += a;
    } else {
      const int o = p - pairs;
      const float* dp = dout + o * ld;
      for (int s = 0; s < nvalid; ++s) a += dp[s];
      gbias[o] += a;
    }
  }
}

__device__ __forceinline__ void f1_sigmoid(float t, float& s, float& ls) {
  const float e = expf(-fabsf(t));
  const float l1p = log1pf(e);
  if (t >= 0.f) { s = 1.f / (1.f + e); ls = -l1p; }
  else { s = e / (1.f + e); ls = t - l1p; }
}

__device__ __forceinline__ void tile_load(const float* __restrict__ g, int n, float* __restrict__ sm, int ld,
                                          long long s0, int nvalid) {
  // global [samples][n] row-major  ->  smem [n][ld]
  const long long base = s0 * n;
  for (int e = threadIdx.x; e < nvalid * n; e += blockDim.x) {
    const int s = e / n, f = e - s * n;
    sm[f * ld + s] = g[base + e];
  }
}
__device__ __forceinline__ void tile_store(float* __restrict__ g, int n, const float* __restrict__ sm, int ld,
                                           long long s0, int nvalid) {
  const long long base = s0 * n;
  for (int e = threadIdx.x; e < nvalid * n; e += blockDim.x) {
    const int s = e / n, f = e - s * n;
    g[base + e] = sm[f * ld + s];
  }
}

// smem (floats): PF[total_fwd] | X[DP*ld] | Y[DP*ld] | A0[(D1+Cc)*ld if Cc] | Ha[hid*ld] | Hb[hid*ld] |
//                ACT[n_act*ld if acts are saved]          with DP = max(D, 2*D2) rows
__global__ void flow1d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ cond,
                                  const float* __restrict__ PF, const float* __restrict__ sl, float* __restrict__ y,
                                  const float* __restrict__ ld_in, float* __restrict__ ld_out,
                                  float* __restrict__ acts, F1Dims d, int B, int reverse) {
  extern __shared__ float sm[];
  const int NT = blockDim.x, ld = NT + 1, s = threadIdx.x;
  const int DP = max(d.D, 2 * d.D2);
  float* W = sm;
  float* X = W + d.total_fwd;
  float* Y = X + DP * ld;
  float* A0 = Y + DP * ld;
  float* Ha = A0 + (d.Cc ? (d.D1 + d.Cc) * ld : 0);
  float* Hb = Ha + d.hid * ld;
  float* ACT = Hb + d.hid * ld;  // only when acts != nullptr: H1..H5, O
  for (int e = threadIdx.x; e < d.total_fwd; e += NT) W[e] = PF[e];
  const float sl0 = sl ? sl[0] : 0.f;
  const int tiles = (B + NT - 1) / NT;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long s0 = static_cast<long long>(t) * NT;
    const int nvalid = min(NT, B - static_cast<int>(s0));
    __syncthreads();
    tile_load(x, d.D, X, ld, s0, nvalid);
    if (d.Cc) tile_load(cond, d.Cc, A0 + d.D1 * ld, ld, s0, nvalid);
    __syncthreads();
    float* Z = X;  // the tensor the coupling acts on
    if (!reverse) {
      lin_fwd<0>(X, d.D, Y, d.D, W + d.offWT[0], W + d.offB[0], d.noutp[0], ld, s);
      Z = Y;
    }
    const float* a0 = Z;
    if (d.Cc) {
      for (int i = 0; i < d.D1; ++i) A0[i * ld + s] = Z[i * ld + s];
      a0 = A0;
    }
    float* H[6];
    if (acts) {
      for (int l = 0; l < 6; ++l) H[l] = ACT + l * d.hid * ld;
    } else {
      H[0] = Ha; H[1] = Hb; H[2] = Ha; H[3] = Hb; H[4] = Ha;
      H[5] = reverse ? Y : X;  // forward: X is dead after the affine; reverse: Y is not used yet
    }
    lin_fwd<1>(a0, d.nin[1], H[0], d.hid, W + d.offWT[1], W + d.offB[1], d.noutp[1], ld, s);
    lin_fwd<1>(H[0], d.hid, H[1], d.hid, W + d.offWT[2], W + d.offB[2], d.noutp[2], ld, s);
    lin_fwd<1>(H[1], d.hid, H[2], d.hid, W + d.offWT[3], W + d.offB[3], d.noutp[3], ld, s);
    lin_fwd<1>(H[2], d.hid, H[3], d.hid, W + d.offWT[4], W + d.offB[4], d.noutp[4], ld, s);
    lin_fwd<2>(H[3], d.hid, H[4], d.hid, W + d.offWT[5], W + d.offB[5], d.noutp[5], ld, s);
    lin_fwd<0>(H[4], d.hid, H[5], 2 * d.D2, W + d.offWT[6], W + d.offB[6], d.noutp[6], ld, s);
    float ldacc = 0.f;
    const float* O = H[5];
    for (int j = 0; j < d.D2; ++j) {
      const float sh = O[(2 * j) * ld + s], lg = O[(2 * j + 1) * ld + s];
      float sg, ls;
      f1_sigmoid(lg + 2.f, sg, ls);
      const float z2 = Z[(d.D1 + j) * ld + s];
      Z[(d.D1 + j) * ld + s] = reverse ? (z2 / sg - sh) : (z2 + sh) * sg;
      ldacc += ls;
    }
    float* OUT = Z;
    if (reverse) {
      // inverse affine last; Y is free (it only ever held the MLP output, consumed above)
      lin_fwd<0>(X, d.D, Y, d.D, W + d.offWT[0], W + d.offB[0], d.noutp[0], ld, s);
      OUT = Y;
      ldacc = -ldacc;
    }
    if (ld_out && s < nvalid) ld_out[s0 + s] = ld_in[s0 + s] + sl0 + ldacc;
    __syncthreads();
    tile_store(y, d.D, OUT, ld, s0, nvalid);
    if (acts) tile_store(acts, d.n_act, ACT, ld, s0, nvalid);
  }
}

// Backward of one 1-D FlowStep (either direction). Reads the step input, the saved MLP activations and the
// incoming gradients; writes the input gradient and accumulates parameter gradients into G (global, pre-zeroed).
// smem (floats): PB[total_bwd] | GA[total_grad] | WTaff[D*noutp0 + noutp0] | X[D*ld] | GZ[D*ld] |
//                A0[(D1+Cc)*ld if Cc] | ACT[n_act*ld] | Da[dmax*ld] | Db[dmax*ld]   dmax >= max(D, hid, 2*D2, D1+Cc)
__global__ void flow1d_bwd_kernel(const float* __restrict__ x_in, const float* __restrict__ cond,
                                  const float* __restrict__ acts, const float* __restrict__ PB,
                                  const float* __restrict__ PF_aff, const float* __restrict__ g_out,
                                  const float* __restrict__ g_ld, float* __restrict__ dx, float* __restrict__ G,
                                  F1Dims d, int B, int reverse, int dmax) {
  extern __shared__ float sm[];
  const int NT = blockDim.x, ld = NT + 1, s = threadIdx.x;
  float* W = sm;
  float* GA = W + d.total_bwd;
  float* WTaff = GA + d.total_grad;  // forward affine (WT_0 | bias_0), to recompute y in the forward direction
  float* X = WTaff + d.D * d.noutp[0] + d.noutp[0];   // (everything above stays 16-byte aligned for LDS.128)
  float* GZ = X + d.D * ld;
  float* A0 = GZ + d.D * ld;
  float* ACT = A0 + (d.Cc ? (d.D1 + d.Cc) * ld : 0);
  float* Da = ACT + d.n_act * ld;
  float* Db = Da + dmax * ld;
  for (int e = threadIdx.x; e < d.total_bwd; e += NT) W[e] = PB[e];
  for (int e = threadIdx.x; e < d.total_grad; e += NT) GA[e] = 0.f;
  if (!reverse)
    for (int e = threadIdx.x; e < d.D * d.noutp[0] + d.noutp[0]; e += NT) WTaff[e] = PF_aff[e];
  const int tiles = (B + NT - 1) / NT;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long s0 = static_cast<long long>(t) * NT;
    const int nvalid = min(NT, B - static_cast<int>(s0));
    const bool live = s < nvalid;
    __syncthreads();
    tile_load(x_in, d.D, X, ld, s0, nvalid);
    tile_load(g_out, d.D, GZ, ld, s0, nvalid);
    tile_load(acts, d.n_act, ACT, ld, s0, nvalid);
    if (d.Cc) tile_load(cond, d.Cc, A0 + d.D1 * ld, ld, s0, nvalid);
    __syncthreads();
    float* H[5];
    for (int l = 0; l < 5; ++l) H[l] = ACT + l * d.hid * ld;
    float* O = ACT + 5 * d.hid * ld;
    const float gl = (g_ld && live) ? g_ld[s0 + s] : 0.f;
    // Z = pre-coupling tensor; GZ = gradient wrt the coupling output (after this block)
    const float* Z;
    if (!reverse) {
      lin_fwd<0>(X, d.D, Da, d.D, WTaff, WTaff + d.D * d.noutp[0], d.noutp[0], ld, s);  // y = W' x + b'
      Z = Da;
    } else {
      // reverse step = coupling^-1 then x = Wi z' + bi: pull g_out through the inverse affine first
      for (int i = 0; i < d.D1; ++i) Da[i * ld + s] = X[i * ld + s];
      for (int j = 0; j < d.D2; ++j) {
        float sg, ls;
        f1_sigmoid(O[(2 * j + 1) * ld + s] + 2.f, sg, ls);
        Da[(d.D1 + j) * ld + s] = X[(d.D1 + j) * ld + s] / sg - O[(2 * j) * ld + s];  // z'
      }
      __syncthreads();
      lin_bwd_w(GZ, d.D, Da, d.D, d.ninp[0], GA + d.offG[0], GA + d.offGB[0], ld, nvalid);
      __syncthreads();
      lin_bwd_in(GZ, d.D, Db, d.D, W + d.offW[0], d.ninp[0], ld, s);  // dL/dz'
      for (int i = 0; i < d.D; ++i) GZ[i * ld + s] = Db[i * ld + s];
      Z = X;
    }
    const float* a0 = Z;
    if (d.Cc) {
      for (int i = 0; i < d.D1; ++i) A0[i * ld + s] = Z[i * ld + s];
      a0 = A0;
    }
    // coupling backward: dO in place over O; GZ[D1+j] becomes the gradient wrt the pre-coupling z2
    for (int j = 0; j < d.D2; ++j) {
      const float sh = O[(2 * j) * ld + s];
      float sg, ls;
      f1_sigmoid(O[(2 * j + 1) * ld + s] + 2.f, sg, ls);
      const float g2 = GZ[(d.D1 + j) * ld + s];
      const float z2 = Z[(d.D1 + j) * ld + s];
      float dsh, dlg, dz2;
      if (!reverse) {
        dsh = g2 * sg;
        dlg = (g2 * (z2 + sh) * sg + gl) * (1.f - sg);
        dz2 = g2 * sg;
      } else {
        dsh = -g2;
        dlg = -(g2 * (z2 / sg) + gl) * (1.f - sg);
        dz2 = g2 / sg;
      }
      O[(2 * j) * ld + s] = dsh;
      O[(2 * j + 1) * ld + s] = dlg;
      GZ[(d.D1 + j) * ld + s] = dz2;
    }
    // MLP backward, layer 6 .. 1. Gradient slots: layer 6 -> Db, layers 5..2 -> the (consumed) activation slot of
    // the layer above, layer 1 -> Db (its input may be wider than hid).
    const float* dcur = O;
    int ncur = 2 * d.D2;
    for (int l = 6; l >= 1; --l) {
      const float* in = (l == 1) ? a0 : H[l - 2];
      const int nin = d.nin[l];
      __syncthreads();
      lin_bwd_w(dcur, ncur, in, nin, d.ninp[l], GA + d.offG[l], GA + d.offGB[l], ld, nvalid);
      __syncthreads();
      float* dn = (l == 6 || l == 1) ? Db : H[l - 1];
      lin_bwd_in(dcur, ncur, dn, nin, W + d.offW[l], d.ninp[l], ld, s);
      if (l >= 2) {
        const float* h = H[l - 2];  // output of layer l-1: tanh for l-1 == 5, ReLU otherwise
        for (int i = 0; i < nin; ++i) {
          const float hv = h[i * ld + s];
          dn[i * ld + s] *= (l == 6) ? (1.f - hv * hv) : (hv > 0.f ? 1.f : 0.f);
        }
      }
      dcur = dn;
      ncur = nin;
    }
    for (int i = 0; i < d.D1; ++i) GZ[i * ld + s] += dcur[i * ld + s];
    const float* DX = GZ;
    if (!reverse) {
      __syncthreads();
      lin_bwd_w(GZ, d.D, X, d.D, d.ninp[0], GA + d.offG[0], GA + d.offGB[0], ld, nvalid);
      __syncthreads();
      lin_bwd_in(GZ, d.D, Da, d.D, W + d.offW[0], d.ninp[0], ld, s);  // dx = W'^T dy  (Da: y is dead now)
      DX = Da;
    }
    __syncthreads();
    tile_store(dx, d.D, DX, ld, s0, nvalid);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d.total_grad; e += NT) {
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

static int f1_threads_fwd(const F1Dims& d, bool save, int* smem_out) {
  for (int nt = 128; nt >= 32; nt >>= 1) {
    const int ld = nt + 1;
    const int DP = d.D > 2 * d.D2 ? d.D : 2 * d.D2;
    long long fl = d.total_fwd + 2LL * DP * ld + (d.Cc ? (d.D1 + d.Cc) * ld : 0) + 2LL * d.hid * ld +
                   (save ? 1LL * d.n_act * ld : 0);
    if (fl * 4 <= 220 * 1024) { *smem_out = static_cast<int>(fl * 4); return nt; }
  }
  return 0;
}

static int f1_threads_bwd(const F1Dims& d, int dmax, int* smem_out) {
  for (int nt = 128; nt >= 32; nt >>= 1) {
    const int ld = nt + 1;
    long long fl = d.total_bwd + d.total_grad + 2LL * d.D * ld + (d.Cc ? (d.D1 + d.Cc) * ld : 0) +
                   1LL * d.n_act * ld + 2LL * dmax * ld + d.D * d.noutp[0] + d.noutp[0];
    if (fl * 4 <= 220 * 1024) { *smem_out = static_cast<int>(fl * 4); return nt; }
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
  const int nt = f1_threads_fwd(d, acts != nullptr, &smem);
  if (!nt) return NFK_ERR_SHAPE;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(flow1d_fwd_kernel), smem)) return rc;
  const int tiles = (B + nt - 1) / nt;
  const int grid = tiles < 2 * 148 ? tiles : 2 * 148;
  flow1d_fwd_kernel<<<grid, nt, smem, static_cast<cudaStream_t>(stream)>>>(x, cond, PF, sl, y, ld_in, ld_out, acts, d,
                                                                          B, reverse);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_flow1d_bwd(const float* x_in, const float* cond, const float* acts, const float* PB,
                              const float* PF, const float* g_out, const float* g_ld, float* dx, float* G, int B,
                              int D, int Cc, int hid, int reverse, void* stream) {
  if (B <= 0 || D < 2 || hid <= 0 || Cc < 0) return NFK_ERR_SHAPE;
  if (!x_in || !acts || !PB || !PF || !g_out || !dx || !G || (Cc && !cond)) return NFK_ERR_ARG;
  const F1Dims d = f1_dims(D, Cc, hid);
  int dmax = D;
  if (hid > dmax) dmax = hid;
  if (2 * d.D2 > dmax) dmax = 2 * d.D2;
  if (d.nin[1] > dmax) dmax = d.nin[1];
  int smem = 0;
  const int nt = f1_threads_bwd(d, dmax, &smem);
  if (!nt) return NFK_ERR_SHAPE;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(flow1d_bwd_kernel), smem)) return rc;
  const int tiles = (B + nt - 1) / nt;
  const int grid = tiles < 148 ? tiles : 148;
  flow1d_bwd_kernel<<<grid, nt, smem, static_cast<cudaStream_t>(stream)>>>(x_in, cond, acts, PB, PF + d.offWT[0],
                                                                          g_out, g_ld, dx, G, d, B, reverse, dmax);
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
