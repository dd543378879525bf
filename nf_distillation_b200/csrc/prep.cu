// Parameter-space kernels ("K0"): everything that depends only on the weights, once per optimiser step.
//
//  * invconv_prep: builds the fused ActNorm o InvertibleConv1x1 matrix of one FlowStep from the LU
//    parametrisation (reference: models/layers.py:376-397 get_weight, :101-142 ActNorm) — forward
//    W' = P L U diag(e^logs), b' = W' b, and inverse diag(e^-logs) U^-1 L^-1 P^T via in-SM triangular solves
//    (replaces three torch.inverse calls) — plus log|det| = sum(logs) + sum(log_s). The non-LU branch
//    (:366-375) runs a Gauss-Jordan slogdet/inverse in shared memory.
//  * invconv_prep_bwd: chain rule from (dW', db') back to actnorm.{bias,logs}, invconv.{lower,upper,log_s}.
//  * coupling_prep: folds the ActNorm affine of Conv2d (models/layers.py:223-228) and the exp(3*logs) output
//    scale of Conv2dZeros (:257-260) into bf16 GEMM operands laid out for the tcgen05 tiles; coupling_prep_bwd
//    maps GEMM-operand gradients back to the reference parameters.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"
#include "launch_util.h"

namespace nfk {

constexpr int PREP_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  // red: >= 32 floats of shared memory
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

// Gauss-Jordan with partial pivoting on A (destroyed) producing Inv = A^-1 and sum log|pivot|.
__device__ float gauss_jordan(float* A, float* Inv, int C, int* piv_row, float* red) {
  const int tid = threadIdx.x;
  for (int i = tid; i < C * C; i += blockDim.x) Inv[i] = (i / C == i % C) ? 1.f : 0.f;
  __syncthreads();
  float logabs = 0.f;
  for (int c = 0; c < C; ++c) {
    if (tid == 0) {
      int best = c;
      float bv = fabsf(A[c * C + c]);
      for (int r = c + 1; r < C; ++r) {
        const float v = fabsf(A[r * C + c]);
        if (v > bv) { bv = v; best = r; }
      }
      *piv_row = best;
    }
    __syncthreads();
    const int pr = *piv_row;
    if (pr != c) {
      for (int j = tid; j < C; j += blockDim.x) {
        float t = A[c * C + j]; A[c * C + j] = A[pr * C + j]; A[pr * C + j] = t;
        t = Inv[c * C + j]; Inv[c * C + j] = Inv[pr * C + j]; Inv[pr * C + j] = t;
      }
    }
    __syncthreads();
    const float pv = A[c * C + c];
    logabs += logf(fabsf(pv));
    __syncthreads();
    const float ip = 1.f / pv;
    for (int j = tid; j < C; j += blockDim.x) { A[c * C + j] *= ip; Inv[c * C + j] *= ip; }
    __syncthreads();
    // eliminate column c from every other row; factors are read before any write of column c (saved in red? no:
    // each (r, j) pair needs A[r][c]; process j != c first, then zero column c).
    for (int idx = tid; idx < C * C; idx += blockDim.x) {
      const int r = idx / C, j = idx % C;
      if (r == c) continue;
      const float f = A[r * C + c];
      if (j != c) A[r * C + j] -= f * A[c * C + j];
      Inv[r * C + j] -= f * Inv[c * C + j];
    }
    __syncthreads();
    for (int r = tid; r < C; r += blockDim.x)
      if (r != c) A[r * C + c] = 0.f;
    __syncthreads();
  }
  (void)red;
  return logabs;
}

struct InvconvParams {
  const float* an_bias;  // [C]
  const float* an_logs;  // [C]
  const float* lower;    // [C,C]  (LU) or nullptr
  const float* upper;    // [C,C]
  const float* log_s;    // [C]
  const float* p;        // [C,C] permutation
  const float* sign_s;   // [C]
  const float* weight;   // [C,C]  (non-LU) or nullptr
};

// smem: Lm, Um, X, Y (C*C each) + perm[C] + red[32] + piv
__global__ void __launch_bounds__(PREP_THREADS)
invconv_prep_kernel(InvconvParams q, int C, int reverse, int transpose, float* __restrict__ outW,
                    float* __restrict__ outb, float* __restrict__ out_sl) {
  extern __shared__ float sm[];
  float* Lm = sm;
  float* Um = Lm + C * C;
  float* X = Um + C * C;
  float* Y = X + C * C;
  float* red = Y + C * C;
  int* perm = reinterpret_cast<int*>(red + 32);
  int* piv = perm + C;
  const int tid = threadIdx.x;
  const int CC = C * C;

  float lsum = 0.f;
  for (int i = tid; i < C; i += blockDim.x) lsum += q.an_logs[i] + (q.weight ? 0.f : q.log_s[i]);
  lsum = block_sum(lsum, red);

  float* Wplain = X;  // forward matrix W (reference layout: z_o = sum_i W[o][i] x_i in 2-D)
  float* Winv = Y;
  if (q.weight) {
    for (int i = tid; i < CC; i += blockDim.x) { Lm[i] = q.weight[i]; Wplain[i] = q.weight[i]; }
    __syncthreads();
    if (reverse) {
      lsum += gauss_jordan(Lm, Winv, C, piv, red);
    } else {
      lsum += gauss_jordan(Lm, Um, C, piv, red);  // only the log|det| is needed
    }
  } else {
    for (int i = tid; i < CC; i += blockDim.x) {
      const int r = i / C, c = i % C;
      Lm[i] = c < r ? q.lower[i] : (c == r ? 1.f : 0.f);
      Um[i] = c > r ? q.upper[i] : (c == r ? q.sign_s[r] * expf(q.log_s[r]) : 0.f);
    }
    for (int r = tid; r < C; r += blockDim.x) {
      int best = 0;
      float bv = q.p[r * C];
      for (int c = 1; c < C; ++c)
        if (q.p[r * C + c] > bv) { bv = q.p[r * C + c]; best = c; }
      perm[r] = best;
    }
    __syncthreads();
    if (!reverse) {
      for (int i = tid; i < CC; i += blockDim.x) {
        const int o = i / C, c = i % C, r = perm[o];
        const int kmax = min(r, c);
        float t = 0.f;
        for (int k = 0; k <= kmax; ++k) t = fmaf(Lm[r * C + k], Um[k * C + c], t);
        Wplain[i] = t;
      }
    } else {
      // X <- L^-1 (forward substitution), Y <- U^-1 (back substitution), one thread per column
      for (int j = tid; j < C; j += blockDim.x) {
        for (int i = 0; i < C; ++i) {
          float x;
          if (i < j) x = 0.f;
          else if (i == j) x = 1.f;
          else {
            x = 0.f;
            for (int k = j; k < i; ++k) x = fmaf(Lm[i * C + k], X[k * C + j], x);
            x = -x;
          }
          X[i * C + j] = x;
        }
        for (int i = C - 1; i >= 0; --i) {
          float x;
          if (i > j) x = 0.f;
          else if (i == j) x = 1.f / Um[j * C + j];
          else {
            x = 0.f;
            for (int k = i + 1; k <= j; ++k) x = fmaf(Um[i * C + k], Y[k * C + j], x);
            x = -x / Um[i * C + i];
          }
          Y[i * C + j] = x;
        }
      }
      __syncthreads();
      // Lm <- (U^-1 L^-1) P^T : column c of the product goes to column where perm[.] == c
      for (int i = tid; i < CC; i += blockDim.x) {
        const int r = i / C, c = i % C, pc = perm[c];
        float t = 0.f;
        for (int k = max(r, pc); k < C; ++k) t = fmaf(Y[r * C + k], X[k * C + pc], t);
        Lm[i] = t;
      }
      __syncthreads();
      for (int i = tid; i < CC; i += blockDim.x) Winv[i] = Lm[i];
    }
  }
  __syncthreads();
  // fused affine in "out" layout: out[a][b] multiplies input channel b into output channel a
  float* Out = Um;  // reuse
  __syncthreads();
  for (int i = tid; i < CC; i += blockDim.x) {
    const int a = i / C, b = i % C;
    float v;
    if (!reverse) v = (transpose ? Wplain[b * C + a] : Wplain[i]) * expf(q.an_logs[b]);
    else v = (transpose ? Winv[b * C + a] : Winv[i]) * expf(-q.an_logs[a]);
    Out[i] = v;
    outW[i] = v;
  }
  __syncthreads();
  for (int a = tid; a < C; a += blockDim.x) {
    float t;
    if (!reverse) {
      t = 0.f;
      for (int b = 0; b < C; ++b) t = fmaf(Out[a * C + b], q.an_bias[b], t);
    } else {
      t = -q.an_bias[a];
    }
    outb[a] = t;
  }
  if (tid == 0) out_sl[0] = reverse ? -lsum : lsum;
}

// Backward of the prep (either direction). Wf is the saved outW of the same direction. Outputs are overwritten.
__global__ void __launch_bounds__(PREP_THREADS)
invconv_prep_bwd_kernel(InvconvParams q, int C, int reverse, int transpose, const float* __restrict__ Wf,
                        const float* __restrict__ dWf, const float* __restrict__ dbf,
                        const float* __restrict__ g_ld, int B, float pixels, float* __restrict__ d_bias,
                        float* __restrict__ d_logs, float* __restrict__ d_lower, float* __restrict__ d_upper,
                        float* __restrict__ d_log_s, float* __restrict__ d_weight) {
  extern __shared__ float sm[];
  float* Lm = sm;
  float* Um = Lm + C * C;
  float* G = Um + C * C;   // later dT
  float* dW = G + C * C;
  float* red = dW + C * C;
  int* perm = reinterpret_cast<int*>(red + 32);
  int* piv = perm + C;
  const int tid = threadIdx.x;
  const int CC = C * C;

  float gs = 0.f;
  if (g_ld)
    for (int b = tid; b < B; b += blockDim.x) gs += g_ld[b];
  float gsum = block_sum(gs, red) * pixels;

  if (!reverse) {
    // Wf[a][b] = W*[a][b] e^{logs_b}, bf = Wf bias
    for (int i = tid; i < CC; i += blockDim.x) {
      const int a = i / C, b = i % C;
      G[i] = dWf[i] + dbf[a] * q.an_bias[b];
    }
    __syncthreads();
    for (int b = tid; b < C; b += blockDim.x) {
      float db = 0.f, dl = 0.f;
      for (int a = 0; a < C; ++a) {
        db = fmaf(Wf[a * C + b], dbf[a], db);
        dl = fmaf(G[a * C + b], Wf[a * C + b], dl);
      }
      d_bias[b] = db;
      d_logs[b] = dl + gsum;
    }
    for (int i = tid; i < CC; i += blockDim.x) {
      const int r = i / C, c = i % C;
      dW[i] = transpose ? G[c * C + r] * expf(q.an_logs[r]) : G[i] * expf(q.an_logs[c]);
    }
    __syncthreads();
  } else {
    // Wf[a][b] = Winv*[a][b] e^{-logs_a}, bf = -bias, log-det enters with a minus sign
    gsum = -gsum;
    for (int a = tid; a < C; a += blockDim.x) {
      float dl = 0.f;
      for (int b = 0; b < C; ++b) dl = fmaf(dWf[a * C + b], Wf[a * C + b], dl);
      d_bias[a] = -dbf[a];
      d_logs[a] = -dl + gsum;
    }
    float* Winv = G;      // plain layout
    float* dWinv = dW;    // plain layout
    for (int i = tid; i < CC; i += blockDim.x) {
      const int a = i / C, b = i % C;
      const float e = expf(q.an_logs[a]);
      const int pi = transpose ? b * C + a : i;
      Winv[pi] = Wf[i] * e;
      dWinv[pi] = dWf[i] / e;
    }
    __syncthreads();
    // Lm <- dWinv Winv^T ; Um <- -Winv^T Lm  (= dL/dW)
    for (int i = tid; i < CC; i += blockDim.x) {
      const int r = i / C, c = i % C;
      float t = 0.f;
      for (int k = 0; k < C; ++k) t = fmaf(dWinv[r * C + k], Winv[c * C + k], t);
      Lm[i] = t;
    }
    __syncthreads();
    for (int i = tid; i < CC; i += blockDim.x) {
      const int r = i / C, c = i % C;
      float t = 0.f;
      for (int k = 0; k < C; ++k) t = fmaf(Winv[k * C + r], Lm[k * C + c], t);
      Um[i] = -t;
    }
    __syncthreads();
    for (int i = tid; i < CC; i += blockDim.x) dW[i] = Um[i];
    __syncthreads();
  }
  if (q.weight) {
    // d slogdet / dW = W^-T
    for (int i = tid; i < CC; i += blockDim.x) Lm[i] = q.weight[i];
    __syncthreads();
    gauss_jordan(Lm, Um, C, piv, red);
    for (int i = tid; i < CC; i += blockDim.x) {
      const int r = i / C, c = i % C;
      d_weight[i] = dW[i] + gsum * Um[c * C + r];
    }
    return;
  }
  for (int i = tid; i < CC; i += blockDim.x) {
    const int r = i / C, c = i % C;
    Lm[i] = c < r ? q.lower[i] : (c == r ? 1.f : 0.f);
    Um[i] = c > r ? q.upper[i] : (c == r ? q.sign_s[r] * expf(q.log_s[r]) : 0.f);
  }
  for (int r = tid; r < C; r += blockDim.x) {
    int best = 0;
    float bv = q.p[r * C];
    for (int c = 1; c < C; ++c)
      if (q.p[r * C + c] > bv) { bv = q.p[r * C + c]; best = c; }
    perm[r] = best;
  }
  __syncthreads();
  float* dT = G;  // dT[perm[o]][i] = dW[o][i]
  for (int i = tid; i < CC; i += blockDim.x) {
    const int o = i / C, c = i % C;
    dT[perm[o] * C + c] = dW[i];
  }
  __syncthreads();
  for (int i = tid; i < CC; i += blockDim.x) {
    const int r = i / C, k = i % C;
    // d_lower[r][k] = sum_{c>=k} dT[r][c] Um[k][c]   (k < r)
    float dl = 0.f;
    if (k < r)
      for (int c = k; c < C; ++c) dl = fmaf(dT[r * C + c], Um[k * C + c], dl);
    d_lower[i] = dl;
    // dUm[r][k] = sum_{rr>=r} Lm[rr][r] dT[rr][k]   (here (r,k) indexes U)
    float du = 0.f;
    if (k >= r)
      for (int rr = r; rr < C; ++rr) du = fmaf(Lm[rr * C + r], dT[rr * C + k], du);
    d_upper[i] = k > r ? du : 0.f;
    if (k == r) d_log_s[r] = du * Um[r * C + r] + gsum;
  }
}

// ------------------------------------------------------------------------------------------ coupling weights
struct CouplingW {
  const float *w1, *b1, *l1;  // Conv2d #1: [hid, cin, 3, 3], actnorm bias/logs [hid]
  const float *w2, *b2, *l2;  // Conv2d #2: [hid, hid, 1, 1]
  const float *w3, *b3, *l3;  // Conv2dZeros: [cout, hid, 3, 3], bias [cout], logs [cout]
};
struct CouplingOps {
  __nv_bfloat16 *B1, *B1T;  // [hid, K1p], [K1p, hid]     k = tap*cin + ci
  __nv_bfloat16 *B2, *B2T;  // [hid, hid] (co, ci), (ci, co)
  __nv_bfloat16 *B3, *B3T;  // [K3p, hid] rows n = tap*cout + co ; [hid, K3p]
  float *bias1, *bias2, *bias3;  // folded biases
};

__global__ void coupling_prep_kernel(CouplingW w, CouplingOps o, int cin, int hid, int cout, int K1p, int K3p,
                                     int with_transposed) {
  const long long n1 = static_cast<long long>(hid) * K1p;
  const long long n2 = static_cast<long long>(hid) * hid;
  const long long n3 = static_cast<long long>(K3p) * hid;
  const long long total = n1 + n2 + n3 + 2 * hid + cout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < n1) {
      const int co = static_cast<int>(i / K1p), k = static_cast<int>(i % K1p);
      float v = 0.f;
      if (k < 9 * cin) {
        const int tap = k / cin, ci = k % cin;
        v = w.w1[(static_cast<long long>(co) * cin + ci) * 9 + tap] * expf(w.l1[co]);
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      o.B1[i] = h;
      if (with_transposed) o.B1T[static_cast<long long>(k) * hid + co] = h;
    } else if (i < n1 + n2) {
      const long long j = i - n1;
      const int co = static_cast<int>(j / hid), ci = static_cast<int>(j % hid);
      const __nv_bfloat16 h = __float2bfloat16_rn(w.w2[j] * expf(w.l2[co]));
      o.B2[j] = h;
      if (with_transposed) o.B2T[static_cast<long long>(ci) * hid + co] = h;
    } else if (i < n1 + n2 + n3) {
      const long long j = i - n1 - n2;
      const int n = static_cast<int>(j / hid), ci = static_cast<int>(j % hid);
      float v = 0.f;
      if (n < 9 * cout) {
        const int tap = n / cout, co = n % cout;
        v = w.w3[(static_cast<long long>(co) * hid + ci) * 9 + tap] * expf(3.f * w.l3[co]);
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      o.B3[j] = h;
      if (with_transposed) o.B3T[static_cast<long long>(ci) * K3p + n] = h;
    } else {
      const int j = static_cast<int>(i - n1 - n2 - n3);
      if (j < hid) o.bias1[j] = w.b1[j] * expf(w.l1[j]);
      else if (j < 2 * hid) o.bias2[j - hid] = w.b2[j - hid] * expf(w.l2[j - hid]);
      else o.bias3[j - 2 * hid] = w.b3[j - 2 * hid] * expf(3.f * w.l3[j - 2 * hid]);
    }
  }
}

struct CouplingGradIn {
  const float *dB1, *dbias1;  // [hid, K1p], [hid]
  const float *dB2, *dbias2;  // [hid, hid], [hid]
  const float *dB3, *dbias3;  // [K3p, hid], [cout]
};
struct CouplingGradOut {
  float *dw1, *db1, *dl1, *dw2, *db2, *dl2, *dw3, *db3, *dl3;
};

// One CTA per output channel: rows [0,hid) -> layer 1, [hid,2hid) -> layer 2, [2hid, 2hid+cout) -> layer 3.
__global__ void __launch_bounds__(PREP_THREADS)
coupling_prep_bwd_kernel(CouplingW w, CouplingGradIn gi, CouplingGradOut go, int cin, int hid, int cout, int K1p,
                         int K3p) {
  __shared__ float red[32];
  const int row = blockIdx.x, tid = threadIdx.x;
  float acc = 0.f;
  if (row < hid) {
    const int co = row;
    const float e = expf(w.l1[co]);
    for (int k = tid; k < 9 * cin; k += blockDim.x) {
      const int tap = k / cin, ci = k % cin;
      const long long wi = (static_cast<long long>(co) * cin + ci) * 9 + tap;
      const float g = gi.dB1[static_cast<long long>(co) * K1p + k];
      go.dw1[wi] = g * e;
      acc = fmaf(g, w.w1[wi] * e, acc);
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
      const float gb = gi.dbias1[co];
      go.db1[co] = gb * e;
      go.dl1[co] = acc + gb * w.b1[co] * e;
    }
  } else if (row < 2 * hid) {
    const int co = row - hid;
    const float e = expf(w.l2[co]);
    for (int ci = tid; ci < hid; ci += blockDim.x) {
      const long long wi = static_cast<long long>(co) * hid + ci;
      const float g = gi.dB2[wi];
      go.dw2[wi] = g * e;
      acc = fmaf(g, w.w2[wi] * e, acc);
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
      const float gb = gi.dbias2[co];
      go.db2[co] = gb * e;
      go.dl2[co] = acc + gb * w.b2[co] * e;
    }
  } else {
    const int co = row - 2 * hid;
    const float e = expf(3.f * w.l3[co]);
    for (int j = tid; j < 9 * hid; j += blockDim.x) {
      const int tap = j / hid, ci = j % hid;
      const long long wi = (static_cast<long long>(co) * hid + ci) * 9 + tap;
      const float g = gi.dB3[static_cast<long long>(tap * cout + co) * hid + ci];
      go.dw3[wi] = g * e;
      acc = fmaf(g, w.w3[wi] * e, acc);
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
      const float gb = gi.dbias3[co];
      go.db3[co] = gb * e;
      go.dl3[co] = 3.f * (acc + gb * w.b3[co] * e);
    }
  }
}

static int prep_smem(int C) { return (4 * C * C + 32 + C + 4) * static_cast<int>(sizeof(float)); }

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_invconv_prep(const float* an_bias, const float* an_logs, const float* lower, const float* upper,
                                const float* log_s, const float* p, const float* sign_s, const float* weight,
                                int C, int reverse, int transpose, float* outW, float* outb, float* out_sl,
                                void* stream) {
  if (C <= 0 || C > 104) return NFK_ERR_SHAPE;
  if (!an_bias || !an_logs || !outW || !outb || !out_sl) return NFK_ERR_ARG;
  if (!weight && (!lower || !upper || !log_s || !p || !sign_s)) return NFK_ERR_ARG;
  InvconvParams q{an_bias, an_logs, lower, upper, log_s, p, sign_s, weight};
  const int smem = prep_smem(C);
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(invconv_prep_kernel), smem)) return rc;
  invconv_prep_kernel<<<1, PREP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(q, C, reverse, transpose, outW,
                                                                                  outb, out_sl);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_invconv_prep_bwd(const float* an_bias, const float* an_logs, const float* lower,
                                    const float* upper, const float* log_s, const float* p, const float* sign_s,
                                    const float* weight, int C, int reverse, int transpose, const float* Wf,
                                    const float* dWf, const float* dbf, const float* g_ld, int B, float pixels,
                                    float* d_bias,
                                    float* d_logs, float* d_lower, float* d_upper, float* d_log_s, float* d_weight,
                                    void* stream) {
  if (C <= 0 || C > 104 || B <= 0) return NFK_ERR_SHAPE;
  if (!Wf || !dWf || !dbf || !d_bias || !d_logs) return NFK_ERR_ARG;
  if (weight ? !d_weight : (!d_lower || !d_upper || !d_log_s)) return NFK_ERR_ARG;
  InvconvParams q{an_bias, an_logs, lower, upper, log_s, p, sign_s, weight};
  const int smem = prep_smem(C);
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(invconv_prep_bwd_kernel), smem)) return rc;
  invconv_prep_bwd_kernel<<<1, PREP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      q, C, reverse, transpose, Wf, dWf, dbf, g_ld, B, pixels, d_bias, d_logs, d_lower, d_upper, d_log_s, d_weight);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_coupling_prep(const float* w1, const float* b1, const float* l1, const float* w2, const float* b2,
                                 const float* l2, const float* w3, const float* b3, const float* l3, int cin, int hid,
                                 int cout, int K1p, int K3p, void* B1, void* B1T, void* B2, void* B2T, void* B3,
                                 void* B3T, float* bias1, float* bias2, float* bias3, int with_transposed,
                                 void* stream) {
  if (cin <= 0 || hid <= 0 || cout <= 0 || hid % 64 || K1p % 64 || K3p % 64 || K1p < 9 * cin || K3p < 9 * cout)
    return NFK_ERR_SHAPE;
  if (with_transposed && (!B1T || !B2T || !B3T)) return NFK_ERR_ARG;
  CouplingW w{w1, b1, l1, w2, b2, l2, w3, b3, l3};
  CouplingOps o{static_cast<__nv_bfloat16*>(B1), static_cast<__nv_bfloat16*>(B1T), static_cast<__nv_bfloat16*>(B2),
                static_cast<__nv_bfloat16*>(B2T), static_cast<__nv_bfloat16*>(B3), static_cast<__nv_bfloat16*>(B3T),
                bias1, bias2, bias3};
  const long long total = static_cast<long long>(hid) * (K1p + hid + K3p) + 2 * hid + cout;
  const int blocks = static_cast<int>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  coupling_prep_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, o, cin, hid, cout, K1p, K3p,
                                                                             with_transposed);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_coupling_prep_bwd(const float* w1, const float* b1, const float* l1, const float* w2,
                                     const float* b2, const float* l2, const float* w3, const float* b3,
                                     const float* l3, int cin, int hid, int cout, int K1p, int K3p, const float* dB1,
                                     const float* dbias1, const float* dB2, const float* dbias2, const float* dB3,
                                     const float* dbias3, float* dw1, float* db1, float* dl1, float* dw2, float* db2,
                                     float* dl2, float* dw3, float* db3, float* dl3, void* stream) {
  if (cin <= 0 || hid <= 0 || cout <= 0) return NFK_ERR_SHAPE;
  CouplingW w{w1, b1, l1, w2, b2, l2, w3, b3, l3};
  CouplingGradIn gi{dB1, dbias1, dB2, dbias2, dB3, dbias3};
  CouplingGradOut go{dw1, db1, dl1, dw2, db2, dl2, dw3, db3, dl3};
  coupling_prep_bwd_kernel<<<2 * hid + cout, PREP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      w, gi, go, cin, hid, cout, K1p, K3p);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
