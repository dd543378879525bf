// The tail of the KD training step as four launches (reference: pl_module.py:257-320 NFModel.loss, :348-363
// configure_optimizers, train.py:46 gradient_clip_val=30):
//
//   kd_nll_loss_fwd   ONE launch: per sample, the multi-level latent MSE (pl_module.py:266-282), the prior
//                     log-density of the last latent and the bits/dim objective (models/layers.py:10-23,
//                     models/kd_flows.py:134-150), the weighted sum of the terms (pl_module.py:306-313) — and the four
//                     batch means the step returns (:315-320), summed in a fixed order by the last CTA to finish.
//   kd_nll_loss_bwd   ONE launch: gradients of all of the above w.r.t. every student tap, the last latent and the
//                     log-det.
//   grad_sqnorm       squared L2 norm of the flat gradient buffer (per-CTA partial sums, fixed order) + step counter.
//   adam_step         clip_grad_norm_(max_norm) folded into Adam / Adamax on the flat parameter buffer.
//
// All four are HBM-bound streaming kernels: 128-bit loads/stores, grid sized to the work (one CTA or one warp per
// sample for the loss, 4 CTAs per SM for the optimiser).
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/nfk.h"

namespace nfk {

constexpr int LT = 256;
constexpr float kLog2PiL = 1.8378770664093453f;

struct LossLevels {
  const float* s[NFK_LOSS_MAX_LEVELS];
  const float* t[NFK_LOSS_MAX_LEVELS];
  float* ds[NFK_LOSS_MAX_LEVELS];
  int n[NFK_LOSS_MAX_LEVELS];
  int L;
};

struct LossArgs {
  const float* z_last;      // [B, nz] or null (then nll_in holds the per-sample objective already)
  const float* prior_mean;  // [nz] or null (zeros)
  const float* prior_logs;  // [nz] or null (zeros)
  const float* logdet;      // [B]
  const float* nll_in;      // [B] or null
  const float* perc;        // [B] or null
  const float* sample_w;    // [B] or null
  int nz, B;
  float nll_scale, w_nll, w_kd, w_perc;
};

// sum over the GROUP threads that share a sample (GROUP == 32: one warp; GROUP == LT: the whole CTA)
template <int GROUP>
__device__ __forceinline__ float group_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (GROUP == 32) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < LT / 32; ++i) t += red[i];
  return t;
}

template <int GROUP>
__device__ __forceinline__ float sq_diff_sum(const float* __restrict__ sp, const float* __restrict__ tp, int n, int l) {
  float a = 0.f;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(tp)) & 15) == 0) {
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
  return a;
}

// scratch: [grid][4] partial sums, then one unsigned counter (zeroed by the host wrapper before the launch)
template <int GROUP>
__global__ void __launch_bounds__(LT)
kd_nll_loss_fwd_kernel(const __grid_constant__ LossLevels lv, const __grid_constant__ LossArgs a,
                       float* __restrict__ nll_out, float* __restrict__ kd_out, float* __restrict__ means,
                       float* __restrict__ scratch) {
  __shared__ float red[LT / 32];
  __shared__ float part[LT / GROUP][4];
  __shared__ bool last;
  constexpr int PER = LT / GROUP;
  const int g = threadIdx.x / GROUP, l = threadIdx.x % GROUP;
  const int b = blockIdx.x * PER + g;
  const bool live = b < a.B;
  const int bb = live ? b : 0;
  // ---- latent MSE over the levels: kd[b] = (1/L) sum_l mean_i (s - t)^2
  float kd = 0.f;
  for (int k = 0; k < lv.L; ++k) {
    const int n = lv.n[k];
    float v = live ? sq_diff_sum<GROUP>(lv.s[k] + static_cast<long long>(bb) * n,
                                        lv.t[k] + static_cast<long long>(bb) * n, n, l) : 0.f;
    v = group_sum<GROUP>(v, red);
    kd += v / (static_cast<float>(n) * static_cast<float>(lv.L));
  }
  // ---- objective: nll[b] = -(logdet[b] + sum_i log N(z_i; mean_i, exp(logs_i))) * scale
  float nll;
  if (a.z_last) {
    float acc = 0.f;
    if (live) {
      const float* zp = a.z_last + static_cast<long long>(bb) * a.nz;
      for (int i = l; i < a.nz; i += GROUP) {
        const float lg = a.prior_logs ? a.prior_logs[i] : 0.f;
        const float d = zp[i] - (a.prior_mean ? a.prior_mean[i] : 0.f);
        acc += -0.5f * (2.f * lg + d * d * expf(-2.f * lg) + kLog2PiL);
      }
    }
    acc = group_sum<GROUP>(acc, red);
    nll = -(a.logdet[bb] + acc) * a.nll_scale;
  } else {
    nll = a.nll_in[bb];
  }
  if (l == 0) {
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;
    if (live) {
      const float pc = a.perc ? a.perc[b] : 0.f;
      float res = a.w_nll * nll + a.w_kd * kd + a.w_perc * pc;
      if (a.sample_w) res *= a.sample_w[b];
      if (nll_out) nll_out[b] = nll;
      if (kd_out) kd_out[b] = kd;
      r0 = nll; r1 = kd; r2 = pc; r3 = res;
    }
    part[g][0] = r0; part[g][1] = r1; part[g][2] = r2; part[g][3] = r3;
  }
  __syncthreads();
  // ---- batch means: per-CTA partials, the last CTA to arrive adds them up in CTA order (deterministic)
  if (threadIdx.x < 4) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) v += part[i][threadIdx.x];
    scratch[blockIdx.x * 4 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  unsigned* counter = reinterpret_cast<unsigned*>(scratch + static_cast<size_t>(gridDim.x) * 4);
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int comp = threadIdx.x >> 6, t64 = threadIdx.x & 63;       // 4 components x 64 threads
  float v = 0.f;
  for (int i = t64; i < static_cast<int>(gridDim.x); i += 64) v += __ldcg(scratch + i * 4 + comp);
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ float fin[8];
  if ((threadIdx.x & 31) == 0) fin[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 4) means[threadIdx.x] = (fin[2 * threadIdx.x] + fin[2 * threadIdx.x + 1]) / static_cast<float>(a.B);
}

struct LossBwdArgs {
  const float* g_means;   // [4] gradient of the four batch means (device)
  const float* g_nll;     // [B] or null: extra gradient on the per-sample objective
  const float* g_kd;      // [B] or null
  float* dz_last;         // [B, nz] or null
  float* dlogdet;         // [B] or null
  float* dnll_in;         // [B] or null (when the objective came in precomputed)
  float* dperc;           // [B] or null
};

template <int GROUP>
__global__ void __launch_bounds__(LT)
kd_nll_loss_bwd_kernel(const __grid_constant__ LossLevels lv, const __grid_constant__ LossArgs a,
                       const __grid_constant__ LossBwdArgs o) {
  constexpr int PER = LT / GROUP;
  const int g = threadIdx.x / GROUP, l = threadIdx.x % GROUP;
  const int b = blockIdx.x * PER + g;
  if (b >= a.B) return;
  const float inv_b = 1.f / static_cast<float>(a.B);
  const float sw = a.sample_w ? a.sample_w[b] : 1.f;
  const float gl = o.g_means[3] * sw;
  const float c_nll = (o.g_means[0] + gl * a.w_nll) * inv_b + (o.g_nll ? o.g_nll[b] : 0.f);
  const float c_kd = (o.g_means[1] + gl * a.w_kd) * inv_b + (o.g_kd ? o.g_kd[b] : 0.f);
  for (int k = 0; k < lv.L; ++k) {
    if (!lv.ds[k]) continue;
    const int n = lv.n[k];
    const float c = 2.f * c_kd / (static_cast<float>(n) * static_cast<float>(lv.L));
    const float* sp = lv.s[k] + static_cast<long long>(b) * n;
    const float* tp = lv.t[k] + static_cast<long long>(b) * n;
    float* dp = lv.ds[k] + static_cast<long long>(b) * n;
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(sp) | reinterpret_cast<uintptr_t>(tp) |
                          reinterpret_cast<uintptr_t>(dp)) & 15) == 0) {
      for (int i = l; i < n / 4; i += GROUP) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(sp) + i);
        const float4 v = __ldg(reinterpret_cast<const float4*>(tp) + i);
        reinterpret_cast<float4*>(dp)[i] = make_float4(c * (u.x - v.x), c * (u.y - v.y), c * (u.z - v.z), c * (u.w - v.w));
      }
    } else {
      for (int i = l; i < n; i += GROUP) dp[i] = c * (sp[i] - tp[i]);
    }
  }
  if (a.z_last) {
    const float gs = c_nll * a.nll_scale;
    if (o.dz_last) {
      const float* zp = a.z_last + static_cast<long long>(b) * a.nz;
      float* dz = o.dz_last + static_cast<long long>(b) * a.nz;
      for (int i = l; i < a.nz; i += GROUP) {
        const float lg = a.prior_logs ? a.prior_logs[i] : 0.f;
        dz[i] = gs * (zp[i] - (a.prior_mean ? a.prior_mean[i] : 0.f)) * expf(-2.f * lg);
      }
    }
    if (l == 0 && o.dlogdet) o.dlogdet[b] = -gs;
  } else if (l == 0 && o.dnll_in) {
    o.dnll_in[b] = c_nll;
  }
  if (l == 0 && o.dperc) o.dperc[b] = (o.g_means[2] + gl * a.w_perc) * inv_b;
}

// ------------------------------------------------------------------------------------------------ clip + Adam
constexpr int OPT_BLOCKS = 592;   // 4 CTAs per SM

__global__ void __launch_bounds__(LT)
grad_sqnorm_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials, int* __restrict__ step) {
  __shared__ float red[LT / 32];
  float a = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(LT) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * LT) {
    const float4 v = __ldg(g4 + i);
    a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += LT) a += g[i] * g[i];
  a = group_sum<LT>(a, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = a;
    if (blockIdx.x == 0 && step) *step += 1;
  }
}

struct AdamArgs {
  float max_norm, lr, beta1, beta2, eps, weight_decay;
  int adamax, nparts;
};

__global__ void __launch_bounds__(LT)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, const float* __restrict__ partials, const int* __restrict__ step, float* __restrict__ norm_out,
                 const AdamArgs a) {
  __shared__ float red[LT / 32];
  // total gradient norm: every CTA adds the same partial sums in the same order
  float s = 0.f;
  for (int i = threadIdx.x; i < a.nparts; i += LT) s += partials[i];
  s = group_sum<LT>(s, red);
  const float norm = sqrtf(s);
  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
  const float coef = a.max_norm > 0.f ? fminf(1.f, a.max_norm / (norm + 1e-6f)) : 1.f;
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = norm;
  const double t = static_cast<double>(*step);
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(a.beta1), t));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(a.beta2), t));
  const float step_size = a.lr / bc1;
  const float sqrt_bc2 = sqrtf(bc2);
  auto upd = [&](float& pw, float gw, float& mw, float& vw) {
    gw *= coef;
    if (a.weight_decay != 0.f) gw = fmaf(a.weight_decay, pw, gw);
    mw = mw + (gw - mw) * (1.f - a.beta1);                         // lerp, as torch.optim.Adam
    if (a.adamax) {                                                 // torch.optim.Adamax
      vw = fmaxf(a.beta2 * vw, fabsf(gw) + a.eps);
      pw -= step_size * mw / vw;
    } else {
      vw = a.beta2 * vw + (1.f - a.beta2) * gw * gw;
      pw -= step_size * mw / (sqrtf(vw) / sqrt_bc2 + a.eps);
    }
  };
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(LT) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * LT) {
    float4 pw = reinterpret_cast<float4*>(p)[i];
    const float4 gw = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mw = reinterpret_cast<float4*>(m)[i], vw = reinterpret_cast<float4*>(v)[i];
    upd(pw.x, gw.x, mw.x, vw.x); upd(pw.y, gw.y, mw.y, vw.y); upd(pw.z, gw.z, mw.z, vw.z); upd(pw.w, gw.w, mw.w, vw.w);
    reinterpret_cast<float4*>(p)[i] = pw;
    reinterpret_cast<float4*>(m)[i] = mw;
    reinterpret_cast<float4*>(v)[i] = vw;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += LT) upd(p[i], g[i], m[i], v[i]);
}

static int fill_levels(const nfk_loss_levels* lv, LossLevels& out, bool bwd) {
  if (!lv || lv->L < 0 || lv->L > NFK_LOSS_MAX_LEVELS) return NFK_ERR_ARG;
  out.L = lv->L;
  for (int k = 0; k < NFK_LOSS_MAX_LEVELS; ++k) {
    out.s[k] = k < lv->L ? lv->s[k] : nullptr;
    out.t[k] = k < lv->L ? lv->t[k] : nullptr;
    out.ds[k] = (k < lv->L && bwd) ? lv->ds[k] : nullptr;
    out.n[k] = k < lv->L ? lv->n[k] : 0;
    if (k < lv->L && (!lv->s[k] || !lv->t[k] || lv->n[k] <= 0)) return NFK_ERR_ARG;
  }
  return NFK_OK;
}

static long long per_sample(const LossLevels& lv, int nz) {
  long long t = nz;
  for (int k = 0; k < lv.L; ++k) t += lv.n[k];
  return t;
}

}  // namespace nfk

using namespace nfk;

extern "C" int nfk_kd_nll_loss_scratch_floats(int B) {
  // per-CTA partial sums of the widest grid (one warp per sample: B / 8 CTAs; one CTA per sample: B) + the counter
  return B <= 0 ? 0 : 4 * B + 4;
}

extern "C" int nfk_kd_nll_loss_fwd(const nfk_loss_levels* levels, const float* z_last, int nz, const float* prior_mean,
                                   const float* prior_logs, const float* logdet, float nll_scale, const float* nll_in,
                                   const float* perc, const float* sample_w, float w_nll, float w_kd, float w_perc,
                                   int B, float* nll_out, float* kd_out, float* means, float* scratch, void* stream) {
  if (B <= 0 || nz < 0) return NFK_ERR_SHAPE;
  LossLevels lv;
  if (int rc = fill_levels(levels, lv, false)) return rc;
  if (!means || !scratch) return NFK_ERR_ARG;
  if (z_last ? (!logdet || nz <= 0) : !nll_in) return NFK_ERR_ARG;
  LossArgs a{z_last, prior_mean, prior_logs, logdet, nll_in, perc, sample_w, nz, B, nll_scale, w_nll, w_kd, w_perc};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool wide = per_sample(lv, z_last ? nz : 0) >= 2048;
  const int grid = wide ? B : (B + 7) / 8;
  if (cudaMemsetAsync(scratch + static_cast<size_t>(grid) * 4, 0, sizeof(unsigned), st) != cudaSuccess)
    return NFK_ERR_LAUNCH;
  if (wide) kd_nll_loss_fwd_kernel<LT><<<grid, LT, 0, st>>>(lv, a, nll_out, kd_out, means, scratch);
  else kd_nll_loss_fwd_kernel<32><<<grid, LT, 0, st>>>(lv, a, nll_out, kd_out, means, scratch);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_kd_nll_loss_bwd(const nfk_loss_levels* levels, const float* z_last, int nz, const float* prior_mean,
                                   const float* prior_logs, float nll_scale, const float* sample_w, float w_nll,
                                   float w_kd, float w_perc, int B, const float* g_means, const float* g_nll,
                                   const float* g_kd, float* dz_last, float* dlogdet, float* dnll_in, float* dperc,
                                   void* stream) {
  if (B <= 0 || nz < 0) return NFK_ERR_SHAPE;
  LossLevels lv;
  if (int rc = fill_levels(levels, lv, true)) return rc;
  if (!g_means) return NFK_ERR_ARG;
  LossArgs a{z_last, prior_mean, prior_logs, nullptr, nullptr, nullptr, sample_w, nz, B, nll_scale, w_nll, w_kd, w_perc};
  LossBwdArgs o{g_means, g_nll, g_kd, dz_last, dlogdet, dnll_in, dperc};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool wide = per_sample(lv, z_last ? nz : 0) >= 2048;
  if (wide) kd_nll_loss_bwd_kernel<LT><<<B, LT, 0, st>>>(lv, a, o);
  else kd_nll_loss_bwd_kernel<32><<<(B + 7) / 8, LT, 0, st>>>(lv, a, o);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_optim_partials(void) { return OPT_BLOCKS; }

extern "C" int nfk_grad_sqnorm(const float* g, long long n, float* partials, int* step, void* stream) {
  if (n <= 0) return NFK_ERR_SHAPE;
  if (!g || !partials) return NFK_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(g) & 15) return NFK_ERR_ALIGN;
  grad_sqnorm_kernel<<<OPT_BLOCKS, LT, 0, static_cast<cudaStream_t>(stream)>>>(g, n, partials, step);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}

extern "C" int nfk_adam_step(float* p, const float* g, float* m, float* v, long long n, const float* partials,
                             const int* step, float max_norm, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int adamax, float* norm_out, void* stream) {
  if (n <= 0) return NFK_ERR_SHAPE;
  if (!p || !g || !m || !v || !partials || !step) return NFK_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return NFK_ERR_ALIGN;
  AdamArgs a{max_norm, lr, beta1, beta2, eps, weight_decay, adamax, OPT_BLOCKS};
  adam_step_kernel<<<OPT_BLOCKS, LT, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, partials, step, norm_out, a);
  return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ERR_LAUNCH;
}
