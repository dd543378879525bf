/* libnfk — C-ABI of the B200 (sm_100a) normalizing-flow kernels.
 *
 * Drop-in boundary for the flow hot path of vklyukin/nf_distillation (pure-Python/PyTorch reference):
 *   models/layers.py  (ActNorm, InvertibleConv1x1, Conv2d, Conv2dZeros, Split2d, gaussian_*)
 *   models/flows.py   (FlowStep.normal_flow / reverse_flow, get_block_2d / get_block_1d)
 *   pl_module.py      (NFModel.loss: multi-level latent MSE KD)
 * Each entry point names the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers + sizes, `void* stream` is a cudaStream_t; no torch types.
 *   - the caller owns and allocates every buffer; nothing here allocates, synchronises or keeps global state.
 *   - every function returns NFK_OK (0) or a negative NFK_ERR_* code; kernels are stream-ordered and re-entrant.
 *   - fp32 tensors are contiguous NCHW ("z path"); coupling-network activations are pixel-major bf16
 *     matrices [pixels, channels] ("h path"), pixels = B*H*W in (b, y, x) order.
 */
#ifndef NFK_H_
#define NFK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define NFK_OK 0
#define NFK_ERR_SHAPE (-1)  /* unsupported / inconsistent sizes */
#define NFK_ERR_ALIGN (-2)  /* pointer or leading dimension not 16-byte aligned */
#define NFK_ERR_ARG (-3)    /* missing / invalid argument */
#define NFK_ERR_LAUNCH (-4) /* CUDA launch failure */
#define NFK_ERR_DRIVER (-5) /* tensor-map encode / driver entry point failure */

/* Epilogues of nfk_gemm_nt_bf16 */
#define NFK_EPI_F32 0            /* out fp32  = acc (+ bias[col])                                   */
#define NFK_EPI_BIAS_RELU_BF16 1 /* out bf16  = relu(acc + bias[col])   (Conv2d+ActNorm+ReLU, folded) */
#define NFK_EPI_MASK_BF16 2      /* out bf16  = acc * (aux > 0); colsum[col] += column sums (ReLU bwd) */

int nfk_version(void);

/* ---- tensor-core GEMM tiles (tcgen05 + TMEM + TMA) --------------------------------------------------------
 * out[M,N] = A[M,K] * B[N,K]^T, A/B bf16 row-major (K contiguous), fp32 accumulate in TMEM.
 * Replaces the cuDNN/cuBLAS calls behind nn.Conv2d in models/layers.py:209,226 (Conv2d) and :249,259
 * (Conv2dZeros) — 3x3 convs arrive here as im2col / col2im GEMMs, 1x1 convs directly — and their dgrads.
 * K % 64 == 0, N % 16 == 0, lda/ldb/ldo % 8 == 0. */
int nfk_gemm_nt_bf16(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, int epi,
                     void* out, long long ldo, const float* bias, const void* aux, long long ldaux, float* colsum,
                     void* stream);

/* out[Mo,No] (fp32, caller-zeroed) += sum_k A[k,Mo] * B[k,No]; A/B bf16 row-major [Kpix, ld]. Weight gradients
 * of the coupling convs (autograd of nn.Conv2d in models/layers.py:209,249), split-K over pixels with
 * red.global.add. No % 64 == 0. */
int nfk_gemm_tn_bf16(const void* A, long long lda, const void* B, long long ldb, int Mo, int No, int Kpix,
                     float* out, long long ldo, int sm_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NFK_H_ */
