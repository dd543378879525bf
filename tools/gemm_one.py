"""Run the conv#2-shaped GEMM (M x 512 x 512, bias+ReLU -> bf16) a few times; used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
epi = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = int(sys.argv[3]) if len(sys.argv) > 3 else 512
N = 512
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.zeros(N, device="cuda")
mask = ops.relu_mask_like(M, N, "cuda"); mask.fill_(-1)  # all active
cs = torch.zeros(N, device="cuda")
for _ in range(5):
    if epi == 1:
        ops.gemm_nt(A, B, M, N, K, ops.EPI_BIAS_RELU_BF16, out, bias=bias)
    else:
        ops.gemm_nt(A, B, M, N, K, ops.EPI_MASK_BF16, out, aux=mask, colsum=cs)
torch.cuda.synchronize()
if os.environ.get("NFK_PROF"):
    from nf_distillation_b200._lib import LIB
    prof = torch.zeros(148, 8, device="cuda", dtype=torch.int64)
    LIB.nfk_gemm_set_prof(prof.data_ptr())
    if epi == 1:
        ops.gemm_nt(A, B, M, N, K, ops.EPI_BIAS_RELU_BF16, out, bias=bias)
    else:
        ops.gemm_nt(A, B, M, N, K, ops.EPI_MASK_BF16, out, aux=mask, colsum=cs)
    torch.cuda.synchronize()
    LIB.nfk_gemm_set_prof(None)
    p = prof.cpu().double()
    lead = p[p[:, 0] > 0]
    print(f"MMA issuer (leaders, {len(lead)}): total {lead[:,0].mean():.0f} cyc, wait accumulator-free {lead[:,1].mean():.0f}, "
          f"wait operands {lead[:,2].mean():.0f}, tiles {lead[:,3].mean():.1f}")
    print(f"epilogue warp: wait accumulator {p[:,4].mean():.0f}, drain {p[:,5].mean():.0f} per CTA; per tile drain "
          f"{(p[:,5].sum() / max(lead[:,3].sum() * (len(p) / max(len(lead),1)), 1)):.0f}")
print("ok")
