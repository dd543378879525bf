"""Building blocks of the bf16x3 precision mode (nf_distillation_b200/precise.py) against fp64 torch: the split GEMM in
its 3-term and 6-term forms, the split weight-gradient GEMM, and the ReLU / mask split kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nf_distillation_b200 import precise as P
dev = "cuda"
torch.manual_seed(0)
rel = lambda a, b: ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def rec(t, K, pat):   # split A operand -> h + m (+ l)
    h, m = t[:, :K].double(), t[:, P.M_OFF[pat] * K:(P.M_OFF[pat] + 1) * K].double()
    return h + m + (t[:, 5 * K:6 * K].double() if pat == P.A6 else 0)


for (M, N, K) in ((10240, 512, 128), (10240, 512, 512), (10240, 64, 512), (1280, 512, 256)):
    A = torch.randn(M, K, device=dev) * torch.rand(M, 1, device=dev) * (torch.rand(M, K, device=dev) > 0.5)
    Wt = torch.randn(N, K, device=dev) * 0.05
    ref = A.double() @ Wt.double().T
    for pa, pb, name in ((P.A3, P.B3, "3-term"), (P.A6, P.B6, "6-term")):
        As, Bs = P.split_rows(A, K, pa), P.split_rows(Wt, K, pb)
        print(f"M={M} N={N} K={K} {name}: split A {rel(rec(As, K, pa), A):.1e}  gemm vs fp64 {rel(P.gemm_split(As, Bs, M, N), ref):.1e}"
              f"  (fp32 torch: {rel(A @ Wt.T, ref):.1e}, plain bf16: {rel(A.bfloat16().float() @ Wt.bfloat16().float().T, ref):.1e})")
    G = torch.randn(M, N, device=dev)
    Gs, As6 = P.split_rows(G, N, P.A3), P.split_rows(A, K, P.A6)
    print(f"    wgrad (A3 x A6) vs fp64 {rel(P.wgrad_split(Gs, N, P.A3, As6, K, P.A6, M), G.double().T @ A.double()):.1e}")
    pre = torch.randn(M, N, device=dev)
    h6 = P.act_split(pre, 0, pattern=P.A6)
    cs = torch.zeros(N, device=dev)
    d3 = P.act_split(G, 1, gate=h6, colsum=cs)
    refd = G * (pre > 0)
    print(f"    relu split6 {rel(rec(h6, N, P.A6), torch.relu(pre)):.1e}  masked split3 {rel(rec(d3, N, P.A3), refd):.1e}  colsum {rel(cs, refd.sum(0)):.1e}")
