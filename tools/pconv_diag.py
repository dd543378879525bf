"""Fused Conv2dZeros+coupling kernel (csrc/pconv_coupling.cu) against the two-kernel path it replaces."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops

dev = torch.device("cuda:0")
F32, BF16 = torch.float32, torch.bfloat16


def run(B, C, H, W, hid=512, reverse=False, reps=10):
    M = B * H * W
    K3p = ops.round_up(9 * C, 64)
    g = torch.Generator(device=dev).manual_seed(B + C)
    h2 = (torch.randn(M, hid, device=dev, generator=g).clamp_min(0) * 0.5).to(BF16)
    B3 = torch.zeros(K3p, hid, device=dev, dtype=BF16)
    B3[:9 * C] = (torch.randn(9 * C, hid, device=dev, generator=g) * 0.02).to(BF16)
    bias3 = torch.randn(C, device=dev, generator=g) * 0.1
    y0 = torch.randn(B, C, H, W, device=dev, generator=g)
    ld0 = torch.randn(B, device=dev, generator=g)
    # reference path
    ya, lda, hsa = y0.clone(), ld0.clone(), torch.empty(M, C, device=dev)
    P = torch.empty(M, K3p, device=dev, dtype=F32)
    ops.gemm_nt(h2, B3, M, K3p, hid, ops.EPI_F32, P)
    ops.coupling_fwd(P, K3p, bias3, ya, hsa, lda, B, C, H, W, reverse=reverse)
    yb, ldb, hsb = y0.clone(), ld0.clone(), torch.empty(M, C, device=dev)
    assert ops.pconv_coupling_supported(C, H, W, hid)
    ops.pconv_coupling_fwd(h2, B3, K3p, bias3, yb, hsb, ldb, B, C, H, W, hid, reverse)
    torch.cuda.synchronize()
    r = lambda a, b: ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
    print(f"B={B} C={C} {H}x{W} rev={int(reverse)}: y {r(yb, ya):.2e}  hsave {r(hsb, hsa):.2e}  ld {r(ldb, lda):.2e}", flush=True)
    # timing in CUDA graphs (rotating y copies are not needed: in-place update of the same maps)
    def t_old():
        ops.gemm_nt(h2, B3, M, K3p, hid, ops.EPI_F32, P)
        ops.coupling_fwd(P, K3p, bias3, ya, hsa, lda, B, C, H, W, reverse=reverse)
    def t_new():
        ops.pconv_coupling_fwd(h2, B3, K3p, bias3, yb, hsb, ldb, B, C, H, W, hid, reverse)
    for name, fn in (("gemm+coupling", t_old), ("fused", t_new)):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps):
                fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        print(f"   {name:14s} {e0.elapsed_time(e1) * 1e3 / reps:8.1f} us", flush=True)


if __name__ == "__main__":
    run(4, 12, 16, 16)
    run(3, 24, 8, 8)
    run(1024, 12, 16, 16)
    run(1024, 12, 16, 16, reverse=True)
    run(1024, 24, 8, 8)
    run(1023, 24, 8, 8)
    run(5, 48, 4, 4)
    run(1024, 48, 4, 4)
    run(1021, 48, 4, 4, reverse=True)
    run(9, 12, 8, 8)
    run(3, 12, 32, 32)
    run(5, 12, 8, 32)
    run(256, 12, 32, 32)
    run(255, 12, 32, 32, reverse=True)
