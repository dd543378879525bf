"""Timing of the Split2d kernels (csrc/split_prior_kd.cu) at the CIFAR KD shapes; NFK_SPLIT_GENERIC=1 selects the
generic (any C) kernels for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for C, H in ((12, 16), (24, 8)):
    W = H
    g = torch.Generator(device=dev).manual_seed(C)
    x = torch.randn(B, C, H, W, device=dev, generator=g)
    w = torch.randn(C, C // 2, 3, 3, device=dev, generator=g) * 0.05
    bias = torch.randn(C, device=dev, generator=g) * 0.1
    logs = torch.randn(C, device=dev, generator=g) * 0.05
    z1 = torch.empty(B, C // 2, H, W, device=dev); ld = torch.zeros(B, device=dev)
    gz1 = torch.randn(B, C // 2, H, W, device=dev, generator=g); gld = torch.randn(B, device=dev, generator=g)
    dx = torch.empty_like(x); dw = torch.zeros_like(w); db = torch.zeros(C, device=dev); dl = torch.zeros(C, device=dev)
    def f(): ops.split2d_fwd(x, w, bias, logs, z1, ld, B, C, H, W)
    def b(): ops.split2d_bwd(x, w, bias, logs, gz1, gld, dx, dw, db, dl, B, C, H, W)
    for name, fn in (("fwd", f), ("bwd", b)):
        fn(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(5): fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        print(f"split2d {name} B={B} C={C} {H}x{W}: {e0.elapsed_time(e1) * 200:.1f} us")
