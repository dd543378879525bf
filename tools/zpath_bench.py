"""Stand-alone timing (CUDA events, rotating buffers > L2) of the HBM-bound z-path kernels at Glow-CIFAR shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = "cuda"
HBM = 6547.5e9


def timeit(fn, nbuf, reps=20):
    """reps launches captured in one CUDA graph (so Python launch overhead is not what is measured)."""
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (C, H) in ((12, 16), (24, 8), (48, 4)):
    W = H
    M = B * H * W
    K1p, K3p = ops.round_up(9 * C // 2, 64), ops.round_up(9 * C, 64)
    nbuf = max(2, int(400e6 / (M * (C * 8 + K3p * 4))) + 1)
    xs = [torch.randn(B, C, H, W, device=dev) for _ in range(nbuf)]
    ys = [torch.empty(B, C, H, W, device=dev) for _ in range(nbuf)]
    cols = [torch.empty(M, K1p, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    Ps = [torch.randn(M, K3p, device=dev) * 0.1 for _ in range(nbuf)]
    Wf = torch.randn(C, C, device=dev) * 0.3
    bf = torch.randn(C, device=dev)
    sl = torch.zeros(1, device=dev)
    b3 = torch.zeros(C, device=dev)
    ld0, ld1 = torch.zeros(B, device=dev), torch.zeros(B, device=dev)
    t = timeit(lambda i: ops.affine1x1_fwd(xs[i], Wf, bf, sl, ys[i], cols[i], K1p, ld0, ld1, B, C, H, W), nbuf)
    by = M * (C * 8 + K1p * 2)
    print(f"C={C:2d} HxW={H}x{W} B={B}: affine1x1_fwd {t:7.1f} us  {by/1e6:7.1f} MB  {by/t*1e6/HBM*100:5.1f}% of HBM peak")
    t = timeit(lambda i: ops.coupling_fwd(Ps[i], K3p, b3, ys[i], None, ld1, B, C, H, W, False), nbuf)
    by = M * (K3p * 4 + C * 4)
    print(f"                         coupling_fwd  {t:7.1f} us  {by/1e6:7.1f} MB  {by/t*1e6/HBM*100:5.1f}% of HBM peak")
    # backward kernels of the student
    gouts = [torch.randn(B, C, H, W, device=dev) for _ in range(nbuf)]
    hsv = [torch.randn(M, C, device=dev) for _ in range(nbuf)]
    dys = [torch.empty(B, C, H, W, device=dev) for _ in range(nbuf)]
    dhcols = [torch.empty(M, K3p, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    gld = torch.randn(B, device=dev)
    db3 = torch.zeros(C, device=dev)
    t = timeit(lambda i: ops.coupling_bwd(gouts[i], gld, ys[i], hsv[i], dys[i], dhcols[i], K3p, db3, B, C, H, W), nbuf)
    by = M * (C * 4 + C * 2 + C * 4 + C * 4 + K3p * 2)
    print(f"                         coupling_bwd  {t:7.1f} us  {by/1e6:7.1f} MB  {by/t*1e6/HBM*100:5.1f}% of HBM peak")
    dcols = [torch.randn(M, K1p, device=dev) * 0.1 for _ in range(nbuf)]
    dxs = [torch.empty(B, C, H, W, device=dev) for _ in range(nbuf)]
    dWf, dbf = torch.zeros(C, C, device=dev), torch.zeros(C, device=dev)
    t = timeit(lambda i: ops.affine1x1_bwd(dys[i], dcols[i], K1p, xs[i], Wf, dxs[i], dWf, dbf, B, C, H, W), nbuf)
    by = M * (C * 12 + K1p * 4)
    print(f"                         affine1x1_bwd {t:7.1f} us  {by/1e6:7.1f} MB  {by/t*1e6/HBM*100:5.1f}% of HBM peak")
