"""Run affine1x1_fwd and the fused Conv2dZeros + coupling kernel at levels 1 and 2 of the CIFAR model (B images) a few
times; used under ncu (python tools/levels_one.py [B])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
hid = 512
for (C, H) in ((24, 8), (48, 4)):
    W = H
    M = B * H * W
    K1p, K3p = ops.round_up(9 * C // 2, 64), ops.round_up(9 * C, 64)
    x, y = torch.randn(B, C, H, W, device=dev), torch.empty(B, C, H, W, device=dev)
    col = torch.empty(M, K1p, device=dev, dtype=torch.bfloat16)
    h2 = (torch.randn(M, hid, device=dev).clamp_min(0) * 0.5).bfloat16()
    Wf, bfv, sl = torch.randn(C, C, device=dev) * 0.3, torch.randn(C, device=dev) * 0.1, torch.zeros(1, device=dev)
    ld, ld1 = torch.zeros(B, device=dev), torch.zeros(B, device=dev)
    B3 = (torch.randn(K3p, hid, device=dev) * 0.02).bfloat16()
    b3 = torch.zeros(C, device=dev)
    for _ in range(3):
        ops.affine1x1_fwd(x, Wf, bfv, sl, y, col, K1p, ld, ld1, B, C, H, W)
        ops.pconv_coupling_fwd(h2, B3, K3p, b3, y, None, ld, B, C, H, W, hid, False)
torch.cuda.synchronize()
print("ok")
