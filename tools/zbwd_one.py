"""Run coupling_bwd and affine1x1_bwd at the level-0 shape of the bench (B images of 16x16, C = 12) a few times; used
under ncu (python tools/zbwd_one.py [B])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
C, H, W = 12, 16, 16
M = B * H * W
K1p, K3p = ops.round_up(9 * C // 2, 64), ops.round_up(9 * C, 64)
x, y = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
gout, dy, dx = torch.randn(B, C, H, W, device=dev), torch.empty(B, C, H, W, device=dev), torch.empty(B, C, H, W, device=dev)
hsv = torch.randn(M, C, device=dev)
dhcol = torch.empty(M, K3p, device=dev, dtype=torch.bfloat16)
dcol = torch.randn(M, K1p, device=dev) * 0.1
gld, db3 = torch.randn(B, device=dev), torch.zeros(C, device=dev)
Wf, dWf, dbf = torch.randn(C, C, device=dev) * 0.3, torch.zeros(C, C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    ops.coupling_bwd(gout, gld, y, hsv, dy, dhcol, K3p, db3, B, C, H, W)
    ops.affine1x1_bwd(dy, dcol, K1p, x, Wf, dx, dWf, dbf, B, C, H, W)
torch.cuda.synchronize()
print("ok")
