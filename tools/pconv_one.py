"""Run the fused Conv2dZeros + coupling kernel at the level-0 shape of the bench (B images of 16x16, C = 12); used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
C, H, W, hid = 12, 16, 16, 512
M, K3p = B * H * W, 128
h2 = (torch.randn(M, hid, device="cuda").clamp_min(0) * 0.5).bfloat16()
B3 = (torch.randn(K3p, hid, device="cuda") * 0.02).bfloat16()
b3 = torch.zeros(C, device="cuda"); y = torch.randn(B, C, H, W, device="cuda"); ld = torch.zeros(B, device="cuda")
for _ in range(5):
    ops.pconv_coupling_fwd(h2, B3, K3p, b3, y, None, ld, B, C, H, W, hid, False)
torch.cuda.synchronize()
print("ok")
