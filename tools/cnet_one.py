"""Run the fused conv#1+conv#2 kernel at the level-0 shape of the bench (M = B*256) a few times; used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M, K1p, hid = B * 256, 64, 512
col = (torch.randn(M, K1p, device="cuda") * 0.5).bfloat16()
B1 = (torch.randn(hid, K1p, device="cuda") * 0.1).bfloat16()
B2 = (torch.randn(hid, hid, device="cuda") * 0.05).bfloat16()
b1 = torch.zeros(hid, device="cuda"); b2 = torch.zeros(hid, device="cuda")
h2 = torch.empty(M, hid, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    ops.cnet_fwd_fused(col, K1p, B1, B2, b1, b2, h2, M, hid)
torch.cuda.synchronize()
print("ok")
