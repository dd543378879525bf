"""One launch of the resident MADE inverse (for ncu): python tools/maf_inv_one.py [B] [D] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nf_distillation_b200.models.maf import MADE  # noqa: E402

B, D, H = (int(a) for a in (sys.argv[1:4] + ["65536", "63", "512"][len(sys.argv) - 1:]))
torch.manual_seed(0)
made = MADE(D, H, flip=True).cuda()
u = torch.randn(B, D, device="cuda")
with torch.no_grad():
    for _ in range(2):
        made(u, logdet=torch.zeros(B, device="cuda"), reverse=True)
torch.cuda.synchronize()
