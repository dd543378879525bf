"""Does programmatic dependent launch survive CUDA-graph capture? A chain of small dependent kernels, timed inside one
graph; run with NFK_PDL=1 and NFK_PDL=0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
dev = "cuda"
B, C, H, W = 64, 12, 16, 16
M, K1p = B * H * W, 64
x = torch.randn(B, C, H, W, device=dev); y = torch.empty_like(x)
col = torch.empty(M, K1p, device=dev, dtype=torch.bfloat16)
Wf = torch.eye(C, device=dev); bf = torch.zeros(C, device=dev); sl = torch.zeros(1, device=dev)
ld0, ld1 = torch.zeros(B, device=dev), torch.zeros(B, device=dev)
hid = 512
B1 = (torch.randn(hid, K1p, device=dev) * 0.1).bfloat16(); B2 = (torch.randn(hid, hid, device=dev) * 0.05).bfloat16()
b1 = torch.zeros(hid, device=dev); h2 = torch.empty(M, hid, device=dev, dtype=torch.bfloat16)
B3 = (torch.randn(128, hid, device=dev) * 0.02).bfloat16(); b3 = torch.zeros(C, device=dev)
def chain(n):
    for _ in range(n):
        ops.affine1x1_fwd(x, Wf, bf, sl, y, col, K1p, ld0, ld1, B, C, H, W)
        ops.cnet_fwd_fused(col, K1p, B1, B2, b1, b1, h2, M, hid)
        ops.pconv_coupling_fwd(h2, B3, 128, b3, y, None, ld1, B, C, H, W, hid, False)
chain(3); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain(100)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"NFK_PDL={os.environ.get('NFK_PDL', '1')}: {e0.elapsed_time(e1) * 1e3 / 300:.2f} us per kernel (graph of 300 dependent launches, B=64 level-0 FlowStep)")
