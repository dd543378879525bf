"""Run the weight-gradient GEMM (out[Mo,No] += A[Kpix,Mo]^T B[Kpix,No]) at a coupling-net shape a few times; used under
ncu and for timing: tn_one.py Kpix Mo No."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
Kpix = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
Mo = int(sys.argv[2]) if len(sys.argv) > 2 else 512
No = int(sys.argv[3]) if len(sys.argv) > 3 else 512
A = (torch.randn(Kpix, Mo, device="cuda") * 0.1).bfloat16(); B = (torch.randn(Kpix, No, device="cuda") * 0.1).bfloat16()
out = torch.zeros(Mo, No, device="cuda")
for _ in range(3):
    ops.gemm_tn(A, B, Mo, No, Kpix, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm_tn(A, B, Mo, No, Kpix, out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
ref = (A[:4096].float().T @ B[:4096].float())
out.zero_(); ops.gemm_tn(A[:4096], B[:4096], Mo, No, 4096, out); torch.cuda.synchronize()
print(f"tn Kpix={Kpix} Mo={Mo} No={No}: {us:.1f} us, {2.0*Kpix*Mo*No/us/1e6:.0f} TFLOP/s, err {(out-ref).abs().max().item()/ref.abs().max().item():.2e}")
