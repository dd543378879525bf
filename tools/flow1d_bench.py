"""Micro-benchmark of the fused 1-D FlowStep kernels (csrc/flow1d.cu): each launch timed inside a CUDA graph of
REPS launches; --ncu runs each once eagerly between cudaProfilerStart/Stop."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--D", type=int, default=63)
ap.add_argument("--ncu", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
B, D = args.B, args.D
REPS = 10


def setup(hid):
    tf, tb, tg, n_act, offs = ops.flow1d_sizes(D, 0, hid)
    g = torch.Generator(device=dev).manual_seed(0)
    PF = torch.randn(tf, device=dev, generator=g) * 0.1
    PB = torch.randn(tb, device=dev, generator=g) * 0.1
    return PF, PB, tg, n_act


def timeit(name, fn, bytes_):
    fn(); torch.cuda.synchronize()
    if args.ncu:
        torch.cuda.profiler.start(); fn(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
        return
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(REPS):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / REPS
    print(f"{name:34s} {us:8.1f} us   {bytes_ / us / 1e3:7.1f} GB/s algorithmic")


x = torch.randn(B, D, device=dev)
ld = torch.zeros(B, device=dev)
sl = torch.zeros(1, device=dev)
for hid in (32, 16):
    PF, PB, tg, n_act = setup(hid)
    y = torch.empty_like(x); ld_out = torch.empty_like(ld)
    acts = torch.empty(B, n_act, device=dev)
    for rev in (0, 1):
        timeit(f"fwd hid={hid} rev={rev} infer", lambda: ops.flow1d_fwd(x, None, PF, sl, y, ld, ld_out, None, B, D, 0, hid, rev),
               2 * B * D * 4)
        timeit(f"fwd hid={hid} rev={rev} save", lambda: ops.flow1d_fwd(x, None, PF, sl, y, ld, ld_out, acts, B, D, 0, hid, rev),
               (2 * D + n_act) * B * 4)
        G = torch.zeros(tg, device=dev); dx = torch.empty_like(x); gy = torch.randn_like(x); gl = torch.randn_like(ld)
        timeit(f"bwd hid={hid} rev={rev}", lambda: ops.flow1d_bwd(x, None, acts, PB, y, gy, gl, dx, G, B, D, 0, hid, rev),
               (4 * D + n_act) * B * 4)
