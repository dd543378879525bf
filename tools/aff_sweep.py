"""Sweep of the affine1x1_fwd launch heuristics (output groups per pixel, images per CTA) through NFK_AFF_OG / NFK_AFF_IPC
(one process per point: the overrides are read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
import torch
from nf_distillation_b200 import ops
dev = "cuda"
def tg(fn, reps, nb):
    for i in range(nb): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn(i %% nb)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
out = []
for B in (2048, 64):
    for (C, H) in ((12, 16), (24, 8), (48, 4)):
        M = B * H * H; K1p = ops.round_up(9 * C // 2, 64); nb = 6
        xs = [torch.randn(B, C, H, H, device=dev) for _ in range(nb)]
        ys = [torch.empty(B, C, H, H, device=dev) for _ in range(nb)]
        cols = [torch.empty(M, K1p, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
        Wf, bfv, sl = torch.randn(C, C, device=dev) * 0.3, torch.randn(C, device=dev) * 0.1, torch.zeros(1, device=dev)
        ld, ld1 = torch.zeros(B, device=dev), torch.zeros(B, device=dev)
        try:
            t = tg(lambda i: ops.affine1x1_fwd(xs[i], Wf, bfv, sl, ys[i], cols[i], K1p, ld, ld1, B, C, H, H), 16, nb)
        except Exception as e:
            t = float("nan")
        out.append("%%6.1f" %% t)
print(" ".join(out))
''' % ROOT
print("og ipc | B=2048: C12 C24 C48 | B=64: C12 C24 C48   (us)")
for og in (0, 1, 2, 4, 8):
    for ipc in (0, 1, 2, 4, 8, 16):
        env = dict(os.environ)
        if og: env["NFK_AFF_OG"] = str(og)
        if ipc: env["NFK_AFF_IPC"] = str(ipc)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        print(f"{og:2d} {ipc:3d} |", r.stdout.strip() or r.stderr.strip()[-200:], flush=True)
