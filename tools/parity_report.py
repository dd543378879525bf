"""Error tables behind the bounds of tests/test_headline_parity_gpu.py: CUDA path vs the CPU oracle in its fp32 and
bf16-operand modes, at hidden 512 (run on the GPU box; prints only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_headline_parity_gpu as T
from oracle import glow_oracle as O
dev = "cuda"


def flowstep(C, H, B, precision="bf16"):
    st, sd = T.make_step(C, 512, 100 + C)
    st.precision = precision
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, H, H, generator=g); ld0 = torch.randn(B, generator=g)
    wz, wl = torch.randn(B, C, H, H, generator=g), torch.randn(B, generator=g)
    res = {}
    for bf16 in (False, True):
        osd = {k: (v.clone().requires_grad_(True) if k in dict(st.named_parameters()) else v) for k, v in sd.items()}
        xo = x.clone().requires_grad_(True)
        with O.bf16_operands(bf16):
            z, ld = O.flowstep(xo, osd, "", ld0, False)
            ((z * wz).sum() + (ld * wl).sum()).backward()
        res[bf16] = (z.detach(), ld.detach(), xo.grad, {k: v.grad for k, v in osd.items() if getattr(v, "grad", None) is not None})
    st = st.to(dev)
    xg = x.to(dev).requires_grad_(True)
    zt, ldt = st(xg, logdet=ld0.to(dev), reverse=False)
    ((zt * wz.to(dev)).sum() + (ldt * wl.to(dev)).sum()).backward()
    print(f"--- FlowStep C={C} H={H} B={B} (M={B*H*H}) precision={precision}")
    for bf16 in (False, True):
        z, ld, dx, gr = res[bf16]
        tag = "bf16-oracle" if bf16 else "fp32-oracle"
        print(f"  {tag}: z {T.rel(zt, z):.2e}  ld {T.rel(ldt, ld):.2e}  dx {T.rel(xg.grad, dx):.2e}")
        for n_, p in st.named_parameters():
            print(f"      {n_:28s} {T.rel(p.grad, gr[n_]):.2e}")


def kd_step():
    from nf_distillation_b200.models import utils as U
    m, s_cfg, t_cfg, s_sd, t_sd = T.kd_models(8, 32, 512)
    names = set(dict(m.student.named_parameters()))
    g = torch.Generator().manual_seed(11)
    B = 32
    x = T.images(B, 32, g)
    n1, n2 = torch.rand(B, 3, 32, 32, generator=g) / 256, torch.rand(B, 3, 32, 32, generator=g) / 256
    import time
    t0 = time.time()
    l32, g32, z32 = T.oracle_kd_grads(s_sd, s_cfg, t_sd, t_cfg, x, n1, n2, names, False)
    t1 = time.time()
    l16, g16, z16 = T.oracle_kd_grads(s_sd, s_cfg, t_sd, t_cfg, x, n1, n2, names, True)
    print(f"--- KD step K32->K8 hid512 B=32: oracle fp32 {t1-t0:.1f}s, bf16 mode {time.time()-t1:.1f}s")
    m.to(dev)
    q = [n1.to(dev), n2.to(dev)]
    orig = U.dequant_noise
    U.dequant_noise = lambda t_, n: q.pop(0)
    out = m.training_step([x.to(dev), None], 0)
    out["loss"].backward()
    U.dequant_noise = orig
    print("  losses cuda", {k: round(v.item(), 6) for k, v in out.items()}, "\n  fp32", l32, "\n  bf16", l16)
    e16 = sorted((T.rel(p.grad, g16[n_]), n_) for n_, p in m.student.named_parameters())
    e32 = sorted((T.rel(p.grad, g32[n_]), n_) for n_, p in m.student.named_parameters())
    for tag, e in (("bf16-oracle", e16), ("fp32-oracle", e32)):
        print(f"  grads vs {tag}: median {e[len(e)//2][0]:.2e}  p90 {e[int(len(e)*0.9)][0]:.2e}  worst:", [(f"{a:.2e}", b) for a, b in e[-4:]])


if __name__ == "__main__":
    what = sys.argv[1:] or ["steps", "kd"]
    if "steps" in what:
        for C, H, B in ((12, 16, 40), (24, 8, 136), (48, 4, 520)):
            flowstep(C, H, B)
    if "x3" in what:
        for C, H, B in ((12, 16, 40), (24, 8, 20)):
            flowstep(C, H, B, "bf16x3")
    if "kd" in what:
        kd_step()
