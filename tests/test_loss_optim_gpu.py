"""GPU: the tail of the KD step as single launches (csrc/loss_optim.cu) against plain torch on the same inputs.
Reference: pl_module.py:257-320 (NFModel.loss), :348-363 (Adam / Adamax), train.py:46 (gradient_clip_val=30)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


@pytest.mark.parametrize("B,shapes,nz,use_prior,use_w,use_perc", [
    (37, [(12, 16, 16), (24, 8, 8), (48, 4, 4), (48, 4, 4)], 768, False, False, False),   # G-CIFAR taps
    (300, [(63,), (63,)], 63, True, True, True),                                          # tabular / RICH style
    (5, [], 96, True, False, False),                                                      # nll only (kd weight 0)
])
def test_fused_loss_kernel_matches_torch(B, shapes, nz, use_prior, use_w, use_perc):
    """kd[b] = mean over levels of per-level MSE, nll[b] = -(logdet + log N(z; mean, exp(logs))) * scale,
    res = (w_nll nll + w_kd kd + w_perc perc) * sample_w, and the four batch means — forward values and the gradients
    of mean(res) + a side term on mean(nll) w.r.t. every student tap, z_last, logdet and perc."""
    from nf_distillation_b200 import functional as Fn
    g = torch.Generator(device=dev).manual_seed(B)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    s_l = [rn(B, *sh).requires_grad_(True) for sh in shapes]
    t_l = [rn(B, *sh) for sh in shapes]
    z = rn(B, nz).requires_grad_(True)
    ld = (rn(B) * 10).requires_grad_(True)
    mean, logs = (rn(nz) * 0.3, rn(nz) * 0.2) if use_prior else (None, None)
    sw = torch.rand(B, device=dev, generator=g) + 0.5 if use_w else None
    perc = torch.rand(B, device=dev, generator=g).requires_grad_(True) if use_perc else None
    w = (0.85, 0.075, 0.075 if use_perc else 0.0)
    scale = 1.0 / (math.log(2.0) * nz)
    spec = {"teacher": t_l, "prior": (mean, logs) if use_prior else None, "nll_scale": scale, "w": w, "sample_w": sw}
    means, nll, kd = Fn.KdNllLossFn.apply(spec, ld, z, perc, *s_l)
    (means[3] + 0.3 * means[0]).backward()
    got = [t.grad.clone() for t in (*s_l, z, ld)] + ([perc.grad.clone()] if use_perc else [])
    for t in (*s_l, z, ld, *([perc] if use_perc else [])):
        t.grad = None
    # torch reference
    mu = mean if use_prior else torch.zeros(nz, device=dev)
    lg = logs if use_prior else torch.zeros(nz, device=dev)
    lp = (-0.5 * (2 * lg + (z - mu) ** 2 * torch.exp(-2 * lg) + math.log(2 * math.pi))).sum(1)
    nll_r = -(ld + lp) * scale
    kd_r = sum(((a - b) ** 2).flatten(1).mean(1) for a, b in zip(s_l, t_l)) / max(len(s_l), 1) if s_l \
        else torch.zeros(B, device=dev)
    pc = perc if use_perc else torch.zeros(B, device=dev)
    res = w[0] * nll_r + w[1] * kd_r + w[2] * pc
    if use_w:
        res = res * sw
    (res.mean() + 0.3 * nll_r.mean()).backward()
    ref = [t.grad for t in (*s_l, z, ld)] + ([perc.grad] if use_perc else [])
    assert rel(nll, nll_r) < 1e-5 and (not s_l or rel(kd, kd_r) < 1e-5)
    for i, r in enumerate((nll_r.mean(), kd_r.mean() if s_l else torch.zeros((), device=dev), pc.mean(), res.mean())):
        assert abs(means[i].item() - r.item()) <= 2e-6 * abs(r.item()) + 1e-7, i
    for a, b in zip(got, ref):
        assert rel(a, b) < 1e-5
    # determinism of the batch means (last-CTA reduction in CTA order)
    m2, _, _ = Fn.KdNllLossFn.apply(spec, ld.detach(), z.detach(), None if perc is None else perc.detach(),
                                    *[t.detach() for t in s_l])
    assert torch.equal(m2, means.detach())


@pytest.mark.parametrize("kind", ["adam", "adamax"])
def test_flat_clip_and_optimiser_match_torch(kind):
    """FlatAdam (train.py) = clip_grad_norm_(30) + torch.optim.Adam / Adamax, five steps on odd-sized tensors with
    gradients large enough to be clipped in some steps and not in others."""
    from nf_distillation_b200.train import FlatAdam
    g = torch.Generator(device=dev).manual_seed(7)
    shapes = [(512, 54, 3), (17,), (1, 48, 1, 1), (3, 3), (1000,)]
    ps = [torch.nn.Parameter(torch.randn(*s, device=dev, generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt_cls = torch.optim.Adam if kind == "adam" else torch.optim.Adamax
    lr = 5e-4 if kind == "adam" else 1e-4
    ropt = opt_cls(ref, lr=lr, weight_decay=0.0)
    fopt = FlatAdam(ps, lr=lr, weight_decay=0.0, kind=kind)
    assert all(p.data_ptr() >= fopt.p.data_ptr() for p in ps), "parameters must live in the flat buffer"
    for step in range(5):
        scale = (50.0, 0.01, 3.0, 200.0, 0.5)[step]
        grads = [torch.randn(*s, device=dev, generator=g) * scale for s in shapes]
        for p, r, gr in zip(ps, ref, grads):
            p.grad, r.grad = gr.clone(), gr.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 30.0)
        ropt.step()
        fopt.gather_grads()
        fopt.step()
        assert abs(fopt.grad_norm.item() - total.item()) < 1e-5 * total.item()
        for p, r in zip(ps, ref):
            assert rel(p, r) < 2e-6, (kind, step)
    assert fopt.step_count.item() == 5
