"""GPU: the BASELINE.json configurations as parity / property cases (at sizes the CPU oracle finishes in seconds, and at
full size through size-independent properties: step-wise invertibility, log-det antisymmetry, finite sampling)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dev = "cuda"


def rel(a, b):
    b = b.to(a.device)
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def make(cfg, seed, std=0.05):
    from nf_distillation_b200.models import create_glow_model
    from nf_distillation_b200.train import randomise_zero_params
    torch.manual_seed(seed)
    m = create_glow_model(cfg)
    randomise_zero_params(m, seed + 1, std)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m.to(dev).eval(), sd


def images(B, H, g):
    return torch.floor(torch.rand(B, 3, H, H, generator=g) * 256) / 256 - 0.5


def test_config3_glow_l3_k32_h512_forward_inverse_logdet(monkeypatch):
    """BASELINE configs[2]: Glow L=3 K=32 hidden 512 on 32x32x3 — forward (vs oracle, B=2), then at B=64 every
    FlowStep's inverse + log-det antisymmetry, and temperature-0.7 sampling."""
    from nf_distillation_b200.models import utils as U
    from nf_distillation_b200.models.flows import FlowStep
    from nf_distillation_b200.train import glow_cfg
    from oracle import glow_oracle as O
    cfg = glow_cfg((32, 32, 3), 32, 3, 512)
    # std 0.01: with untrained N(0, 0.05^2) couplings the 96-step inverse from a Split2d-mean latent overflows to NaN
    # in the reference arithmetic itself (oracle and CUDA path agree on that, but it checks nothing)
    m, sd = make(cfg, 3, std=0.01)
    g = torch.Generator().manual_seed(0)
    x, noise = images(2, 32, g), torch.rand(2, 3, 32, 32, generator=g) / 256
    o_outs, o_bpd = O.glow_forward(sd, cfg, x, noise)
    monkeypatch.setattr(U, "dequant_noise", lambda t, n: noise.to(dev))
    with torch.no_grad():
        outs, bpd, _ = m(x.to(dev), None)
    assert len(outs) == 101 and rel(bpd, o_bpd) < 1e-4          # per-sample log-likelihood, north_star fp32 bound
    for i in (0, 34, 68, 100):                                    # the teacher's KD taps
        assert rel(outs[i], o_outs[i]) < 2e-2
    monkeypatch.undo()
    B = 64
    xb = images(B, 32, g).to(dev)
    with torch.no_grad():
        z, ld = xb, torch.zeros(B, device=dev)
        checked = 0
        for layer in m.flow.layers:
            zin = z
            z, ld = layer(z, logdet=ld, reverse=False)
            if isinstance(layer, FlowStep) and checked % 8 == 0:
                back, ld0 = layer(z, logdet=ld, reverse=True)
                z_prev_ld = ld0
                assert rel(back, zin) < 1e-3
            checked += isinstance(layer, FlowStep)
        assert torch.isfinite(ld).all()
        # full inverse pass (all 101 layers, Split2d returning its mean at temperature 0) against the oracle
        rev = m(z=outs[-1], temperature=0.0, reverse=True)
    o_rev = O.glow_reverse(sd, cfg, o_outs[-1], 0.0)
    assert len(rev) == 101 and rev[-1].shape == (2, 3, 32, 32)
    assert rel(rev[-1], o_rev[-1]) < 5e-2


def test_config5_glow_l4_64x64_shapes(monkeypatch):
    """BASELINE configs[4] shape family: Glow L=4 on 64x64x3 (levels C = 12, 24, 48, 96) — forward vs oracle, inverse."""
    from nf_distillation_b200.models import utils as U
    from nf_distillation_b200.train import glow_cfg
    from oracle import glow_oracle as O
    cfg = glow_cfg((64, 64, 3), 2, 4, 256)
    m, sd = make(cfg, 5)
    g = torch.Generator().manual_seed(1)
    x, noise = images(3, 64, g), torch.rand(3, 3, 64, 64, generator=g) / 256
    o_outs, o_bpd = O.glow_forward(sd, cfg, x, noise)
    monkeypatch.setattr(U, "dequant_noise", lambda t, n: noise.to(dev))
    with torch.no_grad():
        outs, bpd, _ = m(x.to(dev), None)
    assert [tuple(o.shape[1:]) for o in outs[-3:]] == [(96, 4, 4)] * 3
    assert rel(bpd, o_bpd) < 1e-4
    for a, b in zip(outs, o_outs):
        assert rel(a, b) < 2e-2
    with torch.no_grad():
        rev = m(z=outs[-1], temperature=0.0, reverse=True)
    o_rev = O.glow_reverse(sd, cfg, o_outs[-1], 0.0)
    assert rel(rev[-1], o_rev[-1]) < 3e-2


def test_config4_kd_step_l3_gradients_flow_everywhere():
    """BASELINE configs[3]: teacher K=32 -> student K=8 KD step (B=16): every student parameter gets a finite gradient,
    the teacher none, and two identical steps give identical losses (determinism of the forward path)."""
    from nf_distillation_b200.pl_module import NFModel
    from nf_distillation_b200.train import glow_cfg, kd_config, randomise_zero_params
    torch.manual_seed(0)
    m = NFModel(kd_config(glow_cfg((32, 32, 3), 8, 3, 512), glow_cfg((32, 32, 3), 32, 3, 512)))
    randomise_zero_params(m.student, 1)
    randomise_zero_params(m.teacher, 2)
    m.to(dev)
    assert (m.student_kd_indices, m.teacher_kd_indices) == ([0, 10, 20, 28], [0, 34, 68, 100])
    g = torch.Generator().manual_seed(2)
    x = images(16, 32, g).to(dev)
    torch.manual_seed(5)
    out = m.training_step([x.clone(), None], 0)
    out["loss"].backward()
    assert torch.isfinite(out["loss"]) and out["kd"].item() > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.student.parameters())
    assert all(p.grad is None for p in m.teacher.parameters())
    torch.manual_seed(5)
    out2 = m.training_step([x.clone(), None], 0)
    assert abs(out2["nll"].item() - out["nll"].item()) < 1e-6


def test_config5_celeba_64x64_training_gradients_vs_oracle(monkeypatch):
    """BASELINE configs[4] shape family in TRAINING: Glow L=4 on 64x64x3 — the level-0 maps are 32x32, so the coupling
    backward walks row bands with a halo instead of whole images. Student gradients of nll (bpd) + a latent term
    against the CPU oracle's autograd; bf16 GEMM operands: median 1e-2 of max|grad|, worst 0.25 (one element of a 4x4-level weight),
    every tensor's cosine >= 0.995 (the 4x4 top level sums only B*16 pixels, so its bf16 rounding noise is the largest)."""
    from nf_distillation_b200.models import utils as U
    from nf_distillation_b200.train import glow_cfg
    from oracle import glow_oracle as O
    cfg = glow_cfg((64, 64, 3), 1, 4, 128)
    m, sd = make(cfg, 11)
    m.train()
    g = torch.Generator().manual_seed(3)
    B = 8
    x, noise = images(B, 64, g), torch.rand(B, 3, 64, 64, generator=g) / 256
    wz = torch.randn(B, 96, 4, 4, generator=g)
    names = dict(m.named_parameters())
    osd = {k: v.clone().requires_grad_(k in names) for k, v in sd.items()}
    o_outs, o_bpd = O.glow_forward(osd, cfg, x, noise)
    (o_bpd.sum() + (o_outs[-1] * wz).sum() * 1e-2).backward()
    monkeypatch.setattr(U, "dequant_noise", lambda t, n: noise.to(dev))
    outs, bpd, _ = m(x.to(dev), None)
    assert rel(bpd, o_bpd.detach()) < 1e-4
    (bpd.sum() + (outs[-1] * wz.to(dev)).sum() * 1e-2).backward()
    errs = []
    for n_, p in m.named_parameters():
        if osd[n_].grad is None:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n_
        errs.append((rel(p.grad, osd[n_].grad), n_))
        a, b = p.grad.detach().cpu().flatten().double(), osd[n_].grad.flatten().double()
        assert (a @ b / (a.norm() * b.norm() + 1e-30)).item() > 0.995, n_
    errs.sort()
    assert errs[len(errs) // 2][0] < 1e-2 and errs[-1][0] < 0.25, (errs[len(errs) // 2], errs[-1])


def test_eager_training_steps_do_not_accumulate_device_memory():
    """The batched parameter prep hands a token through the autograd graph; nothing of a finished step (operands,
    deposits, activations) may stay reachable once its backward has run, even with the cyclic GC off."""
    import gc
    from nf_distillation_b200.pl_module import NFModel
    from nf_distillation_b200.train import glow_cfg, kd_config, randomise_zero_params
    torch.manual_seed(0)
    m = NFModel(kd_config(glow_cfg((32, 32, 3), 2, 3, 128), glow_cfg((32, 32, 3), 3, 3, 128)))
    randomise_zero_params(m.student, 1)
    randomise_zero_params(m.teacher, 2)
    m.to(dev)
    g = torch.Generator().manual_seed(2)
    x = images(8, 32, g).to(dev)
    gc.collect()
    gc.disable()
    try:
        used = []
        for it in range(6):
            out = m.training_step([x.clone(), None], 0)
            out["loss"].backward()
            for p in m.student.parameters():
                p.grad = None
            del out
            torch.cuda.synchronize()
            used.append(torch.cuda.memory_allocated())
        assert used[5] <= used[2] + (1 << 20), used
    finally:
        gc.enable()


def test_generate_sampling_path_matches_oracle_draw_for_draw():
    """NFModel.generate (pl_module.py:322-346): z ~ prior at temperature 1, then the inverse pass with every Split2d
    drawing its half from the learned conditional Gaussian. With the same torch generator state the kernels consume the
    same draws as the reference code order (prior sample first, then one draw per Split2d in decode order), so the
    samples can be checked against the oracle's inverse fed with those draws. 2-D (L=3, two Split2d) and 1-D."""
    from nf_distillation_b200.pl_module import NFModel
    from nf_distillation_b200.train import glow_cfg, kd_config, randomise_zero_params
    from oracle import glow_oracle as O
    # ---- 2-D
    cfg = glow_cfg((32, 32, 3), 2, 3, 64)
    torch.manual_seed(0)
    m = NFModel(kd_config(cfg, cfg, kd=0.0))
    randomise_zero_params(m.student, 1, 0.01)
    sd = {k: v.clone() for k, v in m.student.state_dict().items()}
    m.to(dev)
    B = 6
    y = torch.nn.functional.one_hot(torch.arange(B) % 10, 10).float().to(dev)
    torch.manual_seed(123)
    xs = m.generate([torch.zeros(B, 3, 32, 32, device=dev), y])
    assert xs.shape == (32, 3, 32, 32) and torch.isfinite(xs).all()   # no y_condition: the prior's default batch of 32
    torch.manual_seed(123)
    z_top = torch.normal(torch.zeros(32, 48, 4, 4, device=dev), torch.ones(32, 48, 4, 4, device=dev))
    eps = [torch.randn(32, 12, 8, 8, device=dev), torch.randn(32, 6, 16, 16, device=dev)]   # shape of each z1
    ref = O.glow_reverse(sd, cfg, z_top.cpu(), 1.0, eps=[e.cpu() for e in eps])[-1]
    assert rel(xs, ref) < 3e-2
    # ---- 1-D (the batch only sizes the prior)
    cfg1 = glow_cfg([21], 3, 1, 32, is_1d=True, y_classes=0)
    torch.manual_seed(1)
    m1 = NFModel(kd_config(cfg1, cfg1, data="hepmass", kd=0.0))
    randomise_zero_params(m1.student, 2, 0.05)
    sd1 = {k: v.clone() for k, v in m1.student.state_dict().items()}
    m1.to(dev)
    torch.manual_seed(7)
    x1 = m1.generate([torch.zeros(100, 21, device=dev)])
    torch.manual_seed(7)
    z1 = torch.normal(torch.zeros(100, 21, device=dev), torch.ones(100, 21, device=dev))
    ref1 = O.glow_reverse(sd1, cfg1, z1.cpu(), 1.0)[-1]
    assert x1.shape == (100, 21) and rel(x1, ref1) < 1e-4
