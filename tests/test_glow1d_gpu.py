"""GPU parity of the fused 1-D (tabular) Glow kernels against vectors recorded from the unmodified reference.
Everything here is fp32 end to end; tolerances: outputs / log-dets / bpd 1e-5 relative, deterministic inverse 1e-3,
loss scalars 1e-5, student gradients (forward AND inverse-pass, i.e. the perceptual term) 1e-4 relative to max|grad|."""
import json

import pytest
import torch

from golden_util import cfg_of, load, state_dict_of, t
from test_glow2d_gpu import nf_config, rel

pytestmark = pytest.mark.gpu
dev = "cuda"


def build(name):
    from nf_distillation_b200.models import create_glow_model
    d = load(name)
    cfg = cfg_of(d)
    m = create_glow_model(cfg)
    m.load_state_dict(state_dict_of(d))
    return d, cfg, m.to(dev).eval()


@pytest.mark.parametrize("name", ["glow1d_d6_k5_h32", "glow1d_d63_k5_h32"])
def test_forward_reverse_golden(name):
    d, cfg, m = build(name)
    x = t(d["x"]).to(dev)
    with torch.no_grad():
        outs, nll, _ = m(x.clone(), None)
    n = sum(1 for k in d if k.startswith("out."))
    assert len(outs) == n
    for i, o in enumerate(outs):
        assert rel(o, t(d[f"out.{i}"])) < 1e-5
    assert rel(nll, t(d["bpd"])) < 1e-5
    with torch.no_grad():
        rev = m(z=t(d[f"out.{n - 1}"]).to(dev), temperature=0.0, reverse=True)
    assert len(rev) == int(d["rev_n"]) and rel(rev[-1], t(d["rev_last"])) < 1e-3
    B = x.shape[0]
    with torch.no_grad():
        out, ld = m.flow.layers[0](x, logdet=torch.zeros(B, device=dev), reverse=False)
        back, ldr = m.flow.layers[0](out, logdet=torch.zeros(B, device=dev), reverse=True)
    assert rel(ld, t(d["step.0.logdet_fwd"])) < 1e-5 and rel(ldr, t(d["step.0.logdet_rev"])) < 1e-5
    assert rel(back, x) < 1e-4


def test_kd_training_step_golden_with_inverse_pass_gradients(monkeypatch):
    import nf_distillation_b200.pl_module as PM
    d = load("kd1d_d63_t5_s3")
    s_cfg, t_cfg = cfg_of(d, "s_cfg"), cfg_of(d, "t_cfg")
    w = json.loads(str(d["weights"]))
    m = PM.NFModel(nf_config(s_cfg, t_cfg, w, "bsds300"))
    m.student.load_state_dict(state_dict_of(d, "s_sd."))
    m.teacher.load_state_dict(state_dict_of(d, "t_sd."))
    m = m.to(dev)
    assert m.student_kd_indices == list(d["s_idx"]) and m.teacher_kd_indices == list(d["t_idx"])
    lat = t(d["latent"]).to(dev)
    monkeypatch.setattr(PM, "gaussian_sample", lambda mean, logs, T: lat)
    out = m.training_step([t(d["x"]).to(dev)], 0)
    for k_, ref in (("nll", "nll"), ("kd", "kd"), ("perceptual", "perceptual"), ("loss", "loss")):
        assert abs(out[k_].item() - float(d[ref])) < 1e-5 * abs(float(d[ref])) + 1e-7, k_
    out["loss"].backward()
    for n_, p in m.student.named_parameters():
        assert p.grad is not None, n_
        assert rel(p.grad, t(d["grad." + n_])) < 1e-4, n_


@pytest.mark.parametrize("D,hid,K", [(6, 32, 5), (63, 16, 3), (8, 32, 2), (43, 32, 2)])
def test_full_batch_roundtrip_and_logdet_antisymmetry(D, hid, K):
    """Reference batch size (65 536, conf/training/tabular.yaml:7): x -> z -> x and logdet_fwd + logdet_rev = 0."""
    from nf_distillation_b200.models import create_glow_model
    from nf_distillation_b200.train import glow_cfg, randomise_zero_params
    torch.manual_seed(D)
    m = create_glow_model(glow_cfg([D], K, 1, hid, is_1d=True, y_classes=0))
    randomise_zero_params(m, 3)
    m = m.to(dev).eval()
    B = 65536 + 17   # ragged last tile
    x = torch.randn(B, D, device=dev)
    with torch.no_grad():
        outs, nll, _ = m(x, None)
        back = m(z=outs[-1], temperature=0.0, reverse=True)[-1]
        assert torch.isfinite(nll).all() and nll.shape == (B,)
        assert rel(back, x) < 1e-3
        z, ld = x, torch.zeros(B, device=dev)
        for layer in m.flow.layers:
            z, ld = layer(z, logdet=ld, reverse=False)
        for layer in reversed(m.flow.layers):
            z, ld = layer(z, logdet=ld, reverse=True)
        assert ld.abs().max().item() < 1e-2 and rel(z, x) < 1e-3


def test_standalone_1d_layers_and_empty_grad_cases():
    from nf_distillation_b200.models.layers import ActNorm1d, InvertibleConv1x1
    torch.manual_seed(0)
    an, iv = ActNorm1d(21).to(dev), InvertibleConv1x1(21, LU_decomposed=True, is_1d=True).to(dev)
    an.inited = True
    with torch.no_grad():
        an.bias.normal_(0, 0.3); an.logs.normal_(0, 0.3)
        x = torch.randn(100, 21, device=dev)
        y, ld = an(x, logdet=torch.zeros(100, device=dev))
        assert rel(y, (x + an.bias) * torch.exp(an.logs)) < 1e-6 and rel(ld, an.logs.sum().expand(100)) < 1e-6
        z, ld2 = iv(x, logdet=torch.zeros(100, device=dev))
        L = torch.tril(iv.lower, -1) + torch.eye(21, device=dev)
        U = torch.triu(iv.upper, 1) + torch.diag(iv.sign_s * torch.exp(iv.log_s))
        Wm = iv.p @ L @ U
        assert rel(z, x @ Wm) < 1e-5 and rel(ld2, iv.log_s.sum().expand(100)) < 1e-5
        xb, ld3 = iv(z, logdet=ld2, reverse=True)
        assert rel(xb, x) < 1e-4 and ld3.abs().max().item() < 1e-4


@pytest.mark.parametrize("D,hid,K,B", [(63, 16, 2, 1003), (63, 32, 1, 131), (6, 32, 2, 517), (43, 24, 1, 257)])
def test_ragged_batch_gradients_match_oracle_both_directions(D, hid, K, B):
    """Parameter and input gradients of forward and inverse passes on batches that do not fill the kernels' sample
    tiles (tail tile, tails that are not a multiple of four), against the CPU oracle's autograd. fp32: 1e-4 of max|grad|."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import glow_oracle as O
    from nf_distillation_b200.models import create_glow_model
    from nf_distillation_b200.train import glow_cfg, randomise_zero_params
    torch.manual_seed(D + B)
    cfg = glow_cfg([D], K, 1, hid, is_1d=True, y_classes=0)
    m = create_glow_model(cfg)
    randomise_zero_params(m, 5)
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k in dict(m.named_parameters()))
          for k, v in m.state_dict().items()}
    x_cpu = torch.randn(B, D)
    zin_cpu = torch.randn(B, D)
    wz, wn, wx = torch.randn(B, D), torch.randn(B), torch.randn(B, D)

    # oracle: loss = <wz, z_last> + <wn, nll> + <wx, x_rev>
    xo = x_cpu.clone().requires_grad_(True)
    zo = zin_cpu.clone().requires_grad_(True)
    outs, nll = O.glow_forward(sd, cfg, xo)
    rev = O.glow_reverse(sd, cfg, zo, 0.0)
    ((outs[-1] * wz).sum() + (nll * wn).sum() + (rev[-1] * wx).sum()).backward()

    m = m.to(dev).train()
    for layer in m.flow.layers:
        if hasattr(layer, "actnorm"):
            layer.actnorm.inited = True
    xg = x_cpu.to(dev).requires_grad_(True)
    zg = zin_cpu.to(dev).requires_grad_(True)
    outs_g, nll_g, _ = m(xg, None)
    rev_g = m(z=zg, temperature=0.0, reverse=True)
    assert rel(outs_g[-1], outs[-1].detach()) < 1e-5 and rel(nll_g, nll.detach()) < 1e-5
    assert rel(rev_g[-1], rev[-1].detach()) < 1e-4
    ((outs_g[-1] * wz.to(dev)).sum() + (nll_g * wn.to(dev)).sum() + (rev_g[-1] * wx.to(dev)).sum()).backward()
    assert rel(xg.grad, xo.grad) < 1e-4 and rel(zg.grad, zo.grad) < 1e-4
    for n_, p in m.named_parameters():
        if sd[n_].grad is None:
            continue
        assert p.grad is not None, n_
        assert rel(p.grad, sd[n_].grad) < 1e-4, n_


@pytest.mark.parametrize("D,Cc,hid,B", [(10, 3, 32, 300), (63, 5, 16, 131)])
def test_conditioned_1d_step_matches_oracle(D, Cc, hid, B):
    """y-conditioned 1-D FlowStep (the coupling MLP sees [z1 | y_onehot], reference flows.py:156-166 with
    condition_features > 0): forward, inverse and parameter / input gradients against the oracle step. fp32."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import glow_oracle as O
    from nf_distillation_b200.models.flows import FlowStep
    torch.manual_seed(D * 7 + Cc)
    step = FlowStep(D, hid, 1.0, "invconv", "affine", True, is_1d=True, condition_features=Cc)
    with torch.no_grad():
        for p in step.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.05)
    step.actnorm.inited = True
    sd = {"s." + k: v.clone().requires_grad_(k in dict(step.named_parameters())) for k, v in step.state_dict().items()}
    x = torch.randn(B, D)
    cond = torch.nn.functional.one_hot(torch.randint(0, Cc, (B,)), Cc).float()
    wz, wl = torch.randn(B, D), torch.randn(B)
    xo = x.clone().requires_grad_(True)
    zo, ldo = O.flowstep(xo, sd, "s.", torch.zeros(B), False, y_onehot=cond)
    xr, ldr = O.flowstep(zo.detach(), sd, "s.", torch.zeros(B), True, y_onehot=cond)
    ((zo * wz).sum() + (ldo * wl).sum()).backward()
    step = step.to(dev).train()
    xg = x.to(dev).requires_grad_(True)
    z, ld = step(xg, y_onehot=cond.to(dev), logdet=torch.zeros(B, device=dev), reverse=False)
    assert rel(z, zo.detach()) < 1e-5 and rel(ld, ldo.detach()) < 1e-5
    with torch.no_grad():
        back, ldb = step(z.detach(), y_onehot=cond.to(dev), logdet=torch.zeros(B, device=dev), reverse=True)
    assert rel(back, xr.detach()) < 1e-4 and rel(ldb, ldr.detach()) < 1e-4 and rel(back, x) < 1e-4
    ((z * wz.to(dev)).sum() + (ld * wl.to(dev)).sum()).backward()
    assert rel(xg.grad, xo.grad) < 1e-4
    for n_, p in step.named_parameters():
        ref = sd["s." + n_].grad
        if ref is None:
            continue
        assert p.grad is not None and rel(p.grad, ref) < 1e-4, n_


@pytest.mark.parametrize("D,Cc,hid,B", [(5, 3, 256, 300), (63, 0, 128, 1000)])
def test_wide_frozen_1d_step_runs_on_the_gemm_path(D, Cc, hid, B):
    """conf/teacher/rich.yaml shape (D=5, 3 condition classes, hidden 256): the coupling MLP does not fit the fused
    kernel's shared memory, so a frozen step runs its six Linear layers as tcgen05 GEMMs (bf16 operands: 2e-2 of the
    output range, 1e-2 on the per-sample log-det); forward and inverse against the oracle, and x -> z -> x."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import glow_oracle as O
    from nf_distillation_b200.models.flows import FlowStep
    torch.manual_seed(D + hid)
    step = FlowStep(D, hid, 1.0, "invconv", "affine", True, is_1d=True, condition_features=Cc)
    with torch.no_grad():
        for p in step.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.05)
    step.actnorm.inited = True
    sd = {"s." + k: v.clone() for k, v in step.state_dict().items()}
    x = torch.randn(B, D)
    cond = torch.nn.functional.one_hot(torch.randint(0, Cc, (B,)), Cc).float() if Cc else None
    zo, ldo = O.flowstep(x, sd, "s.", torch.zeros(B), False, y_onehot=cond)
    xr, ldr = O.flowstep(zo, sd, "s.", torch.zeros(B), True, y_onehot=cond)
    step = step.to(dev).eval()
    cg = None if cond is None else cond.to(dev)
    with torch.no_grad():
        z, ld = step(x.to(dev), y_onehot=cg, logdet=torch.zeros(B, device=dev), reverse=False)
        back, ldb = step(z, y_onehot=cg, logdet=ld, reverse=True)
    assert rel(z, zo) < 2e-2 and (ld.cpu() - ldo).abs().max() < 1e-2 * (ldo.abs().max() + 1)
    assert rel(back, x) < 1e-4 and ldb.abs().max().item() < 1e-4      # exact inverse of its own forward
    with pytest.raises(NotImplementedError):
        step.train()
        step(x.to(dev).requires_grad_(True), y_onehot=cg, logdet=torch.zeros(B, device=dev), reverse=False)


def test_rich_shaped_kd_step_golden(monkeypatch):
    """conf/{teacher,student,training}/rich.yaml shape, recorded from the unmodified reference (oracle/make_golden.py,
    fixture kd1d_rich_t2_s2): D = 5, y-conditioned on 3 classes (coupling MLPs see [z1 | y], the prior gets the
    LinearZeros projection of y), batch = [x, cond, weights], per-sample loss weights, frozen WIDE teacher (hidden 128:
    tensor-core GEMM path, bf16 operands) and a narrow student (fused fp32 kernels), perceptual term through both
    inverse passes. nll is all-fp32 (1e-5); kd / perceptual / loss and the gradients inherit the teacher's bf16 MLP
    (2e-2 of the largest entry)."""
    import nf_distillation_b200.pl_module as PM
    d = load("kd1d_rich_t2_s2")
    s_cfg, t_cfg = cfg_of(d, "s_cfg"), cfg_of(d, "t_cfg")
    w = json.loads(str(d["weights"]))
    cfg = nf_config(s_cfg, t_cfg, w, "rich")
    cfg["data"]["drop_weights"] = False
    m = PM.NFModel(cfg)
    m.student.load_state_dict(state_dict_of(d, "s_sd."))
    m.teacher.load_state_dict(state_dict_of(d, "t_sd."))
    m = m.to(dev)
    assert m.student_kd_indices == list(d["s_idx"]) and m.teacher_kd_indices == list(d["t_idx"])
    lat = t(d["latent"]).to(dev)
    monkeypatch.setattr(PM, "gaussian_sample", lambda mean, logs, T: lat)
    out = m.training_step([t(d["x"]).to(dev), t(d["cond"]).to(dev), t(d["sample_w"]).to(dev)], 0)
    assert abs(out["nll"].item() - float(d["nll"])) < 1e-5 * abs(float(d["nll"]))
    for k_ in ("kd", "perceptual", "loss"):
        assert abs(out[k_].item() - float(d[k_])) < 2e-2 * abs(float(d[k_])) + 1e-6, (k_, out[k_].item(), float(d[k_]))
    out["loss"].backward()
    assert all(p.grad is None for p in m.teacher.parameters())
    for n_, p in m.student.named_parameters():
        ref = t(d["grad." + n_])
        if ref.abs().max() == 0:
            continue
        assert p.grad is not None and rel(p.grad, ref) < 2e-2, n_
