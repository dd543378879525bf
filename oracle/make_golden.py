"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/models, imported read-only) on seeded inputs and randomised weights.

Run in the build container (the reference does not exist on the GPU box):
    python oracle/make_golden.py

Each fixture stores the config (JSON), the full reference ``state_dict`` (reference key names), the inputs
(pre-noise x, the dequantisation noise the reference drew, latents) and the reference outputs: every layer output
the KD path can tap, bpd / nll, the deterministic (temperature 0) inverse, per-layer forward+reverse log-dets, and
for the KD fixtures the four loss scalars of NFModel.loss plus student gradients of a few parameters.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("NF_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import models  # noqa: F401  (reference package)
    from models import create_glow_model
    from models.flows import FlowStep
    from models.layers import SqueezeLayer, Split2d
    return create_glow_model, FlowStep, SqueezeLayer, Split2d


def base_cfg(**kw):
    cfg = dict(image_shape=[32, 32, 3], hidden_channels=64, K=2, L=3, actnorm_scale=1.0,
               flow_permutation="invconv", flow_coupling="affine", LU_decomposed=True, y_classes=10,
               learn_top=False, y_condition=False, is_1d=False)
    cfg.update(kw)
    return cfg


def randomise(model, gen, std=0.05):
    """Zero-initialised tensors (Conv2dZeros, ActNorm bias/logs, LinearZeros) make parity vacuous:
    re-draw them, and perturb the LU factors, so every term of the hot path is exercised."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("actnorm.bias", "actnorm.logs", ".logs")) or ".block.4." in name or \
                    name.startswith("learn_top_fn.") or \
                    (".conv.conv." in name) or name.endswith(("conv.logs",)):
                p.copy_(torch.randn(p.shape, generator=gen) * std)
            elif name.endswith(("invconv.lower", "invconv.upper", "invconv.log_s")):
                p.add_(torch.randn(p.shape, generator=gen) * std)
            elif ".block.10." in name:  # last Linear of the 1-D coupling MLP: default init is fine but small
                p.mul_(0.5)
            elif name.startswith(("project_ycond.linear", "project_class.linear")):   # LinearZeros of y_condition
                p.copy_(torch.randn(p.shape, generator=gen) * std)


def make_x(cfg, B, gen):
    if cfg["is_1d"]:
        return torch.randn(B, cfg["image_shape"][0], generator=gen)
    H, W, C = cfg["image_shape"]
    return torch.floor(torch.rand(B, C, H, W, generator=gen) * 256.0) / 256.0 - 0.5


def kd_indices_reference(student, teacher, is_1d, SqueezeLayer):
    """Restates pl_module.py:81-110 on the reference modules (pl_module itself needs Lightning, absent here)."""
    s_idx = [i for i, l in enumerate(student.flow.layers)
             if isinstance(l, SqueezeLayer) or (is_1d and (i + 1) % 2 == 0) or i + 1 == len(student.flow.layers)]
    t_idx = [i for i, l in enumerate(teacher.flow.layers)
             if isinstance(l, SqueezeLayer) or (is_1d and (i + 1) % 4 == 0) or i + 1 == len(teacher.flow.layers)]
    return s_idx, t_idx


def forward_fixture(name, cfg, B, seed):
    create_glow_model, FlowStep, SqueezeLayer, Split2d = import_reference()
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    model = create_glow_model(dict(cfg))
    randomise(model, gen)
    model.eval()
    if cfg["flow_permutation"] != "invconv":   # Permute2d.indices are plain attributes: record them with the config
        cfg = dict(cfg, perm_indices={str(i): getattr(l, cfg["flow_permutation"]).indices.tolist()
                                      for i, l in enumerate(model.flow.layers) if isinstance(l, FlowStep)})
    x0 = make_x(cfg, B, gen)
    x = x0.clone()
    torch.manual_seed(seed + 2)
    with torch.no_grad():
        outs, bpd, _ = model(x, None)          # mutates x in place with the dequantisation noise (2-D)
    noise = x - x0
    data = {"x": x0.numpy(), "noise": noise.numpy(), "bpd": bpd.numpy()}
    for i, o in enumerate(outs):
        data[f"out.{i}"] = o.numpy()
    # deterministic inverse from the final latent (temperature 0 => Split2d returns its mean)
    with torch.no_grad():
        rev = model(z=outs[-1].clone(), temperature=0.0, reverse=True)
    data["rev_last"] = rev[-1].numpy()
    data["rev_n"] = np.array(len(rev))
    # per-layer log-dets, forward and reverse, on the first FlowStep of each level
    with torch.no_grad():
        inp = x.clone()
        seen = set()
        for i, layer in enumerate(model.flow.layers):
            zero = torch.zeros(B)
            out, ld = layer(inp, logdet=zero, reverse=False)
            if isinstance(layer, FlowStep) and inp.shape[1] not in seen:
                seen.add(inp.shape[1])
                back, ld_r = layer(out, logdet=torch.zeros(B), reverse=True)
                data[f"step.{i}.logdet_fwd"] = ld.numpy()
                data[f"step.{i}.logdet_rev"] = ld_r.numpy()
                data[f"step.{i}.roundtrip"] = back.numpy()
            inp = out
    for k, v in model.state_dict().items():
        data["sd." + k] = v.numpy()
    data["cfg"] = np.array(json.dumps(cfg))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, "layers", len(outs), "bpd", bpd.numpy()[:2])


def kd_fixture(name, s_cfg, t_cfg, B, seed, weights, sample_weights=False):
    create_glow_model, FlowStep, SqueezeLayer, Split2d = import_reference()
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    student, teacher = create_glow_model(dict(s_cfg)), create_glow_model(dict(t_cfg))
    randomise(student, gen)
    randomise(teacher, gen)
    is_1d = s_cfg["is_1d"]
    s_idx, t_idx = kd_indices_reference(student, teacher, is_1d, SqueezeLayer)
    x0 = make_x(s_cfg, B, gen)
    x = x0.clone()
    cond = None
    if s_cfg["y_condition"]:   # RICH: batch = [x, cond, weights] (pl_module.py:212-213), cond one-hot over y_classes
        cond = torch.nn.functional.one_hot(torch.randint(0, s_cfg["y_classes"], (B,), generator=gen),
                                           s_cfg["y_classes"]).float()
    torch.manual_seed(seed + 2)
    # --- pl_module.py:198-255 (forward), restated without Lightning
    s_z, s_nll, _ = student(x, cond)
    noise_s = (x - x0).clone()
    with torch.no_grad():
        t_z, _, _ = teacher(x, cond)
    noise_t = (x - x0) - noise_s
    data = {"x": x0.numpy(), "noise_s": noise_s.numpy(), "noise_t": noise_t.numpy(),
            "s_idx": np.array(s_idx), "t_idx": np.array(t_idx)}
    perceptual = torch.tensor(0.0)
    if weights["perceptual"] > 0:
        mean, logs = student.prior(x, y_onehot=cond)
        latent = (torch.randn(mean.shape, generator=gen) * torch.exp(logs) + mean).detach()
        sx = student(z=latent, temperature=0.7, reverse=True, y_onehot=cond)[-1]
        with torch.no_grad():
            tx = teacher(z=latent, temperature=0.7, reverse=True, y_onehot=cond)[-1]
        data["latent"] = latent.numpy()
        data["student_x"] = sx.detach().numpy()
        perceptual = torch.nn.functional.l1_loss(sx, tx, reduction="none")
        perceptual = perceptual.mean(dim=list(range(perceptual.dim()))[1:])
    # --- pl_module.py:257-320 (loss)
    kd = None
    for si, ti in zip(s_idx, t_idx):
        part = torch.nn.functional.mse_loss(s_z[si], t_z[ti], reduction="none")
        part = part.mean(dim=list(range(part.dim()))[1:])
        kd = part if kd is None else kd + part
    kd = kd / len(s_idx)
    result = weights["nll"] * s_nll + weights["kd"] * kd + weights["perceptual"] * perceptual
    if cond is not None:
        data["cond"] = cond.numpy()
    if sample_weights:          # pl_module.py:311-313
        sw = torch.rand(B, generator=gen) + 0.5
        data["sample_w"] = sw.numpy()
        result = result * sw
    loss = result.mean()
    loss.backward()
    data.update(nll=s_nll.mean().detach().numpy(), kd=kd.mean().detach().numpy(),
                perceptual=perceptual.mean().detach().numpy(), loss=loss.detach().numpy())
    for i in s_idx:
        data[f"student_z.{i}"] = s_z[i].detach().numpy()
    for n, p in student.named_parameters():
        data["grad." + n] = p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)
    for k, v in student.state_dict().items():
        data["s_sd." + k] = v.numpy()
    for k, v in teacher.state_dict().items():
        data["t_sd." + k] = v.numpy()
    data["s_cfg"] = np.array(json.dumps(s_cfg))
    data["t_cfg"] = np.array(json.dumps(t_cfg))
    data["weights"] = np.array(json.dumps(weights))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, "taps", s_idx, t_idx, "loss", float(loss))


def main(only=None):
    # 2-D Glow, CIFAR-shaped, small K / hidden so the fixture stays small (reference default config otherwise)
    forward_fixture("glow2d_cifar_k2_h64", base_cfg(K=2, L=3, hidden_channels=64), B=4, seed=42)
    # 2-D Glow, 16x16, L=2 (exercises one Split2d), odd batch
    forward_fixture("glow2d_16_k1_h64", base_cfg(image_shape=[16, 16, 3], K=1, L=2, hidden_channels=64), B=3, seed=7)
    # the optional 2-D variants no shipped config uses (flows.py:85-95,157-158): additive coupling, fixed permutations
    forward_fixture("glow2d_16_additive_shuffle_k2_h64",
                    base_cfg(image_shape=[16, 16, 3], K=2, L=2, hidden_channels=64, flow_permutation="shuffle",
                             flow_coupling="additive"), B=3, seed=17)
    forward_fixture("glow2d_16_affine_reverse_k2_h64",
                    base_cfg(image_shape=[16, 16, 3], K=2, L=2, hidden_channels=64, flow_permutation="reverse"),
                    B=3, seed=19)
    # learn_top: the prior's mean / logs come from a Conv2dZeros over the (all-zero) prior buffer
    forward_fixture("glow2d_16_learntop_k1_h64",
                    base_cfg(image_shape=[16, 16, 3], K=1, L=2, hidden_channels=64, learn_top=True), B=3, seed=29)
    # 1-D Glow (tabular): POWER-shaped D=6 and BSDS300-shaped D=63 (odd D: z1=31, z2=32)
    forward_fixture("glow1d_d6_k5_h32", base_cfg(image_shape=[6], K=5, L=1, hidden_channels=32, is_1d=True,
                                                  y_classes=0), B=64, seed=11)
    forward_fixture("glow1d_d63_k5_h32", base_cfg(image_shape=[63], K=5, L=1, hidden_channels=32, is_1d=True,
                                                   y_classes=0), B=32, seed=13)
    # KD steps (pl_module.py forward+loss), image weights of conf/training/cifar.yaml, tabular of tabular.yaml
    kd_fixture("kd2d_cifar_t4_s2_h64", base_cfg(K=2, hidden_channels=64), base_cfg(K=4, hidden_channels=64), B=4,
               seed=21, weights={"nll": 0.9, "kd": 0.1, "perceptual": 0.0})
    kd_fixture("kd1d_d63_t5_s3", base_cfg(image_shape=[63], K=3, L=1, hidden_channels=16, is_1d=True, y_classes=0),
               base_cfg(image_shape=[63], K=5, L=1, hidden_channels=32, is_1d=True, y_classes=0), B=64, seed=23,
               weights={"nll": 0.85, "kd": 0.05, "perceptual": 0.1})
    # RICH-shaped (conf/{teacher,student,training}/rich.yaml): D=5, y-conditioned on 3 classes, wide frozen teacher,
    # per-sample loss weights; small K / teacher width 128 so the fixture stays small
    rich = dict(image_shape=[5], is_1d=True, y_classes=3, y_condition=True)
    kd_fixture("kd1d_rich_t2_s2", base_cfg(K=2, L=1, hidden_channels=32, **rich),
               base_cfg(K=1, L=2, hidden_channels=128, **rich), B=96, seed=31,
               weights={"nll": 0.85, "kd": 0.075, "perceptual": 0.075}, sample_weights=True)


if __name__ == "__main__":
    main()
