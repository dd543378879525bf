"""GPU diagnostic: run the CUDA modules on the committed golden fixtures and print error magnitudes."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import load, cfg_of, state_dict_of, t
from nf_distillation_b200.models import create_glow_model
from nf_distillation_b200.models import utils as U
from nf_distillation_b200.pl_module import NFModel

dev = "cuda"


def rel(a, b):
    b = b.to(a.device)
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def fwd(name):
    d = load(name)
    cfg, sd = cfg_of(d), state_dict_of(d)
    m = create_glow_model(cfg)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    x = t(d["x"]).to(dev)
    if not cfg["is_1d"]:
        noise = t(d["noise"]).to(dev)
        U.dequant_noise = lambda x_, n: noise
    with torch.no_grad():
        outs, bpd, _ = m(x.clone(), None)
    print(f"== {name}: bpd rel err {rel(bpd, t(d['bpd'])):.3e}  bpd {bpd[:2].tolist()} ref {d['bpd'][:2].tolist()}")
    for i, o in enumerate(outs):
        print(f"   out.{i} shape {tuple(o.shape)} rel err {rel(o, t(d[f'out.{i}'])):.3e}")
    n = len(outs)
    with torch.no_grad():
        rev = m(z=t(d[f"out.{n-1}"]).to(dev), temperature=0.0, reverse=True)
    print(f"   reverse last rel err {rel(rev[-1], t(d['rev_last'])):.3e}")
    # per-step logdets
    with torch.no_grad():
        inp = x + (t(d["noise"]).to(dev) if not cfg["is_1d"] else 0)
        for i, layer in enumerate(m.flow.layers):
            out, ld = layer(inp, logdet=torch.zeros(x.shape[0], device=dev), reverse=False)
            if f"step.{i}.logdet_fwd" in d:
                back, ldr = layer(out, logdet=torch.zeros(x.shape[0], device=dev), reverse=True)
                print(f"   step {i}: logdet fwd rel {rel(ld, t(d[f'step.{i}.logdet_fwd'])):.3e} rev rel "
                      f"{rel(ldr, t(d[f'step.{i}.logdet_rev'])):.3e} roundtrip rel {rel(back, inp):.3e}")
            inp = t(d[f"out.{i}"]).to(dev)


def kd(name):
    d = load(name)
    s_cfg, t_cfg = cfg_of(d, "s_cfg"), cfg_of(d, "t_cfg")
    w = json.loads(str(d["weights"]))
    cfg = {"data": {"name": "cifar" if not s_cfg["is_1d"] else "bsds300"}, "student": dict(s_cfg), "teacher": dict(t_cfg),
           "loss": {"nll": {"weight": w["nll"]}, "kd": {"weight": w["kd"], "name": "mse"},
                    "perceptual": {"weight": w["perceptual"], "name": "l1"}},
           "optimizer": "adam", "learning_rate": 5e-4, "weight_decay": 0.0}
    m = NFModel(cfg)
    m.student.load_state_dict(state_dict_of(d, "s_sd."))
    m.teacher.load_state_dict(state_dict_of(d, "t_sd."))
    m = m.to(dev)
    assert m.student_kd_indices == list(d["s_idx"]) and m.teacher_kd_indices == list(d["t_idx"])
    x = t(d["x"]).to(dev)
    noises = [t(d["noise_s"]).to(dev), t(d["noise_t"]).to(dev)]
    U.dequant_noise = lambda x_, n: noises.pop(0)
    if "latent" in d:
        import nf_distillation_b200.pl_module as PM
        lat = t(d["latent"]).to(dev)
        PM.gaussian_sample = lambda mean, logs, T: lat
    batch = [x] if s_cfg["is_1d"] else [x, None]
    out = m.training_step(batch)
    for k_, ref in (("nll", "nll"), ("kd", "kd"), ("perceptual", "perceptual"), ("loss", "loss")):
        print(f"   {k_}: {out[k_].item():.6f} ref {float(d[ref]):.6f} rel {abs(out[k_].item()-float(d[ref]))/(abs(float(d[ref]))+1e-12):.3e}")
    out["loss"].backward()
    worst = []
    for n_, p in m.student.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        ref = t(d["grad." + n_]).to(dev)
        e = ((g - ref).abs().max() / (ref.abs().max() + 1e-12)).item()
        worst.append((e, n_, ref.abs().max().item()))
    worst.sort(reverse=True)
    print(f"== {name}: worst grad rel errs:")
    for e, n_, s in worst[:12]:
        print(f"   {e:.3e}  {n_}  (ref max {s:.3e})")
    import statistics
    print("   median grad rel err", statistics.median([e for e, _, _ in worst]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["2d"]
    if "2d" in which:
        fwd("glow2d_cifar_k2_h64")
        fwd("glow2d_16_k1_h64")
        kd("kd2d_cifar_t4_s2_h64")
    if "1d" in which:
        fwd("glow1d_d6_k5_h32")
        fwd("glow1d_d63_k5_h32")
        kd("kd1d_d63_t5_s3")
