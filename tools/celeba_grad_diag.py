import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200.models import create_glow_model, utils as U
from nf_distillation_b200.train import glow_cfg, randomise_zero_params
from oracle import glow_oracle as O
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
hid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = glow_cfg((64, 64, 3), 1, 4, hid)
torch.manual_seed(11)
m = create_glow_model(cfg); randomise_zero_params(m, 12, 0.05)
sd = {k: v.clone() for k, v in m.state_dict().items()}
m = m.to(dev).train()
g = torch.Generator().manual_seed(3)
x = torch.floor(torch.rand(B, 3, 64, 64, generator=g) * 256) / 256 - 0.5
noise = torch.rand(B, 3, 64, 64, generator=g) / 256
wz = torch.randn(B, 96, 4, 4, generator=g)
names = dict(m.named_parameters())
osd = {k: v.clone().requires_grad_(k in names) for k, v in sd.items()}
o_outs, o_bpd = O.glow_forward(osd, cfg, x, noise)
(o_bpd.sum() + (o_outs[-1] * wz).sum() * 1e-2).backward()
U.dequant_noise = lambda t, n: noise.to(dev)
outs, bpd, _ = m(x.to(dev), None)
(bpd.sum() + (outs[-1] * wz.to(dev)).sum() * 1e-2).backward()
rows = []
for n_, p in m.named_parameters():
    og = osd[n_].grad
    if og is None: continue
    a, b = p.grad.detach().cpu().flatten().double(), og.flatten().double()
    rel = ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()
    cos = (a @ b / (a.norm() * b.norm() + 1e-30)).item()
    rows.append((rel, cos, b.abs().max().item(), n_))
rows.sort(reverse=True)
for r in rows[:14]: print("%.4f cos=%.5f max=%.3e %s" % r)
