"""TEST INFRASTRUCTURE ONLY — plain-PyTorch fp32 MAF/MADE, PARITY UNPINNED.

The reference repository ships no MAF/MADE code (README.md:7 names the model, nothing implements it; pl_module.py:153-156
rejects every architecture but "glow"), so there are no reference outputs to pin this against. It restates the
published algorithm (Papamakarios et al. 2017, eq. 3-4; Germain et al. 2015 masks) with the SAME mask / degree / flip
conventions as nf_distillation_b200/models/maf.py and is only a self-consistency checker for the CUDA path."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def masks(D, deg1, deg2):
    i = torch.arange(1, D + 1)
    m1 = (deg1[:, None] >= i[None, :]).float()
    m2 = (deg2[:, None] >= deg1[None, :]).float()
    r = (torch.arange(2 * D) % D + 1)
    m3 = (r[:, None] > deg2[None, :]).float()
    return m1, m2, m3


def made_net(x, sd, pre, D):
    m1, m2, m3 = masks(D, sd[pre + "deg1"].long(), sd[pre + "deg2"].long())
    h = torch.relu(F.linear(x, sd[pre + "fc1.weight"] * m1, sd[pre + "fc1.bias"]))
    h = torch.relu(F.linear(h, sd[pre + "fc2.weight"] * m2, sd[pre + "fc2.bias"]))
    out = F.linear(h, sd[pre + "fc3.weight"] * m3, sd[pre + "fc3.bias"])
    return out[:, :D], out[:, D:]


def made_forward(x, sd, pre, D, flip=True):
    mu, alpha = made_net(x, sd, pre, D)
    u = (x - mu) * torch.exp(-alpha)
    return (u.flip(1) if flip else u), -alpha.sum(1)


def made_inverse(u_out, sd, pre, D, flip=True):
    u = u_out.flip(1) if flip else u_out
    x = torch.zeros_like(u)
    for i in range(D):
        mu, alpha = made_net(x, sd, pre, D)
        x = x.clone()
        x[:, i] = u[:, i] * torch.exp(alpha[:, i]) + mu[:, i]
    _, alpha = made_net(x, sd, pre, D)
    return x, alpha.sum(1)


def maf_forward(sd, D, n_layers, x):
    outs, ld = [], torch.zeros(x.shape[0], dtype=x.dtype)
    z = x
    for l in range(n_layers):
        z, d = made_forward(z, sd, f"flow.layers.{l}.", D)
        ld = ld + d
        outs.append(z)
    nll = -(ld + (-0.5 * (z ** 2 + math.log(2 * math.pi))).sum(1))
    return outs, nll


def maf_inverse(sd, D, n_layers, z):
    for l in reversed(range(n_layers)):
        z, _ = made_inverse(z, sd, f"flow.layers.{l}.", D)
    return z


def hidden_degrees(D, H):
    """Sorted degrees 1..D-1, each ~H/(D-1) times (Germain et al. 2015, eq. 12-13, deterministic assignment); changes
    fall on multiples of 8 units when there are at least D-1 such groups (same rule as the product's models/maf.py)."""
    if D <= 1:
        return torch.ones(H, dtype=torch.int32)
    if H % 8 == 0 and H // 8 >= D - 1:
        tiles = H // 8
        return (torch.arange(tiles, dtype=torch.int64) * (D - 1) // tiles + 1).repeat_interleave(8).to(torch.int32)
    return (torch.arange(H, dtype=torch.int64) * (D - 1) // H + 1).to(torch.int32)


def random_state_dict(D, H, n_layers, seed):
    """Seeded nn.Linear-style weights for `n_layers` MADE layers (timing input for bench.py's reference arm)."""
    g = torch.Generator().manual_seed(seed)
    u = lambda *s, fan: (torch.rand(*s, generator=g) * 2 - 1) / math.sqrt(fan)
    sd = {}
    for l in range(n_layers):
        pre = f"flow.layers.{l}."
        deg = hidden_degrees(D, H)
        sd[pre + "deg1"], sd[pre + "deg2"] = deg.clone(), deg.clone()
        sd[pre + "fc1.weight"], sd[pre + "fc1.bias"] = u(H, D, fan=D), u(H, fan=D)
        sd[pre + "fc2.weight"], sd[pre + "fc2.bias"] = u(H, H, fan=H), u(H, fan=H)
        sd[pre + "fc3.weight"], sd[pre + "fc3.bias"] = u(2 * D, H, fan=H) * 0.1, torch.zeros(2 * D)
    return sd
