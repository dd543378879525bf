"""MAF / MADE pinned to something the builder's masked-MLP code did not produce.

The reference names MAF (README.md:7) but ships no code for it, so there is no reference output to record. Instead:
  * tests/golden/maf_kat_d3.npz — hand-derived known answers (oracle/make_maf_kat.py: a D = 3 MADE layer reduced to
    closed-form scalar expressions of Papamakarios et al. 2017 eq. 3-4, evaluated in float64 with no masks and no
    matrix products, plus decoy weights on every kind of connection a mask must remove);
  * the defining properties of Germain et al. 2015's construction, checked on the masks themselves and on the
    Jacobian.
The CPU tests hold oracle/maf_oracle.py to these; the GPU tests hold the CUDA path (csrc/maf.cu, maf_inverse.cu) to them.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from golden_util import load, state_dict_of, t  # noqa: E402


def kat():
    d = load("maf_kat_d3")
    return d, state_dict_of(d)


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_maf_oracle_reproduces_the_hand_derived_known_answers():
    from oracle import maf_oracle as MO
    d, sd = kat()
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    x = t(d["x"])
    u, ld = MO.made_forward(x, sd64, "flow.layers.0.", 3, flip=False)
    assert (u - t(d["u_layer1_x_order"])).abs().max() < 1e-12
    assert (ld - t(d["logdet_layer1"])).abs().max() < 1e-12
    outs, nll = MO.maf_forward(sd64, 3, 2, x)
    assert (outs[0] - t(d["out_layer1"])).abs().max() < 1e-12
    assert (outs[1] - t(d["out_layer2"])).abs().max() < 1e-12
    assert (nll - t(d["nll"])).abs().max() < 1e-11
    assert (MO.maf_inverse(sd64, 3, 2, outs[1]) - x).abs().max() < 1e-11


@pytest.mark.parametrize("D,H", [(2, 64), (3, 64), (6, 512), (21, 128), (63, 512), (63, 192), (100, 64)])
def test_masks_have_the_made_autoregressive_structure(D, H):
    """Germain et al. 2015, eq. 12-13: with degrees m(k), M1[k,d] = [m(k) >= d], M2[k',k] = [m(k') >= m(k)],
    M3[d,k] = [d > m(k)]. Then (a) output d depends on input d' ONLY IF d' < d: M3 M2 M1 is strictly lower triangular,
    and (b) when every degree 1..D-1 is present (H >= D-1) it depends on EVERY d' < d (full connectivity, no capacity
    silently dropped). Checked for the module's degree assignment and the oracle's mask rule; the GPU test below
    checks the CUDA prep kernel's zero pattern against the same rule."""
    from nf_distillation_b200.models.maf import hidden_degrees
    from oracle import maf_oracle as MO
    deg = hidden_degrees(D, H).long()
    assert deg.min() >= 1 and deg.max() <= max(D - 1, 1) and bool((deg[1:] >= deg[:-1]).all())
    assert torch.equal(deg, MO.hidden_degrees(D, H).long())
    if H >= D - 1 and D > 1:
        assert set(deg.tolist()) == set(range(1, D)), "every degree must occur when there are enough units"
    m1, m2, m3 = MO.masks(D, deg, deg)
    for half in (m3[:D], m3[D:]):                       # mu rows and alpha rows
        conn = half @ m2 @ m1                           # [out d, in d'] number of paths
        assert torch.equal(conn.triu(0), torch.zeros_like(conn)), "output d must not see inputs >= d"
        if H >= D - 1:
            assert bool((conn.tril(-1)[torch.tril(torch.ones(D, D), -1).bool()] > 0).all())


def test_maf_oracle_jacobian_is_triangular_and_logdet_is_its_log_determinant():
    from oracle import maf_oracle as MO
    D, H = 5, 64
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in MO.random_state_dict(D, H, 1, 3).items()}
    x = torch.randn(D, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    f = lambda v: MO.made_forward(v[None], sd, "flow.layers.0.", D, flip=False)[0][0]
    J = torch.autograd.functional.jacobian(f, x)
    assert J.triu(1).abs().max() < 1e-14
    _, ld = MO.made_forward(x[None], sd, "flow.layers.0.", D, flip=False)
    assert abs(torch.linalg.slogdet(J)[1].item() - ld.item()) < 1e-10


# ------------------------------------------------------------------------------------------------ GPU: the kernels
@pytest.mark.gpu
def test_cuda_maf_reproduces_the_hand_derived_known_answers():
    """The CUDA path against the float64 closed forms. Layer 1: every weight and input is exactly representable in
    bf16, so the tensor-core products are exact and what is left is fp32 rounding of exp() and the affine (2e-6) —
    forward u, log-det, and both inverse kernels (whose recursion feeds back dyadic x_d, again exact). Layer 2 takes
    layer 1's output, which bf16 rounds on its way into the masked GEMM: conditioner inputs carry 2^-9 relative
    rounding, so the two-layer outputs / nll are held to 1e-2."""
    from nf_distillation_b200.models.maf import create_maf_model
    d, sd = kat()
    m = create_maf_model(dict(image_shape=[3], hidden_channels=64, K=2))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = t(d["x"]).float().cuda()
    rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
    layer1 = m.flow.layers[0]
    with torch.no_grad():
        u, ld = layer1(x, logdet=torch.zeros(x.shape[0], device="cuda"))
        assert rel(u, t(d["out_layer1"])) < 2e-6 and rel(ld, t(d["logdet_layer1"])) < 2e-6
        for resident in (True, False):                  # shared-memory-resident inverse and the D-pass GEMM inverse
            layer1.resident_inverse = resident
            back, ldb = layer1(t(d["out_layer1"]).float().cuda(), logdet=ld, reverse=True)
            assert rel(back, t(d["x"])) < 5e-6, resident
            assert ldb.abs().max().item() < 1e-5, resident          # logdet(fwd) + logdet(inverse) = 0
        outs, nll, _ = m(x, None)
        assert rel(outs[0], t(d["out_layer1"])) < 2e-6
        assert rel(outs[1], t(d["out_layer2"])) < 1e-2 and rel(nll, t(d["nll"])) < 1e-2
        assert rel(m(z=outs[1], reverse=True)[-1], t(d["x"])) < 1e-2
    # training path (activation-saving forward) gives the same numbers
    outs_t, nll_t, _ = m(x.clone().requires_grad_(True), None)
    assert rel(outs_t[0].detach(), t(d["out_layer1"])) < 2e-6 and rel(nll_t.detach(), t(d["nll"])) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("D,H", [(3, 64), (6, 512), (63, 512), (63, 192)])
def test_cuda_made_prep_applies_exactly_the_made_masks(D, H):
    """nfk_made_prep's bf16 operands: zero wherever Germain's rule says zero, the (rounded) weight elsewhere."""
    from nf_distillation_b200 import ops
    from nf_distillation_b200.models.maf import MADE
    from oracle import maf_oracle as MO
    torch.manual_seed(D + H)
    made = MADE(D, H, flip=True).cuda()
    with torch.no_grad():
        for p in made.parameters():
            p.copy_(torch.rand_like(p) + 0.5)           # no accidental zeros
    B1, _, B2, _, B3, _, _ = made._operands(tuple(p.detach() for p in made._params()), False)
    m1, m2, m3 = MO.masks(D, made.deg1.cpu().long(), made.deg2.cpu().long())
    assert torch.equal(B1[:, :D].float().cpu() != 0, m1.bool()) and (B1[:, D:] == 0).all()
    assert torch.equal(B2.float().cpu() != 0, m2.bool())
    assert torch.equal(B3[:2 * D].float().cpu() != 0, m3.bool()) and (B3[2 * D:] == 0).all()
    assert torch.equal(B2.float().cpu(), (made.fc2.weight.detach().cpu() * m2).bfloat16().float())


@pytest.mark.gpu
def test_cuda_maf_is_autoregressive_at_full_size():
    """BASELINE configs[1] shape (D = 63, hidden 512, B >= 8192: the fused two-GEMM kernel). Perturbing feature j must
    leave every u_i with i < j bit-identical (strictly triangular Jacobian), and the inverse must undo the forward."""
    from nf_distillation_b200.models.maf import MADE
    D, H, B = 63, 512, 8192
    torch.manual_seed(1)
    made = MADE(D, H, flip=False).cuda().eval()
    with torch.no_grad():
        made.fc3.weight.mul_(5.0)
        x = torch.randn(B, D, device="cuda")
        u0, ld0 = made(x, logdet=torch.zeros(B, device="cuda"))
        for j in (0, 1, 17, 31, 62):
            xp = x.clone()
            xp[:, j] += 0.37
            u1, _ = made(xp, logdet=torch.zeros(B, device="cuda"))
            assert torch.equal(u1[:, :j], u0[:, :j]), j
            assert (u1[:, j] - u0[:, j]).abs().min() > 0
        back, ldb = made(u0, logdet=ld0, reverse=True)
        assert ((back - x).abs().max() / x.abs().max()).item() < 2e-2      # bf16 operands in the conditioner
        assert ldb.abs().max().item() < 2e-2 * (ld0.abs().max().item() + 1)


# ------------------------------------------------------------------------------------------------ BASELINE shapes
def graph_kat(D):
    d = load(f"maf_kat_graph_d{D}")
    sd = {k[3:]: torch.from_numpy(v.copy()) for k, v in d.items() if k.startswith("sd.")}
    return d, sd


@pytest.mark.parametrize("D", [6, 63])
def test_maf_oracle_matches_the_graph_interpreter_at_baseline_shapes(D):
    """oracle/make_maf_kat_graph.py: a sparse MADE layer at D = 6 / 63, hidden 512 (BASELINE configs[0] / [1]) evaluated
    connection by connection in float64 — no masks, no matrix products — with decoy weights on illegal positions.
    The paper restatement must reproduce it, and the fixture's degrees must be the product's own assignment."""
    from nf_distillation_b200.models.maf import hidden_degrees
    from oracle import maf_oracle as MO
    d, sd = graph_kat(D)
    assert torch.equal(sd["deg1"].long(), hidden_degrees(D, 512).long())
    sd64 = {"l." + k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    x = t(d["x"])
    u, ld = MO.made_forward(x, sd64, "l.", D, flip=False)
    assert (u - t(d["u_x_order"])).abs().max() < 1e-12 and (ld - t(d["logdet"])).abs().max() < 1e-12
    xr, _ = MO.made_inverse(t(d["u_x_order"]), sd64, "l.", D, flip=False)
    assert (xr - x).abs().max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("D", [6, 63])
def test_cuda_made_layer_matches_the_graph_interpreter_at_baseline_shapes(D):
    """The CUDA MADE layer at the benchmark shapes against the float64 graph interpreter. All weights / inputs / hidden
    activations are dyadic and fit bf16, so the tensor-core products are exact: forward u and log-det 3e-6 through BOTH
    forward paths (the per-layer GEMMs with k-block skipping at B = 64, and the fused two-GEMM kernel with its skipped
    B2 tiles at B = 8256), the inverse through the push kernel, the pull kernel and the D-pass GEMM fallback."""
    from nf_distillation_b200.models.maf import MADE
    d, sd = graph_kat(D)
    made = MADE(D, 512, flip=False)
    made.load_state_dict(sd)
    made = made.cuda().eval()
    rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()
    x = t(d["x"]).float().cuda()
    u_ref, ld_ref = t(d["u_x_order"]), t(d["logdet"])
    with torch.no_grad():
        u, ld = made(x, logdet=torch.zeros(x.shape[0], device="cuda"))
        assert rel(u, u_ref) < 3e-6 and rel(ld, ld_ref) < 3e-6
        reps = 129
        xb = x.repeat(reps, 1)                                   # 8256 rows: the fused kernel's path
        ub, ldb = made(xb, logdet=torch.zeros(xb.shape[0], device="cuda"))
        assert rel(ub, u_ref.repeat(reps, 1)) < 3e-6 and rel(ldb, ld_ref.repeat(reps)) < 3e-6
        for push, resident in ((True, True), (False, True), (False, False)):
            made.push_inverse, made.resident_inverse = push, resident
            back, ld0 = made(u_ref.float().cuda(), logdet=ld_ref.float().cuda(), reverse=True)
            assert rel(back, t(d["x"])) < 1e-5, (push, resident)
            assert ld0.abs().max().item() < 1e-4, (push, resident)
    # training forward (activation-saving kernels) agrees too, and gradients reach only legal connections
    made.push_inverse = made.resident_inverse = True
    xg = x.clone().requires_grad_(True)
    ug, ldg = made(xg, logdet=torch.zeros(x.shape[0], device="cuda"))
    assert rel(ug.detach(), u_ref) < 3e-6
    (ug.sum() + ldg.sum()).backward()
    assert xg.grad.triu(0).shape == xg.grad.shape            # (shape sanity)
    from oracle import maf_oracle as MO
    m1, m2, m3 = MO.masks(D, made.deg1.cpu().long(), made.deg2.cpu().long())
    assert (made.fc1.weight.grad.cpu()[m1 == 0] == 0).all() and (made.fc2.weight.grad.cpu()[m2 == 0] == 0).all()
    assert (made.fc3.weight.grad.cpu()[m3 == 0] == 0).all()
