"""TEST INFRASTRUCTURE ONLY — hand-derived known-answer vectors for the MAF / MADE path (tests/golden/maf_kat_d3.npz).

The reference ships no MAF code (README.md:7 names the model only), so nothing of the reference can pin the MADE
kernels. This script pins them to something that is NOT a restatement of the masked-MLP code: a D = 3 MADE layer whose
weights are zero except for six hand-placed connections, so that the autoregressive conditioner is a closed-form
scalar expression (Papamakarios et al. 2017, eq. 3-4:  u_d = (x_d - mu_d(x_<d)) exp(-alpha_d(x_<d)),
log|det| = -sum_d alpha_d ; inverse x_d = u_d exp(alpha_d) + mu_d, d = 1..D in order). The expressions below are
evaluated directly in numpy float64 — no masks, no matrix products — and written out with the weight tensors that
realise them in the module's state_dict layout. All weights and inputs are dyadic rationals that bf16 represents
exactly, so the tensor-core products are exact and the expected agreement is fp32 rounding (1e-6).

Degrees for D = 3, H = 64 (Germain et al. 2015: hidden unit k gets m(k) in 1..D-1; here units 0-31 have m = 1, units
32-63 have m = 2; input d connects to unit k iff m(k) >= d; unit k connects to unit k' iff m(k') >= m(k); unit k
connects to output d iff d > m(k)). The hand-placed units:

    a  = layer-1 unit 0  (m=1):  a  = relu(x1 + 1/2)
    b  = layer-1 unit 32 (m=2):  b  = relu(x1 - x2)
    a' = layer-2 unit 0  (m=1):  a' = relu(2 a - 1)
    b' = layer-2 unit 32 (m=2):  b' = relu(a + b / 2)
    mu1 = -1/2              alpha1 = 1/4                    (biases: output 1 sees no hidden unit)
    mu2 = a'/2 + 1/4        alpha2 = -a'/2                  (output 2 sees m = 1 units only)
    mu3 = b' - a'           alpha3 = b'/4 + 1/8             (output 3 sees every unit)

Run:  python oracle/make_maf_kat.py
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D, H = 3, 64


def conditioner(x1, x2):
    r = lambda v: np.maximum(v, 0.0)
    a, b = r(x1 + 0.5), r(x1 - x2)
    a2, b2 = r(2 * a - 1), r(a + b / 2)
    mu = (np.full_like(x1, -0.5), a2 / 2 + 0.25, b2 - a2)
    al = (np.full_like(x1, 0.25), -a2 / 2, b2 / 4 + 0.125)
    return mu, al


def made_forward(x):
    """x [N, 3] -> (u [N, 3] in x order, logdet [N])."""
    mu, al = conditioner(x[:, 0], x[:, 1])
    u = np.stack([(x[:, d] - mu[d]) * np.exp(-al[d]) for d in range(3)], 1)
    return u, -(al[0] + al[1] + al[2])


def made_inverse(u):
    """u [N, 3] in x order -> x, one coordinate at a time."""
    x1 = u[:, 0] * np.exp(0.25) - 0.5
    r = lambda v: np.maximum(v, 0.0)
    a2 = r(2 * r(x1 + 0.5) - 1)
    x2 = u[:, 1] * np.exp(-a2 / 2) + (a2 / 2 + 0.25)
    mu, al = conditioner(x1, x2)
    x3 = u[:, 2] * np.exp(al[2]) + mu[2]
    return np.stack([x1, x2, x3], 1)


def weights():
    w1, b1 = np.zeros((H, D)), np.zeros(H)
    w2, b2 = np.zeros((H, H)), np.zeros(H)
    w3, b3 = np.zeros((2 * D, H)), np.zeros(2 * D)
    w1[0, 0], b1[0] = 1.0, 0.5                    # a
    w1[32, 0], w1[32, 1] = 1.0, -1.0              # b
    w2[0, 0], b2[0] = 2.0, -1.0                   # a'
    w2[32, 0], w2[32, 32] = 1.0, 0.5              # b'
    b3[0], b3[3] = -0.5, 0.25                     # mu1, alpha1
    w3[1, 0], b3[1] = 0.5, 0.25                   # mu2
    w3[4, 0] = -0.5                               # alpha2
    w3[2, 32], w3[2, 0] = 1.0, -1.0               # mu3
    w3[5, 32], b3[5] = 0.25, 0.125                # alpha3
    # decoys on connections the masks must remove: if a kernel ignored a mask these would change the answer
    w1[0, 1] = 3.0        # unit a (m=1) must not see x2
    w1[32, 2] = -2.0      # no unit sees x3
    w2[0, 32] = 1.5       # a' (m=1) must not see b (m=2)
    w3[1, 32] = 2.0       # mu2 must not see b' (m=2)
    w3[0, 0] = 1.0        # mu1 sees nothing
    deg = np.repeat(np.array([1, 2], dtype=np.int32), 32)
    return dict(w1=w1, b1=b1, w2=w2, b2=b2, w3=w3, b3=b3, deg=deg)


def main():
    # dyadic inputs covering both sides of every ReLU
    g = np.array([-1.0, -0.5, -0.25, 0.0, 0.25, 0.5, 1.0, 1.5])
    x = np.array([(p, q, s) for p in g for q in g[::2] for s in (-0.75, 0.125, 2.0)], dtype=np.float64)
    u1, ld1 = made_forward(x)
    assert np.abs(made_inverse(u1) - x).max() < 1e-12
    # the module reverses the feature order after every layer (MAF's order alternation): two layers, same weights
    y = u1[:, ::-1]
    u2, ld2 = made_forward(y)
    z = u2[:, ::-1]
    nll = -((ld1 + ld2) + (-0.5 * (z ** 2 + np.log(2 * np.pi))).sum(1))
    w = weights()
    out = {"x": x, "u_layer1_x_order": u1, "logdet_layer1": ld1, "out_layer1": y, "out_layer2": z,
           "logdet_total": ld1 + ld2, "nll": nll, "D": np.int64(D), "H": np.int64(H)}
    for l in range(2):
        pre = f"sd.flow.layers.{l}."
        out[pre + "fc1.weight"], out[pre + "fc1.bias"] = w["w1"].astype(np.float32), w["b1"].astype(np.float32)
        out[pre + "fc2.weight"], out[pre + "fc2.bias"] = w["w2"].astype(np.float32), w["b2"].astype(np.float32)
        out[pre + "fc3.weight"], out[pre + "fc3.bias"] = w["w3"].astype(np.float32), w["b3"].astype(np.float32)
        out[pre + "deg1"], out[pre + "deg2"] = w["deg"], w["deg"]
    out["sd.prior_h"] = np.zeros((1, 2 * D), dtype=np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "maf_kat_d3.npz"), **out)
    print("wrote tests/golden/maf_kat_d3.npz:", x.shape[0], "samples")


if __name__ == "__main__":
    main()
