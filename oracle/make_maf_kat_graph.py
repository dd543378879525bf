"""TEST INFRASTRUCTURE ONLY — known answers for MADE layers at the BASELINE shapes (D = 6 / 63, hidden 512), computed by a
SCALAR GRAPH INTERPRETER that never forms a mask or a matrix product (tests/golden/maf_kat_graph_d{6,63}.npz).

oracle/make_maf_kat.py pins a D = 3 layer to closed-form expressions written out by hand. This script scales the same
idea to the shapes the benchmarks run: a sparse MADE layer is DEFINED as an explicit list of connections
    (layer, target unit, source unit, weight)        weights and inputs dyadic, so every product is exact in bf16
each of which is legal under Germain et al.'s degree rule by construction (the generator only draws sources whose
degree the target is allowed to see), and it is evaluated unit by unit in float64:
    h1[k]  = relu(b1[k] + sum w * x[src])        h2[k] = relu(b2[k] + sum w * h1[src])
    mu[d]  = b3[d]   + sum w * h2[src]           alpha[d] = b3[D+d] + sum w * h2[src]
    u[d]   = (x[d] - mu[d]) * exp(-alpha[d]),    log|det| = -sum alpha                     (Papamakarios 2017, eq. 3-4)
and inverted coordinate by coordinate (x[d] needs only x[<d]). The weight TENSORS written next to the answers contain
these connections PLUS decoys on illegal positions (a source the target must not see): an implementation that applied
a wrong mask, skipped a needed k-block or kept a forbidden one would change the answer.

The degrees are those of models/maf.py::hidden_degrees at these shapes (D-1 groups of units on 8-unit boundaries); they
are written into the fixture and asserted by the test, not imported from the product.

Run:  python oracle/make_maf_kat_graph.py
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def degrees(D, H):
    tiles = H // 8
    assert H % 8 == 0 and tiles >= D - 1
    return np.repeat(np.arange(tiles) * (D - 1) // tiles + 1, 8).astype(np.int32)


def build(D, H, seed, n1=160, n2=160, n3=6):
    rng = np.random.default_rng(seed)
    deg = degrees(D, H)
    conns = {1: [], 2: [], 3: []}           # (target, source, weight)
    w1, w2, w3 = np.zeros((H, D)), np.zeros((H, H)), np.zeros((2 * D, H))
    b1 = rng.integers(-2, 3, H) / 4.0
    b2 = rng.integers(-2, 3, H) / 4.0
    b3 = np.concatenate([rng.integers(-4, 5, D) / 8.0, rng.integers(-2, 3, D) / 16.0])
    # layer 1: unit k (degree m) may see inputs 1..m (0-based 0..m-1); two sources at most, weights +-1/2, +-1
    for k in rng.choice(H, n1, replace=False):
        for src in rng.choice(deg[k], min(2, deg[k]), replace=False):
            w = rng.choice([-1.0, -0.5, 0.5, 1.0])
            conns[1].append((k, src, w)); w1[k, src] = w
    active1 = sorted({k for k, _, _ in conns[1]})
    # layer 2: unit k' (degree m') may see layer-1 units of degree <= m'; ONE source, weight +-1 (keeps h2 in 8 bits)
    for k in rng.choice(H, n2, replace=False):
        legal = [s for s in active1 if deg[s] <= deg[k]]
        if legal:
            src = rng.choice(legal)
            w = rng.choice([-1.0, 1.0])
            conns[2].append((k, src, w)); w2[k, src] = w
    active2 = sorted({k for k, _, _ in conns[2]})
    # layer 3: output d (1-based d+1) may see layer-2 units of degree < d+1
    for r in range(2 * D):
        d = r % D
        legal = [s for s in active2 if deg[s] < d + 1]
        for src in (rng.choice(legal, min(n3, len(legal)), replace=False) if legal else []):
            w = rng.choice([-0.25, 0.25, 0.125, -0.125]) if r >= D else rng.choice([-0.5, 0.5, 0.25, -0.25])
            conns[3].append((r, src, w)); w3[r, src] = w
    # decoys: weights on ILLEGAL positions only
    for _ in range(400):
        k, s = rng.integers(H), rng.integers(D)
        if s + 1 > deg[k]:
            w1[k, s] = 3.0
        k, s = rng.integers(H), rng.integers(H)
        if deg[s] > deg[k]:
            w2[k, s] = -2.0
        r, s = rng.integers(2 * D), rng.integers(H)
        if not deg[s] < (r % D) + 1:
            w3[r, s] = 1.5
    return deg, conns, (w1, b1, w2, b2, w3, b3)


def evaluate(D, H, conns, biases, x):
    """(mu [N, D], alpha [N, D]) of the sparse layer, unit by unit."""
    b1, b2, b3 = biases
    N = x.shape[0]
    h1 = {k: np.full(N, b1[k]) for k in {t for t, _, _ in conns[1]}}
    for t, s, w in conns[1]:
        h1[t] = h1[t] + w * x[:, s]
    h1 = {k: np.maximum(v, 0.0) for k, v in h1.items()}
    h2 = {k: np.full(N, b2[k]) for k in {t for t, _, _ in conns[2]}}
    for t, s, w in conns[2]:
        h2[t] = h2[t] + w * h1[s]
    h2 = {k: np.maximum(v, 0.0) for k, v in h2.items()}
    out = np.tile(b3, (N, 1))
    for r, s, w in conns[3]:
        out[:, r] += w * h2[s]
    # units without an incoming connection still fire relu(bias) and may feed later layers only through connections
    # listed above, all of whose sources are 'active' units by construction
    return out[:, :D], out[:, D:]


def main():
    for D, seed in ((6, 11), (63, 12)):
        H = 512
        deg, conns, (w1, b1, w2, b2, w3, b3) = build(D, H, seed)
        rng = np.random.default_rng(seed + 100)
        x = rng.integers(-8, 9, (64, D)) / 4.0
        mu, al = evaluate(D, H, conns, (b1, b2, b3), x)
        u = (x - mu) * np.exp(-al)
        ld = -al.sum(1)
        # inverse, coordinate by coordinate, from u alone
        xr = np.zeros_like(x)
        for d in range(D):
            mu_d, al_d = evaluate(D, H, conns, (b1, b2, b3), xr)
            xr[:, d] = u[:, d] * np.exp(al_d[:, d]) + mu_d[:, d]
        assert np.abs(xr - x).max() < 1e-9
        assert np.abs(al).max() < 3 and np.abs(mu).max() < 40
        out = {"x": x, "u_x_order": u, "logdet": ld, "mu": mu, "alpha": al, "D": np.int64(D), "H": np.int64(H),
               "n_connections": np.array([len(conns[1]), len(conns[2]), len(conns[3])]),
               "sd.fc1.weight": w1.astype(np.float32), "sd.fc1.bias": b1.astype(np.float32),
               "sd.fc2.weight": w2.astype(np.float32), "sd.fc2.bias": b2.astype(np.float32),
               "sd.fc3.weight": w3.astype(np.float32), "sd.fc3.bias": b3.astype(np.float32),
               "sd.deg1": deg, "sd.deg2": deg}
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"maf_kat_graph_d{D}.npz"), **out)
        print(f"D={D}: {x.shape[0]} samples, connections {[len(conns[i]) for i in (1, 2, 3)]}, "
              f"|alpha| max {np.abs(al).max():.3f}, |u| max {np.abs(u).max():.2f}")


if __name__ == "__main__":
    main()
