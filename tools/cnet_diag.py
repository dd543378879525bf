"""Fused conv#1+conv#2 vs the two separate GEMMs: bitwise comparison (h2, and h1 + masks in training mode) + timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import ops
from nf_distillation_b200._lib import LIB
dev = "cuda"
torch.manual_seed(0)
CASES = ((256, 64), (1000, 64), (65536, 64), (16384, 128), (4096, 256), (262144, 64))
if len(sys.argv) > 1:   # e.g. `cnet_diag.py 262144 64` for an ncu capture of one shape
    CASES = ((int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 64),)
for (M, K1p) in CASES:
    hid = 512
    col = (torch.randn(M, K1p, device=dev) * 0.5).bfloat16()
    B1 = (torch.randn(hid, K1p, device=dev) * 0.1).bfloat16()
    B2 = (torch.randn(hid, hid, device=dev) * 0.05).bfloat16()
    b1, b2 = torch.randn(hid, device=dev) * 0.1, torch.randn(hid, device=dev) * 0.1
    h1a = torch.empty(M, hid, device=dev, dtype=torch.bfloat16); h2a = torch.empty_like(h1a)
    h1b = torch.full_like(h1a, float("nan")); h2b = torch.full_like(h1a, float("nan")); h2c = torch.full_like(h1a, float("nan"))
    m1a, m2a = ops.relu_mask_like(M, hid, dev), ops.relu_mask_like(M, hid, dev)
    m1b, m2b = torch.zeros_like(m1a), torch.zeros_like(m2a)
    def unfused():
        ops.gemm_nt(col, B1, M, hid, K1p, ops.EPI_BIAS_RELU_BF16, h1a, bias=b1, aux=m1a)
        ops.gemm_nt(h1a, B2, M, hid, hid, ops.EPI_BIAS_RELU_BF16, h2a, bias=b2, aux=m2a)
    def fused_train():
        ops.cnet_fwd_fused(col, K1p, B1, B2, b1, b2, h2b, M, hid, h1=h1b, mask1=m1b, mask2=m2b)
    def fused():
        ops.cnet_fwd_fused(col, K1p, B1, B2, b1, b2, h2c, M, hid)
    unfused(); fused_train(); fused(); torch.cuda.synchronize()
    ok = (torch.equal(h2a, h2b), torch.equal(h2a, h2c), torch.equal(h1a, h1b), torch.equal(m1a, m1b), torch.equal(m2a, m2b))
    def t(fn):
        for _ in range(2): fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10 * 1e3
    tu, tf, tt = t(unfused), t(fused), t(fused_train)
    fl = 2.0 * M * hid * (K1p + hid)
    print(f"M={M} K1p={K1p}: equal(h2 train, h2 infer, h1, m1, m2)={ok} | unfused {tu:.1f} us, fused {tf:.1f} us ({fl/tf/1e6:.0f} TFLOP/s), fused+save {tt:.1f} us")
    if M == 262144:
        prof = torch.zeros(168, 8, device=dev, dtype=torch.int64)
        LIB.nfk_cnet_set_prof(prof.data_ptr()); fused(); torch.cuda.synchronize(); LIB.nfk_cnet_set_prof(None)
        p = prof.cpu().double()
        tl = p[148:151].clone(); tw = p[152:168].clone(); p = p[:148]
        pe = p[1::2]; pe = pe[pe[:, 0] > 0]
        p = p[0::2]; p = p[p[:, 0] > 0]
        names = ["total", "wait operands", "wait acc-free", "wait h1", "of operands: A tile", "weights GEMM 1", "weights GEMM 2 first quarter"]
        print("   MMA issuer cycles per CTA:", {n: int(p[:, i].mean()) for i, n in enumerate(names)}, "tiles/CTA %.1f" % (1024 / 74))
        if len(pe):
            en = ["total", "wait acc GEMM1", "wait acc GEMM2", "wait h1-free", "wait staging", "tcgen05.ld GEMM2", "stage+fence+store GEMM2"]
            print("   epilogue warp 2 (odd CTA):", {n: int(pe[:, i].mean()) for i, n in enumerate(en)})
        if tl.abs().sum() > 0:   # timeline of one tile (cycles relative to the MMA issuer reaching the tile)
            t0 = tl[2, 0]
            f = lambda r: [int(v - t0) if v > 0 else None for v in r]
            print("   timeline, MMA issuer [tile start, G1 q0 acc ready, G1 issued, G2 q0 start, q1, q2, q3, G2 issued]:", f(tl[2]))
            for wsi in (0, 1):
                print(f"   timeline, epilogue set {wsi} [G1 a: acc full, ld done, math done, h1 free, st done, handed | G1 b: acc full, handed]:", f(tl[wsi]))
            if tw.abs().sum() > 0:
                print("   per epilogue warp (ew: set = ew>>3, hf = (ew>>2)&1, quadrant = (ew+2)&3): G1 a [acc full, ld done, math done, handed], G1 b [acc full, handed]")
                for ew in range(16):
                    r = tw[ew]
                    print(f"     ew {ew:2d}:", [int(r[i] - t0) if r[i] > 0 else None for i in (0, 1, 2, 5, 6, 7)])
