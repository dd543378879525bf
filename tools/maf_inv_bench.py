"""Time the MADE inverse: resident one-launch kernel (16 / 32 samples per warp) against the D-pass GEMM inverse.
Run on the GPU box: python tools/maf_inv_bench.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nf_distillation_b200.models.maf import MADE  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    for D, H in [(63, 512), (6, 512)]:
        torch.manual_seed(0)
        made = MADE(D, H, flip=True).cuda()
        u = torch.randn(B, D, device="cuda")
        ld = torch.zeros(B, device="cuda")
        with torch.no_grad():
            res = {}
            made.push_inverse = True
            if made._inverse_jobs(u.device)[1]:
                res["resident_push"] = timeit(lambda: made(u, logdet=ld, reverse=True))
            made.push_inverse = False
            for mt in (1, 2):
                made.resident_mtiles = mt
                res[f"resident_mt{mt}"] = timeit(lambda: made(u, logdet=ld, reverse=True))
            made.push_inverse = True
            made.resident_inverse = False
            res["dpass"] = timeit(lambda: made(u, logdet=ld, reverse=True), n=2)
            made.resident_inverse = True
            fwd = timeit(lambda: made(u, logdet=ld))
        macs = sum(int(m.sum()) for m in __import__("oracle.maf_oracle", fromlist=["masks"]).masks(
            D, made.deg1.cpu().long(), made.deg2.cpu().long()))
        print(f"D={D} H={H} B={B}: forward {fwd:.0f} us; " + "; ".join(f"{k} {v:.0f} us" for k, v in res.items())
              + f"; masked MACs/sample {macs}; resident best = {2 * macs * B / min(v for k, v in res.items() if k != 'dpass') / 1e6:.1f} TFLOP/s",
              flush=True)
