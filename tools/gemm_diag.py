"""GPU diagnostic for the tcgen05 GEMM tiles: compares against torch fp32 matmul and prints error structure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nf_distillation_b200 import _lib

L = _lib.LIB
dev = "cuda"
torch.manual_seed(0)


def report(name, got, ref, tol):
    err = (got.float() - ref).abs()
    scale = ref.abs().max().item() + 1e-6
    rel = err.max().item() / scale
    ok = rel < tol
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max_abs_err={err.max().item():.4e} rel={rel:.3e} ref_max={scale:.3e}")
    if not ok:
        M, N = ref.shape
        # error by 128-row block and 16-col block
        eb = err[: (M // 32) * 32, : (N // 8) * 8].reshape(M // 32, 32, N // 8, 8).amax(dim=(1, 3))
        print("   err by (32-row, 8-col) block, first 8x16:\n", (eb[:8, :16] / scale).cpu().numpy().round(3))
        print("   got[0,:8]", got[0, :8].float().cpu().numpy(), "\n   ref[0,:8]", ref[0, :8].cpu().numpy())
    return ok


def nt(M, N, K, epi):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    ref = A.float() @ B.float().t()
    st = torch.cuda.current_stream().cuda_stream
    if epi == 0:
        out = torch.full((M, N), float("nan"), device=dev)
        bias = torch.randn(N, device=dev)
        rc = L.nfk_gemm_nt_bf16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, 0, out.data_ptr(), N, bias.data_ptr(), None, 0, None, st)
        torch.cuda.synchronize()
        assert rc == 0, rc
        return report(f"nt f32 M{M} N{N} K{K}", out, ref + bias, 2e-3)
    if epi == 1:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        bias = torch.randn(N, device=dev)
        mask = torch.zeros((N + 31) // 32, M, device=dev, dtype=torch.int32)
        rc = L.nfk_gemm_nt_bf16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, 1, out.data_ptr(), N, bias.data_ptr(), mask.data_ptr(), mask.stride(0), None, st)
        torch.cuda.synchronize()
        assert rc == 0, rc
        ok = report(f"nt relu M{M} N{N} K{K}", out, torch.relu(ref + bias), 1e-2)
        bits = ((mask.t()[:, :, None] >> torch.arange(32, device=dev, dtype=torch.int32)) & 1).reshape(M, -1)[:, :N].bool()
        same = (bits == (out > 0)).all().item()
        print(f"   relu bitmask consistent with output: {same}")
        return ok and same
    if epi == 2:
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        aux = torch.relu(torch.randn(M, N, device=dev)).bfloat16()
        w = (aux > 0).reshape(M, N // 32, 32).to(torch.int64) << torch.arange(32, device=dev, dtype=torch.int64)
        mask = w.sum(-1)
        mask = torch.where(mask >= 2 ** 31, mask - 2 ** 32, mask).to(torch.int32).t().contiguous()
        cs = torch.zeros(N, device=dev)
        rc = L.nfk_gemm_nt_bf16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, 2, out.data_ptr(), N, None, mask.data_ptr(), mask.stride(0), cs.data_ptr(), st)
        torch.cuda.synchronize()
        assert rc == 0, rc
        refm = ref * (aux > 0)
        a = report(f"nt mask M{M} N{N} K{K}", out, refm, 1e-2)
        b = report(f"   colsum", cs[None], refm.sum(0)[None], 5e-3)   # sums of the bf16-rounded outputs
        return a and b


def tn(Kpix, Mo, No):
    A = (torch.randn(Kpix, Mo, device=dev) * 0.5).bfloat16()
    B = (torch.randn(Kpix, No, device=dev) * 0.5).bfloat16()
    ref = A.float().t() @ B.float()
    out = torch.zeros(Mo, No, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = L.nfk_gemm_tn_bf16(A.data_ptr(), Mo, B.data_ptr(), No, Mo, No, Kpix, out.data_ptr(), No, 148, st)
    torch.cuda.synchronize()
    assert rc == 0, rc
    return report(f"tn Kpix{Kpix} Mo{Mo} No{No}", out, ref, 2e-3)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "libnfk", L.nfk_version())
    ok = True
    ok &= nt(128, 16, 64, 0)
    ok &= nt(128, 64, 128, 0)
    ok &= nt(256, 256, 512, 0)
    ok &= nt(1000, 112, 512, 0)
    ok &= nt(16384, 512, 512, 1)
    ok &= nt(16384, 512, 64, 1)
    ok &= nt(4096, 512, 128, 2)
    ok &= nt(65536, 512, 128, 2)
    ok &= nt(70000, 128, 512, 0)
    ok &= nt(65536 + 77, 512, 512, 1)
    ok &= nt(16384, 448, 512, 0)
    ok &= tn(64, 128, 64)
    ok &= tn(4096, 512, 64)
    ok &= tn(16384, 512, 512)
    ok &= tn(16384 + 40, 512, 128)
    ok &= tn(5000, 448, 512)
    # timing of the conv2-shaped GEMM
    M, N, K = 65536, 512, 512
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16); bias = torch.zeros(N, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        L.nfk_gemm_nt_bf16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, 1, out.data_ptr(), N, bias.data_ptr(), None, 0, None, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.nfk_gemm_nt_bf16(A.data_ptr(), K, B.data_ptr(), K, M, N, K, 1, out.data_ptr(), N, bias.data_ptr(), None, 0, None, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"nt 65536x512x512 relu-bf16: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
    e0.record()
    for _ in range(20):
        torch.relu(A @ B.t())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"torch bf16 matmul+relu same shape: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
    print("ALL OK" if ok else "SOME BAD")
