#!/usr/bin/env python
"""Benchmark of the KD training hot path (BASELINE.json: "KD-train samples/sec (Glow 32x32x3, ...)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference] [--workload NAME]

One process per GPU (torchrun for N > 1). A step = one KD training step (student fwd, teacher fwd, multi-level
latent MSE, backward, clip 30, Adam) on one synthetic CIFAR-shaped batch per rank. Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (image_shape, L, hidden, teacher K, student K)
    # per-GPU batch 2048 (weak scaling). The reference's config uses 64 (conf/training/cifar.yaml:7) on its single
    # GPU; throughput on synthetic data is quoted at a batch that fills a B200 (--batch overrides; measured on one
    # B200: 39.0 k samples/s at 1024, 42.8 k at 2048, 43.9 k at 4096 — the small upper-level kernels stop being
    # latency-bound).
    "glow_cifar_kd_t32_s8": dict(image=(32, 32, 3), L=3, hidden=512, tK=32, sK=8, batch=2048),
    # BASELINE configs[4]: Glow L=4 K=32 on CelebA-shaped 64x64x3, KD training (level-0 GEMM shape of batch 256 equals
    # CIFAR at batch 1024: 262 144 pixels)
    "glow_celeba_kd_t32_s8": dict(image=(64, 64, 3), L=4, hidden=512, tK=32, sK=8, batch=256),
    # the reference's own CelebA pair (conf/teacher/celeba.yaml, conf/student/celeba.yaml): L=3, teacher K=32 hidden
    # 512, student K=16 hidden 256
    "glow_celeba_ref_kd_t32h512_s16h256": dict(image=(64, 64, 3), L=3, hidden=512, s_hidden=256, tK=32, sK=16,
                                               batch=256),
    # BASELINE configs[2]: Glow L=3 K=32 hidden 512, forward + inverse + log-det (no gradients); metric = samples/s
    # through one x -> z (+ per-sample log-det / bpd) pass followed by one z -> x sampling pass
    "glow_cifar_fwd_inv_k32": dict(image=(32, 32, 3), L=3, hidden=512, tK=32, sK=32, batch=1024, mode="fwd_inv"),
    "glow_celeba_fwd_inv_k32": dict(image=(64, 64, 3), L=4, hidden=512, tK=32, sK=32, batch=256, mode="fwd_inv"),
    # MAF density evaluation + sampling: 10 MADE layers, D = 63, hidden 512 (parity unpinned, see oracle/maf_oracle.py);
    # the z -> x pass is the shared-memory-resident sequential inverse (csrc/maf_inverse.cu), one launch per layer
    "maf_bsds300_fwd_inv_k10": dict(image=(63,), L=1, hidden=512, tK=10, sK=10, batch=65536, is_1d=True, arch="maf",
                                    mode="fwd_inv", data="bsds300"),
    # secondary workloads (BASELINE configs[1]): BSDS300-shaped tabular KD, D = 63, reference batch 65 536
    # (conf/training/tabular.yaml: nll 0.85, kd 0.05, perceptual-L1 0.1 through the inverse pass)
    "glow1d_bsds300_kd_t5_s3": dict(image=(63,), L=1, hidden=32, s_hidden=16, tK=5, sK=3, batch=65536, is_1d=True,
                                    weights=(0.85, 0.05, 0.1), data="bsds300"),
    # MAF teacher 10 MADE layers -> student 3 layers, hidden 512 (no reference implementation exists: parity unpinned)
    "maf_bsds300_kd_t10_s3": dict(image=(63,), L=1, hidden=512, tK=10, sK=3, batch=65536, is_1d=True, arch="maf",
                                  weights=(0.9, 0.1, 0.0), data="bsds300"),
}
METRIC = "kd_train_samples_per_sec"


def synthetic_images(n, shape, seed):
    """floor(U[0,1)*256)/256 - 0.5, the value range produced by the reference's preprocess (data/src/utils.py:7-18)."""
    g = torch.Generator().manual_seed(seed)
    H, W, C = shape
    return torch.floor(torch.rand(n, C, H, W, generator=g) * 256.0) / 256.0 - 0.5


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        out = self.proc.communicate()[0]
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ forward + inverse mode
def run_fwd_inv(args, wl, cfg_desc, warmup):
    """BASELINE configs[2]: one x -> z pass (all layer outputs, per-sample bpd) and one z -> x pass of the K=32 model,
    no gradients, captured in one CUDA graph. value = samples/s through the pair of passes."""
    from nf_distillation_b200 import ops
    from nf_distillation_b200.models import create_glow_model
    from nf_distillation_b200.train import glow_cfg, init_distributed, randomise_zero_params
    import torch.distributed as dist
    rank, world, device = init_distributed()
    B = args.batch or wl["batch"]
    maf = wl.get("arch") == "maf"
    torch.manual_seed(42)
    n_pool = 4
    if maf:
        from nf_distillation_b200.models.maf import create_maf_model
        cfg = dict(image_shape=[wl["image"][0]], hidden_channels=wl["hidden"], K=wl["tK"])
        model = create_maf_model(cfg).to(device).eval()
        host_pool = [torch.randn(B, wl["image"][0], generator=torch.Generator().manual_seed(1000 + rank * 100 + i))
                     .pin_memory() for i in range(n_pool)]
        x = torch.empty(B, wl["image"][0], device=device)
    else:
        H, W, C = wl["image"]
        cfg = glow_cfg(wl["image"], wl["tK"], wl["L"], wl["hidden"])
        model = create_glow_model(cfg)
        randomise_zero_params(model, 43, std=0.01)
        model = model.to(device).eval()
        host_pool = [synthetic_images(B, wl["image"], 1000 + rank * 100 + i).pin_memory() for i in range(n_pool)]
        x = torch.empty(B, C, H, W, device=device)
    dev_pool = [h.to(device) for h in host_pool]
    res = {}

    def passes():
        with torch.no_grad():
            outs, bpd, _ = model(x.clone(), None)
            rev = model(z=outs[-1], temperature=0.0, reverse=True)
            res["bpd"], res["x"] = bpd, rev[-1]
            res["stat"] = torch.stack([bpd.mean(), rev[-1].abs().mean()])

    x.copy_(dev_pool[0])
    passes()
    torch.cuda.synchronize()
    c0 = ops.launch_count()
    passes()
    launches = ops.launch_count() - c0
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        passes()
        with torch.cuda.graph(gr):
            passes()
    torch.cuda.current_stream().wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        x.copy_(dev_pool[i % n_pool]); gr.replay()
    sampler = ClockSampler(device.index or 0)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        x.copy_(dev_pool[i % n_pool]); gr.replay()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    # end to end: pinned host batch -> H2D -> both passes -> D2H of (mean bpd, mean |x|)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        x.copy_(host_pool[i % n_pool], non_blocking=True); gr.replay()
        stat = res["stat"].cpu()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        if maf:
            cb, dt = cpu_maf_fwd_inv(sd, wl, 4096, 1)
            what = "oracle/maf_oracle.py (paper restatement, parity unpinned)"
        else:
            from oracle import glow_oracle as O
            cb = args.cpu_batch
            xc = synthetic_images(cb, wl["image"], 7)
            with torch.no_grad():
                t0 = time.perf_counter()
                outs, _ = O.glow_forward(sd, cfg, xc)
                O.glow_reverse(sd, cfg, outs[-1], 0.0)
                dt = time.perf_counter() - t0
            what = "oracle/glow_oracle.py"
        cpu_base = {"value": cb / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"one forward + one inverse pass of {cb} samples ({what}, torch CPU fp32)"}
    if rank == 0:
        gb = B * world
        cfg_desc.update(per_gpu_batch=B, global_batch=gb, parallelism=f"dp{world}", cuda_graphs=True,
                        passes=("x->z (all layer outputs, nll) + z->x (sequential inverse of every MADE layer)" if maf else
                                "x->z (all outputs, bpd) + z->x (split parts at the prior mean, temperature 0)"),
                        l2="activations (~GBs per pass) exceed the 126 MB L2; inputs rotate over 4 batches")
        cfg_desc.pop("student", None); cfg_desc.pop("loss", None); cfg_desc.pop("optimizer", None)
        print(json.dumps({"metric": "fwd_inv_samples_per_sec", "value": gb / (ms * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                          "data": "synthetic", "config": cfg_desc, "clocks": clocks,
                          "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                                  "h2d_bytes_per_step": host_pool[0].numel() * 4, "d2h_bytes_per_step": 8},
                          "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
                          "last": {"bpd_mean": float(stat[0]), "x_abs_mean": float(stat[1])},
                          "roofline": None, "cpu_baseline": cpu_base}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_maf_fwd_inv(sd, wl, cb, n):
    """Seconds for one x -> z + one z -> x pass of `cb` samples through the paper restatement (oracle/maf_oracle.py)."""
    from oracle import maf_oracle as MO
    D = wl["image"][0]
    xc = torch.randn(cb, D, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(n):
            outs, _ = MO.maf_forward(sd, D, wl["tK"], xc)
            MO.maf_inverse(sd, D, wl["tK"], outs[-1])
        return cb, (time.perf_counter() - t0) / n


def cpu_kd_step_fn(wl, batch, seed=42):
    """The reference's CPU algorithm for this path, restated in oracle/ (the reference itself is Python and does not
    travel to the GPU box): KD train step = fwd student+teacher, loss, backward, clip 30, Adam."""
    from oracle import glow_oracle as O
    from nf_distillation_b200.models import create_glow_model
    from nf_distillation_b200.train import glow_cfg, randomise_zero_params
    is_1d = wl.get("is_1d", False)
    torch.manual_seed(seed)
    if wl.get("arch") == "maf":
        from oracle import maf_oracle as MO
        from nf_distillation_b200.models.maf import create_maf_model
        D = wl["image"][0]
        t_model = create_maf_model(dict(image_shape=[D], hidden_channels=wl["hidden"], K=wl["tK"]))
        s_model = create_maf_model(dict(image_shape=[D], hidden_channels=wl["hidden"], K=wl["sK"]))
        pn = dict(s_model.named_parameters())
        s_sd = {k: v.clone().requires_grad_(k in pn) for k, v in s_model.state_dict().items()}
        t_sd = {k: v.clone() for k, v in t_model.state_dict().items()}
        params = [v for v in s_sd.values() if v.requires_grad]
        opt = torch.optim.Adam(params, lr=5e-4)
        x = torch.randn(batch, D)
        wn, wk, _ = wl["weights"]

        def step():
            opt.zero_grad(set_to_none=True)
            s_z, s_nll = MO.maf_forward(s_sd, D, wl["sK"], x)
            with torch.no_grad():
                t_z, _ = MO.maf_forward(t_sd, D, wl["tK"], x)
            kd = O.kd_loss(s_z, t_z, [1, 2], [3, 7])
            loss = (wn * s_nll + wk * kd).mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 30.0)
            opt.step()
            return float(loss.detach())
        return step
    if is_1d:
        s_cfg = glow_cfg(wl["image"], wl["sK"], wl["L"], wl.get("s_hidden", wl["hidden"]), is_1d=True, y_classes=0)
        t_cfg = glow_cfg(wl["image"], wl["tK"], wl["L"], wl["hidden"], is_1d=True, y_classes=0)
    else:
        s_cfg = glow_cfg(wl["image"], wl["sK"], wl["L"], wl.get("s_hidden", wl["hidden"]))
        t_cfg = glow_cfg(wl["image"], wl["tK"], wl["L"], wl["hidden"])
    t_model, s_model = create_glow_model(t_cfg), create_glow_model(s_cfg)   # parameter containers only (CPU)
    randomise_zero_params(s_model, seed + 1)
    randomise_zero_params(t_model, seed + 2)
    s_sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k in dict(s_model.named_parameters()))
            for k, v in s_model.state_dict().items()}
    t_sd = {k: v.clone() for k, v in t_model.state_dict().items()}
    params = [v for v in s_sd.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=5e-4)
    if is_1d:
        x = torch.randn(batch, wl["image"][0])
        wn, wk, wp = wl["weights"]
        weights = {"nll": wn, "kd": wk, "perceptual": wp}
    else:
        x = synthetic_images(batch, wl["image"], seed)
        weights = {"nll": 0.9, "kd": 0.1, "perceptual": 0.0}

    def step():
        n1 = n2 = latent = None
        if is_1d:
            latent = torch.randn_like(x)
        else:
            n1, n2 = torch.rand_like(x) / 256, torch.rand_like(x) / 256
        opt.zero_grad(set_to_none=True)
        out = O.kd_step(s_sd, s_cfg, t_sd, t_cfg, x, weights, n1, n2, latent)
        out["result_loss"].backward()
        torch.nn.utils.clip_grad_norm_(params, 30.0)
        opt.step()
        return float(out["result_loss"].detach())
    return step


def run_cpu(wl, batch, steps, warmup):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_kd_step_fn(wl, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, cores


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: workload's)")
    ap.add_argument("--workload", default="glow_cifar_kd_t32_s8", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: warm up, then run ONE eager step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench value")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "native" else max(args.warmup, 1)
    arch = wl.get("arch", "glow")
    cfg_desc = {"workload": args.workload, "teacher": f"{arch} K={wl['tK']} L={wl['L']} hidden={wl['hidden']}",
                "student": f"{arch} K={wl['sK']} L={wl['L']} hidden={wl.get('s_hidden', wl['hidden'])}",
                "image": list(wl["image"]),
                "loss": "0.9 nll + 0.1 kd(mse, 4 levels)" if "weights" not in wl else
                        "%.2f nll + %.2f kd(mse) + %.2f perceptual(l1, inverse pass)" % wl["weights"],
                "optimizer": "adam 5e-4, clip 30"}

    if args.impl == "reference":
        if rank != 0:
            return
        cb = args.cpu_batch if not wl.get("is_1d", False) else 65536
        metric = METRIC
        if wl.get("mode") == "fwd_inv" and wl.get("arch") == "maf":
            from nf_distillation_b200.models.maf import create_maf_model
            metric = "fwd_inv_samples_per_sec"
            torch.manual_seed(42)
            model = create_maf_model(dict(image_shape=[wl["image"][0]], hidden_channels=wl["hidden"], K=wl["tK"]))
            sd = {k: v.detach() for k, v in model.state_dict().items()}
            n = max(1, min(args.steps, 3))
            cpu_maf_fwd_inv(sd, wl, 256, 1)
            cb, dt = cpu_maf_fwd_inv(sd, wl, 4096, n)
            ms = dt * 1e3
            value, cores = cb / dt, torch.get_num_threads()
        elif wl.get("mode") == "fwd_inv":   # forward + inverse + log-det of the K=32 model through the oracle port
            from oracle import glow_oracle as O
            from nf_distillation_b200.models import create_glow_model
            from nf_distillation_b200.train import glow_cfg, randomise_zero_params
            metric = "fwd_inv_samples_per_sec"
            torch.manual_seed(42)
            cfg = glow_cfg(wl["image"], wl["tK"], wl["L"], wl["hidden"])
            model = create_glow_model(cfg)
            randomise_zero_params(model, 43, std=0.01)
            sd = {k: v.detach() for k, v in model.state_dict().items()}
            xc = synthetic_images(cb, wl["image"], 7)
            n = max(1, min(args.steps, 3))
            with torch.no_grad():
                for it in range(n + 1):
                    if it == 1:
                        t0 = time.perf_counter()
                    outs, _ = O.glow_forward(sd, cfg, xc)
                    O.glow_reverse(sd, cfg, outs[-1], 0.0)
            ms = (time.perf_counter() - t0) * 1e3 / n
            value, cores = cb / (ms * 1e-3), torch.get_num_threads()
        else:
            value, ms, cores = run_cpu(wl, cb, max(1, min(args.steps, 3)), 1)
        cfg_desc.update(per_gpu_batch=cb, global_batch=cb, parallelism="cpu")
        if wl.get("mode") == "fwd_inv":
            for k in ("student", "loss", "optimizer"):
                cfg_desc.pop(k, None)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg_desc,
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{cb}-sample steps of the same workload (oracle/ port of the reference, "
                                       f"torch CPU fp32, {cores} threads)"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if wl.get("mode") == "fwd_inv":
        return run_fwd_inv(args, wl, cfg_desc, warmup)
    import torch.distributed as dist
    from nf_distillation_b200 import ops
    from nf_distillation_b200.train import KDTrainer, glow_cfg, init_distributed, kd_config
    rank, world, device = init_distributed()
    B = args.batch or wl["batch"]
    is_1d = wl.get("is_1d", False)
    if is_1d:
        D = wl["image"][0]
        H = W = 2
        C = D
        s_cfg = glow_cfg(wl["image"], wl["sK"], wl["L"], wl.get("s_hidden", wl["hidden"]), is_1d=True, y_classes=0)
        t_cfg = glow_cfg(wl["image"], wl["tK"], wl["L"], wl["hidden"], is_1d=True, y_classes=0)
        if wl.get("arch") == "maf":
            s_cfg["architecture"] = t_cfg["architecture"] = "maf"
        wn, wk, wp = wl["weights"]
        config = kd_config(s_cfg, t_cfg, data=wl["data"], nll=wn, kd=wk, perceptual=wp)
        shape = (B, D)
    else:
        H, W, C = wl["image"]
        config = kd_config(glow_cfg(wl["image"], wl["sK"], wl["L"], wl.get("s_hidden", wl["hidden"])),
                           glow_cfg(wl["image"], wl["tK"], wl["L"], wl["hidden"]))
        shape = (B, C, H, W)
    trainer = KDTrainer(config, shape, device, use_graphs=not args.no_graphs)
    n_pool = 4

    def synth(i):
        if is_1d:   # z-scored tabular features (data/src/power.py:42-52): N(0, 1)
            g = torch.Generator().manual_seed(1000 + rank * 100 + i)
            return torch.randn(B, wl["image"][0], generator=g)
        return synthetic_images(B, wl["image"], 1000 + rank * 100 + i)
    host_pool = [synth(i).pin_memory() for i in range(n_pool)]
    dev_pool = [h.to(device) for h in host_pool]
    trainer.x.copy_(dev_pool[0])
    if args.ncu_step:
        trainer.use_graphs = False
        trainer.warmup(iters=3)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        trainer.step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_step": True, "batch": B, "losses": trainer.losses.cpu().tolist()}))
        return
    trainer.warmup(iters=1)                      # eager step(s) + graph capture
    launches_eager_step = None
    # count kernels of ONE step: run one extra eager step outside the graphs
    c0 = ops.launch_count()
    trainer._forward_backward()
    launches_per_step = ops.launch_count() - c0
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for i in range(warmup):
        trainer.x.copy_(dev_pool[i % n_pool])
        trainer.step_device()
    sampler = ClockSampler(device.index or 0)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.x.copy_(dev_pool[i % n_pool])
        trainer.step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    losses = trainer.losses.cpu().tolist()

    # ---- end-to-end timing through the public API: pinned host batch -> H2D -> step -> D2H of the loss scalars
    for i in range(2):
        trainer.step(host_pool[i % n_pool])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = trainer.step(host_pool[i % n_pool])
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- roofline of the dominant kernel (profiles/r01_launches_step_b1024.txt: cnet_fwd_fused_kernel, ~30 % of the
    #      step): fused conv#1+conv#2 of the coupling net at the level-0 shape, M = B*H/2*W/2 pixels, timed alone with
    #      CUDA events on its launch stream, 10 launches per CUDA graph so host launch overhead is not measured, and
    #      rotating buffers larger than L2. Algorithmic work per launch = 2*M*512*(K1p+512) FLOP.
    roof = None
    if rank == 0 and not is_1d:
        pk = peaks()
        M, hid, K1p = B * (H // 2) * (W // 2), wl["hidden"], 64
        nbuf = 3
        cols = [(torch.randn(M, K1p, device=device) * 0.5).bfloat16() for _ in range(nbuf)]
        outs_ = [torch.empty(M, hid, device=device, dtype=torch.bfloat16) for _ in range(nbuf)]
        W1 = (torch.randn(hid, K1p, device=device) * 0.1).bfloat16()
        W2 = (torch.randn(hid, hid, device=device) * 0.05).bfloat16()
        b1, b2 = torch.zeros(hid, device=device), torch.zeros(hid, device=device)
        fused = ops.cnet_fused_supported(hid, K1p)

        def run(i):
            if fused:
                ops.cnet_fwd_fused(cols[i], K1p, W1, W2, b1, b2, outs_[i], M, hid)
            else:
                ops.gemm_nt(outs_[(i + 1) % nbuf], W2, M, hid, hid, ops.EPI_BIAS_RELU_BF16, outs_[i], bias=b2)
        for i in range(3):
            run(i % nbuf)
        torch.cuda.synchronize()
        reps = 12
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(reps):
                run(i % nbuf)
        gr.replay()
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        gr.replay()
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / reps
        flops = 2.0 * M * hid * ((K1p if fused else 0) + hid)
        ach = flops / (kms * 1e-3) / 1e12
        alg_bytes = 2.0 * (M * K1p + M * hid + hid * K1p + hid * hid)
        roof = {"bound": "tensor",
                "kernel": "cnet_fwd_fused_kernel (conv#1+conv#2 of the coupling net, level 0)" if fused
                else "gemm_nt_pair_kernel<BIAS_RELU_BF16> (conv#2)",
                "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                # dram__bytes_read+write per launch from profiles/r01_prof_cnet_r1.txt (ncu --set full, B=1024)
                "traffic": 242.1e6 * (B / 1024.0) if fused else None,
                "peak_src": pk["src"] + " burst (kernel timed alone)", "us_per_launch": kms * 1e3,
                "flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes}
        del cols, outs_

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch if not is_1d else (65536 if wl.get("arch") != "maf" else 8192)
        v, cms, cores = run_cpu(wl, cb, 2, 1)
        cpu_base = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                    "sample": f"2 KD train steps of {cb} samples, same model/config "
                              f"(oracle/ on torch CPU fp32, {cores} threads), {cms:.0f} ms/step"}

    if rank == 0:
        gb = B * world
        cfg_desc.update(per_gpu_batch=B, global_batch=gb, parallelism=f"dp{world}",
                        cuda_graphs=not args.no_graphs,
                        l2="per-step working set (activations ~GBs) exceeds the 126 MB L2; inputs rotate over 4 batches")
        out = {"metric": METRIC, "value": gb / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
               "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
               "data": "synthetic", "config": cfg_desc, "clocks": clocks,
               "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": host_pool[0].numel() * 4, "d2h_bytes_per_step": 16},
               "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
               "losses_last_step": dict(zip(("nll", "kd", "perceptual", "loss"), losses)),
               "roofline": roof, "cpu_baseline": cpu_base}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
