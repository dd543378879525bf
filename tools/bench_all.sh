#!/bin/bash
# Every single-GPU bench line of the round (one JSON file per workload under gpurun_out/), each under its own timeout.
run() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r03_bench_$name.json 2> gpurun_out/r03_bench_$name.err || echo "FAILED $name"; tail -c 300 gpurun_out/r03_bench_$name.json | head -c 0; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r03_bench_{n}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"{n:28s} {d['value']:14.1f} {d['unit']}  {d['ms_per_step']:8.3f} ms  launches {d.get('gpu_launches_per_step')}  roofline {r.get('frac')}  logp_delta {(d.get('logp_delta') or {}).get('max_rel_err_per_sample_logp')}  cpu {(d.get('cpu_baseline') or {}).get('value')}")
except Exception as e:
    print(n, "unreadable:", e)
PY
}
run default --steps 20
run b64 --steps 50 --batch 64 --no-cpu-baseline
run b256 --steps 30 --batch 256 --no-cpu-baseline
run b1024 --steps 20 --batch 1024 --no-cpu-baseline
run u8 --steps 20 --u8-input --no-cpu-baseline
run bf16x3_b512 --steps 5 --batch 512 --precision bf16x3 --no-cpu-baseline
run cifar_fwd_inv --steps 10 --workload glow_cifar_fwd_inv_k32
run glow1d --steps 20 --workload glow1d_bsds300_kd_t5_s3
run maf_kd --steps 20 --workload maf_bsds300_kd_t10_s3
run maf_fwd_inv --steps 10 --workload maf_bsds300_fwd_inv_k10
run celeba_kd --steps 10 --workload glow_celeba_kd_t32_s8 --no-cpu-baseline
run celeba_fwd_inv --steps 10 --workload glow_celeba_fwd_inv_k32 --no-cpu-baseline
run celeba_ref_pair --steps 10 --workload glow_celeba_ref_kd_t32h512_s16h256 --no-cpu-baseline
