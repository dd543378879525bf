mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r3_tests.log 2>&1; tail -3 gpurun_out/r3_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3_bench_ts.json 2> gpurun_out/r3_bench_ts.err; python - <<'PY'
import json
for f in ("gpurun_out/r3_bench_ts.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"])
        for k in d.get("kernels", []): print("   ", k.get("kernel","")[:60], round(k.get("frac",0),3), round(k.get("us_per_launch",0),1))
    except Exception as e: print(f, "ERR", e)
PY
NFK_CNET_TS=0 timeout 300 python bench.py --no-cpu-baseline --no-kernel-table > gpurun_out/r3_bench_smem.json 2> gpurun_out/r3_bench_smem.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3_bench_smem.json").read().strip().splitlines()[-1])
print("NFK_CNET_TS=0", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"])
PY
timeout 300 python bench.py --no-cpu-baseline --no-kernel-table --batch 64 --steps 50 > gpurun_out/r3_bench_b64.json 2> gpurun_out/r3_bench_b64.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3_bench_b64.json").read().strip().splitlines()[-1])
print("B=64", d["value"], d["ms_per_step"])
PY
