timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r3_bench_ts3.json 2> gpurun_out/r3_bench_ts3.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3_bench_ts3.json").read().strip().splitlines()[-1])
print("TS", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["us_per_launch"])
for k in d.get("kernels", []): print("   ", k.get("kernel","")[:60], round(k.get("frac",0),3), round(k.get("us_per_launch",0),1))
PY
timeout 300 python bench.py --no-cpu-baseline --no-kernel-table --batch 64 --steps 50 > gpurun_out/r3_bench_b64b.json 2> gpurun_out/r3_bench_b64b.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3_bench_b64b.json").read().strip().splitlines()[-1])
print("B=64", d["value"], d["ms_per_step"])
PY
