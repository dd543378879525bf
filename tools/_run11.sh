mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r03_bench_default.json 2> gpurun_out/r03_bench_default.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03_bench_default.json").read().strip().splitlines()[-1])
print("default", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["us_per_launch"], "launches", d["gpu_launches_per_step"], "logp", d["logp_delta"], "cpu", d["cpu_baseline"]["value"], d["clocks"])
for k in d.get("kernels", []): print("   ", k.get("kernel","")[:60], round(k.get("frac",0),3), round(k.get("us_per_launch",0),1))
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --no-kernel-table --batch 64 --steps 50 > gpurun_out/r03_bench_b64.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03_bench_b64.json").read().strip().splitlines()[-1])
print("B=64", d["value"], d["ms_per_step"], d["gpu_launches_per_step"])
PY
# profiles of the final tree
timeout 200 python tools/cnet_one.py 2048 > gpurun_out/r3_plain_cnet2.log 2>&1
timeout 300 python bench.py --ncu-step > gpurun_out/r3_plain_step2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
  --log-file gpurun_out/r3_launches_final.csv python bench.py --ncu-step > gpurun_out/r3_ncu_step2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnet_fwd_ts -s 2 -c 2 -f \
  -o gpurun_out/r3_prof_cnet_ts_final python tools/cnet_one.py 2048 > gpurun_out/r3_ncu_cnet2.log 2>&1
ls -la gpurun_out/r3_launches_final.csv gpurun_out/r3_prof_cnet_ts_final.ncu-rep
