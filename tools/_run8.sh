mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for b in 64 256; do
timeout 300 python bench.py --no-cpu-baseline --no-kernel-table --batch $b --steps 50 > gpurun_out/r3_ws_b$b.json 2> gpurun_out/r3_ws_b$b.err; python - $b <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r3_ws_b{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("B=%s" % sys.argv[1], d["value"], d["ms_per_step"], d["gpu_launches_per_step"])
PY
NFK_WGRAD_STREAM_MAX_M=0 timeout 300 python bench.py --no-cpu-baseline --no-kernel-table --batch $b --steps 50 > gpurun_out/r3_nows_b$b.json 2> gpurun_out/r3_nows_b$b.err; python - $b <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r3_nows_b{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("B=%s single stream" % sys.argv[1], d["value"], d["ms_per_step"])
PY
done
timeout 300 python bench.py --no-cpu-baseline --no-kernel-table > gpurun_out/r3_ws_default.json 2> gpurun_out/r3_ws_default.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3_ws_default.json").read().strip().splitlines()[-1])
print("B=2048", d["value"], d["ms_per_step"])
PY
