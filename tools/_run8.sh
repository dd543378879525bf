for m in 65536 200000 100000000; do
NFK_WGRAD_STREAM_MAX_M=$m timeout 300 python bench.py --no-cpu-baseline --no-kernel-table > gpurun_out/r3_ws_m$m.json 2> gpurun_out/r3_ws_m$m.err; python - $m <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/r3_ws_m{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("MAX_M=%s B=2048" % sys.argv[1], d["value"], d["ms_per_step"])
PY
done
