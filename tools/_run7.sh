mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r03_bench_2gpu.json 2> gpurun_out/r03_bench_2gpu.err; tail -c 600 gpurun_out/r03_bench_2gpu.json | head -c 600; echo
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r03_bench_2gpu.json").read().strip().splitlines()[-1])
print("2 GPUs:", d["value"], d["ms_per_step"], d["n_gpus"])
PY
