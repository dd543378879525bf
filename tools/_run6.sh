mkdir -p gpurun_out
bash tools/bench_all.sh 2>&1 | tee gpurun_out/r03_bench_all.log
