mkdir -p gpurun_out
# 1) plain runs first (a number printed under ncu is never a bench value)
timeout 200 python tools/cnet_one.py 2048 > gpurun_out/r3_plain_cnet.log 2>&1 || exit 1
timeout 300 python bench.py --ncu-step > gpurun_out/r3_plain_step.log 2>&1 || exit 1
# 2) launch list of one step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
  --log-file gpurun_out/r3_launches_step.csv python bench.py --ncu-step > gpurun_out/r3_ncu_step.log 2>&1
# 3) full capture of the dominant kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnet_fwd_ts -s 2 -c 2 -f \
  -o gpurun_out/r3_prof_cnet_ts python tools/cnet_one.py 2048 > gpurun_out/r3_ncu_cnet.log 2>&1
ls -la gpurun_out/r3_*
