timeout 200 python tools/zbwd_one.py 2048 > gpurun_out/r3_plain_zbwd.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"coupling_bwd|affine1x1_bwd" -s 2 -c 2 -f \
  -o gpurun_out/r3_prof_zbwd python tools/zbwd_one.py 2048 > gpurun_out/r3_ncu_zbwd.log 2>&1
ls -la gpurun_out/r3_prof_zbwd*
