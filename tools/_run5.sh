timeout 600 ncu --set full --clock-control none --import-source on -k regex:"affine1x1_fwd|pconv_coupling" -s 2 -c 2 -f \
  -o gpurun_out/r3_prof_levels24 python tools/levels_one.py 2048 > gpurun_out/r3_ncu_levels24.log 2>&1
ls -la gpurun_out/r3_prof_levels24*
