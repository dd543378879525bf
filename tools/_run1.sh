mkdir -p gpurun_out
timeout 120 python tools/cnet_diag.py 256 64 > gpurun_out/ts_a.log 2>&1; echo "rc=$?" >> gpurun_out/ts_a.log
timeout 200 python tools/cnet_diag.py > gpurun_out/ts_diag.log 2>&1; echo "rc=$?" >> gpurun_out/ts_diag.log
NFK_CNET_TS=0 timeout 200 python tools/cnet_diag.py > gpurun_out/ts_diag_old.log 2>&1
timeout 300 python -m pytest tests/test_headline_parity_gpu.py -x -q -m gpu -k "cnet" > gpurun_out/ts_tests.log 2>&1
tail -3 gpurun_out/ts_a.log; cat gpurun_out/ts_diag.log; cat gpurun_out/ts_diag_old.log | tail -8; tail -5 gpurun_out/ts_tests.log
