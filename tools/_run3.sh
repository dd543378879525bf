timeout 60 python tools/cnet_diag.py 1000 64 2>&1 | tail -1
timeout 200 python tools/cnet_diag.py 262144 64 2>&1 | tail -2
timeout 200 python tools/cnet_diag.py 65536 256 2>&1 | tail -1
