timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/levels_bench.py 2048 2>&1 | tail -5
timeout 200 python tools/levels_bench.py 64 2>&1 | tail -5
