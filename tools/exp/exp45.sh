python tools/shape_perf.py 2>&1 | tail -3
