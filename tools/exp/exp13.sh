set -x
python tools/shape_perf.py 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --no-cfg3 2>&1 | tail -1 > gpurun_out/bench1.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
