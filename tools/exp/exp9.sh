set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/profile_op.py contoursB | head -8
python tools/profile_op.py contoursA | head -5
python bench.py --steps 20 --warmup 5 --no-cfg3 2>&1 | tail -1 > gpurun_out/bench1.json
python tools/shape_perf.py 2>&1 | tail -3
