set -x
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
python tools/shape_perf.py 2>&1 | tail -8
PPP_DEBUG=1 python tools/shape_perf.py 20000 2>&1 | grep -E "ingest|chosen" | sort | uniq -c
python bench.py --steps 20 --warmup 5 --no-cfg3 2>&1 | tail -1 > gpurun_out/bench1.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
