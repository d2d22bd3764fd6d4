set -x
timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py -x -q 2>&1 | tail -8
PPP_PROJECTIONS=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_paths.py -x -q 2>&1 | tail -8
PPP_DEBUG=1 python tools/shape_perf.py 2>&1 | grep -v "knn fast path" | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cfg3 2>&1 | tail -1 > gpurun_out/bench1.json
python tools/profile_op.py contoursB 1000000 5000 | head -12
