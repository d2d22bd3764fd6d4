set -x
timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_host_adapter.py tests/test_gpu_edge.py -x -q 2>&1 | tail -15
