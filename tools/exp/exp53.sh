timeout 600 python -m pytest tests/test_gpu_edge.py -q -k "exact_ties" 2>&1 | tail -6
