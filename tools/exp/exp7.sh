set -x
python tools/pcie_probe3.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_shapes.py -x -q 2>&1 | tail -15
python tools/shape_perf.py 2>&1 | tail -4
