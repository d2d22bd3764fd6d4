timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
PPP_GPU_LIB=$PWD/polishpathplanning_b200/libppp_gpu_check.so python tools/memcheck_small.py 2>&1 | tail -12
