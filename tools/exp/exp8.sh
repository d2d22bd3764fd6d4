set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_shapes.py -x -q 2>&1 | tail -8
for op in knn32 knn24; do python tools/profile_op.py $op | head -3; done
PPP_KNN32_OLD=1 python tools/profile_op.py knn32 | head -3
python tools/profile_op.py knn32 10000000 | head -3
