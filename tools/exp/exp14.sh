set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_shapes.py -x -q 2>&1 | tail -5
for F in 1.35 1.3 1.25 1.2 1.15; do echo "== F=$F"; PPP_DEBUG=1 PPP_CELL_FACTOR=$F python tools/profile_op.py knn16 2>&1 | grep -E "knn_normals|knn_retry|knn_redo|fast path" | sort | uniq | head -5; done
echo "== one level"; PPP_KNN_ONE_LEVEL=1 python tools/profile_op.py knn16 | head -3
