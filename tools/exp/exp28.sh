python tools/debug_knn.py 32 100000 2 | head -3
python tools/debug_knn.py 24 4000 1 | head -3
python tools/debug_knn.py 17 100000 2 | head -3
python tools/debug_knn.py 64 100000 2 | head -3
python tools/debug_knn.py 50 4000 1 | head -3
for v in "PPP_KNN32_OLD=1" "PPP_X=1" "PPP_KNNF_REGS=1" "PPP_KNNF_REGS=-1" "PPP_KNN_BD=128" "PPP_KNN_BD=128 PPP_KNNF_REGS=-1"; do
  echo "== k32 $v"; env $v PPP_DEBUG=1 python tools/profile_op.py knn32 > /tmp/o.txt 2>&1; grep -E "fast path" /tmp/o.txt | tail -1; grep -E "knn_normals|knn_redo" /tmp/o.txt
done
for v in "PPP_KNN64_OLD=1" "PPP_X=1" "PPP_KNNF_REGS=1" "PPP_KNN_BD=128" "PPP_KNN_BD=128 PPP_KNNF_REGS=1"; do
  echo "== k64 $v"; env $v PPP_DEBUG=1 python tools/profile_op.py knn64 > /tmp/o.txt 2>&1; grep -E "fast path" /tmp/o.txt | tail -1; grep -E "knn_normals|knn_redo" /tmp/o.txt
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_shapes.py -q 2>&1 | tail -4
