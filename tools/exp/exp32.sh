export PPP_KNN16_TILE=1
python tools/debug_knn.py 16 100000 2 | head -3
python tools/debug_knn.py 8 4000 1 | head -3
python tools/debug_knn.py 16 4000 1 | head -3
unset PPP_KNN16_TILE
for v in "PPP_PDL=0" "PPP_X=1" "PPP_KNN16_TILE=1" "PPP_KNN16_TILE=1 PPP_KNN16_REGS72=1" "PPP_KNN16_TILE=1 PPP_KNN_BD=64" "PPP_KNN16_TILE=1 PPP_KNN_BD=96"; do
  echo "== $v"; env $v PPP_DEBUG=1 python tools/profile_op.py knn16 > /tmp/o.txt 2>&1; grep -E "fast path" /tmp/o.txt | tail -1; grep -E "knn_normals|knn_redo" /tmp/o.txt
done
PPP_KNN16_TILE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_shapes.py -q 2>&1 | tail -4
echo "== jitter PDL on"; python tools/step_jitter.py 200
echo "== jitter PDL off"; PPP_PDL=0 python tools/step_jitter.py 200
echo "== jitter PDL on, TILE"; PPP_KNN16_TILE=1 python tools/step_jitter.py 200
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
