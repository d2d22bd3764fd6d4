for v in "PPP_X=1" "PPP_KNN16_REGS72=1"; do
  echo "== $v"; env $v PPP_DEBUG=1 python tools/profile_op.py knn16 > /tmp/o.txt 2>&1; grep -E "fast path" /tmp/o.txt | tail -1; grep -E "knn_normals|knn_redo" /tmp/o.txt
done
python tools/debug_knn.py 16 100000 2 | head -3
python tools/debug_knn.py 8 100000 2 | head -3
