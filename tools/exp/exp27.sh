for op in knn16 knn32 knn64 radius knn8; do
  echo "== $op"; python tools/profile_op.py $op > /tmp/o.txt 2>&1; grep -E "knn_normals|knn_redo|radius" /tmp/o.txt
done
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/step_jitter.py 200
