timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
for op in knn16 knn8; do echo "== $op"; PPP_DEBUG=1 python tools/profile_op.py $op > /tmp/o.txt 2>&1; grep -E "fast path" /tmp/o.txt | tail -1; grep -E "knn_normals|knn_redo" /tmp/o.txt; done
PPP_DEBUG=1 python tools/shape_perf.py 2>&1 | grep -E "fast path|ms/step" | sort | uniq -c | sort -rn | head -30
