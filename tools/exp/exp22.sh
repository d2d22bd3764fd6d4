python tools/debug_knn.py 8 4000 1 | head -3
python tools/debug_knn.py 16 4000 1 | head -3
python tools/debug_knn.py 16 100000 2 | head -3
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
PPP_DEBUG=1 python tools/profile_op.py knn16 > /tmp/o.txt 2>&1; grep -E "fast path" /tmp/o.txt | tail -1; grep -E "knn_normals|knn_redo" /tmp/o.txt
python tools/step_breakdown.py
