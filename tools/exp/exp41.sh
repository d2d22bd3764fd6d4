timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/profile_op.py knn16 | grep -E "knn_normals|knn_redo"
python tools/step_jitter.py 200
python tools/step_timeline.py | tail -18
