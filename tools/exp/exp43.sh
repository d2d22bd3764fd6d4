timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/profile_op.py knn16 | grep -E "scan_chained|cell_"
python tools/profile_op.py knn16 20000000 | grep -E "scan_chained|cell_"
python tools/step_jitter.py 200
