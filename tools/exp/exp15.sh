set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/profile_op.py knn16 | head -3
python tools/profile_op.py knn32 | head -3
python tools/profile_op.py radius | head -4
python tools/shape_perf.py 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
