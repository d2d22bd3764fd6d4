set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q 2>&1 | tail -3
for BD in 128 96 64; do echo "== BD=$BD"; PPP_KNN_BD=$BD python tools/profile_op.py knn16 | head -3; done
for cfg in "1 1.35" "1 1.2" "3 1.35" "3 1.25" "2 1.25" "2 1.45"; do set -- $cfg; echo "== R0=$1 F=$2"; PPP_KNN_R0=$1 PPP_CELL_FACTOR=$2 python tools/profile_op.py knn16 | head -3; done
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-600
