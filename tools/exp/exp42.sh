timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/profile_op.py contoursB 1000000 200 | grep -E "pair_nodes|cell_scatter|slice_order"
python tools/step_jitter.py 200
