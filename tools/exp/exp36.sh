timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_host_adapter.py -q 2>&1 | tail -5
for v in "PPP_X=1"; do echo "== $v"; env $v python tools/profile_op.py contoursB 1000000 200 | grep -E "slice_order|pair_nodes"; done
for v in "PPP_X=1"; do echo "== 25 slices $v"; env $v python tools/profile_op.py contoursB 1000000 25 | grep -E "slice_order|pair_nodes"; done
for v in "PPP_X=1"; do echo "== 71 slices 11k $v"; env $v python tools/profile_op.py contoursB 3000000 71 | grep -E "slice_order|pair_nodes"; done
python tools/shape_perf.py 2>&1 | tail -3
python tools/step_jitter.py 200
