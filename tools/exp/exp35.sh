timeout 900 python -m pytest tests/test_gpu_edge.py -q -k "clusters or corner" 2>&1 | tail -15
for v in "PPP_SLICE_CLUSTER=1" "PPP_X=1" "PPP_SLICE_CLUSTER=2" "PPP_SLICE_CLUSTER=4"; do echo "== $v"; env $v python tools/profile_op.py contoursB 1000000 200 | grep -E "slice_order|pair_nodes"; done
for v in "PPP_SLICE_CLUSTER=1" "PPP_X=1" "PPP_SLICE_CLUSTER=2" "PPP_SLICE_CLUSTER=4"; do echo "== 8M/71 $v"; env $v python tools/profile_op.py contoursB 1000000 25 | grep -E "slice_order|pair_nodes"; done
for v in "PPP_SLICE_CLUSTER=1" "PPP_X=1"; do echo "== shapes $v"; env $v python tools/shape_perf.py 2>&1 | tail -3; done
