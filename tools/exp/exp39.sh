for cfg in "3000000 71" "2000000 141" "1400000 100"; do
for v in "PPP_SLICE_CLUSTER=1" "PPP_X=1" "PPP_SLICE_CLUSTER=2" "PPP_SLICE_CLUSTER=4"; do echo "== $cfg $v"; env $v python tools/profile_op.py contoursB $cfg | grep -E "slice_order"; done
done
