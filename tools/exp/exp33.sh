echo "== jitter PDL on (wait only)"; python tools/step_jitter.py 200
echo "== jitter PDL off"; PPP_PDL=0 python tools/step_jitter.py 200
echo "== jitter PDL on (wait only)"; python tools/step_jitter.py 200
echo "== jitter PDL off"; PPP_PDL=0 python tools/step_jitter.py 200
for v in "PPP_PDL=0" "PPP_X=1"; do echo "== $v"; env $v python tools/profile_op.py knn16 2>&1 | grep -E "knn_normals|knn_redo"; done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
