echo "== new"; python tools/step_jitter.py 300
echo "== 3phase scan"; PPP_SCAN_3PHASE=1 python tools/step_jitter.py 300
echo "== old knn"; PPP_KNN16_OLD=1 python tools/step_jitter.py 300
