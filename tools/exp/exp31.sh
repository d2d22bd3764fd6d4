PPP_KNN16_OLD=1 PPP_DEBUG=1 python tools/shape_perf.py 2>&1 | grep -E "fast path|ms/step" | sort | uniq -c | sort -rn | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn16f -s 3 -c 1 -f -o gpurun_out/r02c_knn16f python tools/profile_op.py knn16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knnf -s 3 -c 1 -f -o gpurun_out/r02c_knn32f python tools/profile_op.py knn32 > /dev/null 2>&1
ls -la gpurun_out/r02c_*.ncu-rep
