set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn16c -s 3 -c 1 -f -o gpurun_out/r02a_knn16c python tools/profile_op.py knn16 > gpurun_out/exp2_ncu.log 2>&1
ls -la gpurun_out/
