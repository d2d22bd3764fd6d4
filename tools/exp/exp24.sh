timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn16f -s 3 -c 1 -f -o gpurun_out/r02b_knn16f python tools/profile_op.py knn16 > /dev/null 2>&1
ls -la gpurun_out/r02b_knn16f.ncu-rep
