set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_1gpu_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg3 > gpurun_out/r02_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn16c -s 3 -c 1 -f -o gpurun_out/r02_knn16c python tools/profile_op.py knn16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn32c -s 3 -c 1 -f -o gpurun_out/r02_knn32c python tools/profile_op.py knn32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_nodes|k_slice_order|k_knn_warp" -s 9 -c 3 -f -o gpurun_out/r02_slicing python bench.py --steps 1 --warmup 3 --no-cfg3 > /dev/null 2>&1
python tools/sweep.py > gpurun_out/r02_sweep_one_gpu.txt 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
python tools/profile_op.py knn64 | head -3
python tools/profile_op.py contoursB 1000000 5000 | head -4
python tools/shape_perf.py 2>&1 | tail -3 > gpurun_out/r02_shapes.txt; cat gpurun_out/r02_shapes.txt
