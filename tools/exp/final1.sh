set -x
T=r02f
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_1gpu_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg3 > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knn16f -s 3 -c 1 -f -o gpurun_out/${T}_knn16f python tools/profile_op.py knn16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_knnf -s 3 -c 1 -f -o gpurun_out/${T}_knn32f python tools/profile_op.py knn32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pair_nodes|k_slice_order|k_knn_warp" -s 9 -c 3 -f -o gpurun_out/${T}_slicing python bench.py --steps 1 --warmup 3 --no-cfg3 > /dev/null 2>&1
python tools/sweep.py > gpurun_out/${T}_sweep_one_gpu.txt 2>&1
python tools/shape_perf.py 2>&1 | tail -3 > gpurun_out/${T}_shapes.txt; cat gpurun_out/${T}_shapes.txt
python tools/step_timeline.py > gpurun_out/${T}_timeline.txt 2>&1
python tools/cfg1_main_flow.py > gpurun_out/${T}_cfg1.txt 2>&1; tail -5 gpurun_out/${T}_cfg1.txt
ls -la gpurun_out/${T}_*.ncu-rep
