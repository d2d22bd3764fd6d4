timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg3 > gpurun_out/r02f_ncu_bench.log 2>&1
tail -2 gpurun_out/r02f_ncu_bench.log | cut -c1-200
