python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_1gpu.json 2> gpurun_out/r02h_bench_1gpu.err; tail -c 300 gpurun_out/r02h_bench_1gpu.json
python tools/step_timeline.py > gpurun_out/r02h_timeline.txt 2>&1
