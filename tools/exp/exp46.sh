timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench_1gpu.json 2> gpurun_out/r02g_bench_1gpu.err; tail -c 600 gpurun_out/r02g_bench_1gpu.json
python tools/shape_perf.py 2>&1 | tail -3
