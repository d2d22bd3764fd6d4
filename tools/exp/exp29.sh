timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
for op in knn16 knn32 knn64; do echo "== $op"; python tools/profile_op.py $op > /tmp/o.txt 2>&1; grep -E "knn_normals|knn_redo" /tmp/o.txt; done
echo "== contoursB 200"; python tools/profile_op.py contoursB 1000000 200 | head -8
echo "== contoursB 5000"; python tools/profile_op.py contoursB 1000000 5000 | head -5
echo "== contoursA"; python tools/profile_op.py contoursA 1000000 200 | head -4
python tools/shape_perf.py 2>&1 | tail -3
python tools/step_jitter.py 200
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; tail -c 1500 gpurun_out/r02c_bench_1gpu.json
