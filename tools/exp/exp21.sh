python tools/debug_knn.py 8 4000 1
python tools/debug_knn.py 16 4000 1 | head -3
python tools/debug_knn.py 4 4000 1 | head -3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_shapes.py -q 2>&1 | tail -15
python tools/step_timeline.py > gpurun_out/timeline_cfg2_c.txt 2>&1; tail -22 gpurun_out/timeline_cfg2_c.txt
python tools/step_breakdown.py
