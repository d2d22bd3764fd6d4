timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tools/shape_perf.py 2>&1 | tail -3
python tools/step_jitter.py 200
python tools/step_timeline.py > gpurun_out/timeline_cfg2_d.txt 2>&1; tail -22 gpurun_out/timeline_cfg2_d.txt
