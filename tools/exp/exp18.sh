set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/step_timeline.py > gpurun_out/timeline_cfg2.txt 2>&1; tail -40 gpurun_out/timeline_cfg2.txt
python tools/step_breakdown.py
