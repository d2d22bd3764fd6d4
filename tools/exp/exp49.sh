timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/step_jitter.py 200
python tools/step_timeline.py | tail -16
python tools/profile_op.py contoursB 1000000 5000 | head -5
