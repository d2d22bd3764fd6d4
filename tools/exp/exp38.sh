timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tools/profile_op.py contoursB 1000000 200 | head -12
PPP_BANDS_SEARCH=1 python tools/profile_op.py contoursB 1000000 200 | grep -E "band_"
python tools/profile_op.py contoursB 1000000 5000 | head -5
python tools/step_jitter.py 200
python tools/step_timeline.py 2>&1 | tail -20
