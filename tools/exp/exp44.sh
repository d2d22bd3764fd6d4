timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
PPP_DEBUG=1 python tools/shape_perf.py 2>&1 | grep -E "ms/step|fast path" | sort | uniq -c | sort -rn | head -14
