timeout 1400 python -m pytest tests -m gpu -q 2>&1 | tail -5
python tools/latency_check.py 2>&1 | tail -4
