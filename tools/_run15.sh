timeout 1400 python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/latency_check.py 2>&1 | tail -4
