timeout 1400 python -m pytest tests -m gpu -q 2>&1 | tail -4
python tools/bo_iteration_bench.py 2>&1 | tail -3
