timeout 300 python -m pytest tests/test_gpu_cols.py -x -q --timeout 200 2>&1 | tail -12
