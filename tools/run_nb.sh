timeout 300 python -m pytest tests/test_gpu_cols.py -x -q --timeout 200 2>&1 | tail -12
export COLS_PERF_CONFIGS="12,0,0;10,0,0;8,0,0" COLS_PERF_DEBUGS="0,31,16,1" COLS_PERF_NBS="2"
timeout 250 python tools/cols_perf.py 1024 8 2>&1 | grep -E "loopback|failed|emulated|one GPU"
