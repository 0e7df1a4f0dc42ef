timeout 600 python -m pytest tests/test_gpu_graph.py -x -q --timeout 300 -k "orderset or bit_identical" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_gpu_cols.py -x -q --timeout 200 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -s --timeout 500 2>&1 | grep -vE "^\s*$" | tail -45
export COLS_PERF_MODE=loop COLS_PERF_CONFIGS="16,0,0;12,0,0" COLS_PERF_DEBUGS="0" COLS_PERF_NBS="2"
timeout 250 python tools/cols_perf.py 512 8 2>&1 | grep -E "loopback|failed|emulated|one GPU"
