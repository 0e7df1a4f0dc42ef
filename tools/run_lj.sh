#!/bin/bash
# com-LiveJournal shape on G GPUs, column-sharded, graph built in HBM: bash tools/run_lj.sh G
G=${1:-2}
AMMSB_STAGE_EVENTS=1 AMMSB_BENCH_VARIANTS="${VARIANTS:-nosplit:AMMSB_COLS_NOSPLIT=1}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $G --steps 100 --warmup 10 --store columns --shape com-LiveJournal --graph device 2>gpurun_out/r2_lj_g${G}_cols.err | tail -1 > gpurun_out/r2_lj_g${G}_cols.json
echo "rc=$?"; tail -4 gpurun_out/r2_lj_g${G}_cols.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2_lj_g${G}_cols.json'))
print({k:d.get(k) for k in ('value','ms_per_step','heldout_perplexity')}, d.get('parity_vs_n1','')[:40]); print('e2e', d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')}); print(d.get('stages_in_run_ms')); print({k:(v['value'],v['ms_per_step'],v.get('stages_in_run_ms')) for k,v in d.get('variants',{}).items()})
PY
