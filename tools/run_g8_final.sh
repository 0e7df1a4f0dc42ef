#!/bin/bash
# 8-GPU runs of the column-sharded layout with per-stage events and the k_cols_phi3 A/B in the same process
run() { # name, args
  name=$1; shift
  AMMSB_STAGE_EVENTS=1 AMMSB_BENCH_VARIANTS="${VARIANTS:-}" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 100 --warmup 10 "$@" 2>gpurun_out/r2b_g8_$name.err | tail -1 > gpurun_out/r2b_g8_$name.json
  echo "== $name rc=$?"; grep -E "parity|Error|error" gpurun_out/r2b_g8_$name.err | tail -3
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b_g8_$name.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','heldout_perplexity')}); print('e2e', d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')}); print(d.get('stages_in_run_ms')); print({k:(v['value'],v['ms_per_step'],v.get('stages_in_run_ms')) for k,v in d.get('variants',{}).items()})
except Exception as e:
    print('no json', e)
PY
}
run dblp_cols --store columns
run fr_cols --store columns --shape com-Friendster --graph device
