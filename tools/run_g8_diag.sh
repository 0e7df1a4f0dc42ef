#!/bin/bash
# 8-GPU stage breakdown of the sharded layouts (per-stage CUDA events inside the timed run)
run() { # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env AMMSB_STAGE_EVENTS=1 "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${G:-8} --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus ${G:-8} --steps 100 --warmup 10 --no-e2e "$@" 2>gpurun_out/diag_$name.err | tail -1 > gpurun_out/diag_$name.json
  echo "== $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/diag_$name.json'))
    print({k:d.get(k) for k in ('value','ms_per_step')}, d.get('parity_vs_n1','')[:12]); print(d.get('stages_in_run_ms')); print({k:d['roofline'].get(k) for k in ('achieved','frac','share_of_step')})
except Exception as e:
    print('no json', e)
PY
}
run cols -- --store columns
run cols_nolangevin AMMSB_COLS_DEBUG=16 AMMSB_BENCH_NO_PARITY=1 -- --store columns
run repl AMMSB_BENCH_NO_PARITY=1 -- --store replicated
