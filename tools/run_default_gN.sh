#!/bin/bash
# the driver's own launch of the contract run on G GPUs (defaults), with stage events: bash tools/run_default_gN.sh G
G=${1:-8}
AMMSB_STAGE_EVENTS=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $G --steps 200 --warmup 20 2>gpurun_out/r2_default_g$G.err | tail -1 > gpurun_out/r2_default_g$G.json
echo "rc=$?"; grep -E "parity|Error|error|sampler" gpurun_out/r2_default_g$G.err | tail -4
python - <<PY
import json
d=json.load(open('gpurun_out/r2_default_g$G.json'))
print({k:d.get(k) for k in ('value','ms_per_step','heldout_perplexity')}, d['config']['parallelism'][:30]); print('e2e', d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step','h2d_bytes_per_step')}, d['e2e'] and d['e2e']['api'][:110]); print(d.get('stages_in_run_ms')); print(d['roofline'])
PY
