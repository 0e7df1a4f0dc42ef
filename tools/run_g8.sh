run() { # name, extra args
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 100 --warmup 10 "$@" 2>gpurun_out/r2_g8_$name.err | tail -1 > gpurun_out/r2_g8_$name.json
  echo "== $name rc=$?"; grep -E "parity|device graph|sampler" gpurun_out/r2_g8_$name.err | tail -4
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_g8_$name.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','heldout_perplexity','perplexity_eval_s')}); print('e2e', d['e2e'] and {k:d['e2e'][k] for k in ('value','ms_per_step')}); print({k:d['roofline'].get(k) for k in ('kernel','achieved','frac','share_of_step','nvlink_outbound_GBps')})
except Exception as e:
    print('no json', e)
PY
}
run dblp_cols --store columns
run dblp_repl --store replicated --no-e2e
run lj_cols --store columns --shape com-LiveJournal --graph device
run fr_cols --store columns --shape com-Friendster --graph device
