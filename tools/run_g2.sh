timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 30 --warmup 5 --store columns 2>gpurun_out/bench_g2_cols.err | tail -1 > gpurun_out/bench_g2_cols.json
tail -3 gpurun_out/bench_g2_cols.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_g2_cols.json'))
print({k:d[k] for k in ('value','ms_per_step','parity_vs_n1','heldout_perplexity')}); print({k:d['e2e'][k] for k in ('value','ms_per_step')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 30 --warmup 5 --store columns --shape com-Friendster-eighth --graph device 2>gpurun_out/bench_g2_fr8.err | tail -1 > gpurun_out/bench_g2_fr8.json
tail -8 gpurun_out/bench_g2_fr8.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_g2_fr8.json'))
print({k:d[k] for k in ('value','ms_per_step','parity_vs_n1','heldout_perplexity','perplexity_eval_s')}); print({k:d['e2e'][k] for k in ('value','ms_per_step')}); print(d['roofline'])
PY
