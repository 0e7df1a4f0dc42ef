set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_dist_gpu.py columns peer 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 40 --warmup 5 --store columns --no-e2e 2>gpurun_out/bench_g2_cols.err | tail -1 > gpurun_out/bench_g2_cols.json
tail -5 gpurun_out/bench_g2_cols.err; cat gpurun_out/bench_g2_cols.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 40 --warmup 5 --store replicated --no-e2e 2>gpurun_out/bench_g2_repl.err | tail -1 > gpurun_out/bench_g2_repl.json
cat gpurun_out/bench_g2_repl.json
