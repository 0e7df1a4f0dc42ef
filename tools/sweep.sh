#!/bin/bash
# K / mini-batch / neighbor-count sweep at the DBLP shape (BASELINE.json configs[4]); run on a B200.
cd "$(dirname "$0")/.."
for K in 64 128 256 512 1024 2048 4096; do
  timeout 300 python tools/microbench.py --K $K --iters 10 2>&1 | grep -E "GB/s" | sed "s/^/K=$K m=16384 n=32 | /"
done
for m in 1024 4096 65536; do
  timeout 300 python tools/microbench.py --K 1024 --m $m --iters 10 2>&1 | grep -E "GB/s" | sed "s/^/K=1024 m=$m n=32 | /"
done
for n in 64 128; do
  timeout 300 python tools/microbench.py --K 1024 --n $n --iters 10 2>&1 | grep -E "GB/s" | sed "s/^/K=1024 m=16384 n=$n | /"
done
