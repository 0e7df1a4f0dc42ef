#!/usr/bin/env python3
"""mcmc::ShardedLearner (C++, one process, one rank per device) on the DBLP shape: iterations/s through
Run() with host mini-batches.  usage: cxx_sharded_bench.py G [K] [m_per_gpu]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-ammsb-gpu_b200")]
import pymcmc  # noqa: E402
import synth  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128 * G
m = (int(sys.argv[3]) if len(sys.argv) > 3 else 16384) * G
N, E = 317080, 1049866
cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=32, heldout_ratio=0.1, strategy="Node")
cfg.set_graph(N, synth.make_edges(N, E, 1))
lrn = pymcmc.ShardedLearner(cfg, list(range(G)))
lrn.run(20)
e0, t0 = lrn.edges_processed(), time.perf_counter()
lrn.run(200)
dt = time.perf_counter() - t0
print("mcmc::ShardedLearner %d GPUs K=%d m=%d: %.3f ms/step, %.1f M edges/s, perplexity %.4f" %
      (G, K, m, 1e3 * dt / 200, (lrn.edges_processed() - e0) / dt / 1e6, lrn.heldout_perplexity()))
