#!/usr/bin/env python3
"""Peer-gather probe: update_phi on GPU 0 with pi node-partitioned over 2 GPUs of ONE process
(direct peer access, no IPC import), to separate the cost of NVLink row gathers from the cost of
how the peer memory is mapped.  Usage: peer_probe.py [N] [K]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-ammsb-gpu_b200")]
import pyammsb as A  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3997962
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n, V = 32, 16385
c0, c1 = A.Ctx(0), A.Ctx(1)
s0, s1 = A.Store(c0, N, K, 2, 0), A.Store(c1, N, K, 2, 1)
s0.attach_local(1, s1)
s1.attach_local(0, s0)
s0.init_pi()
s1.init_pi()
c0.sync(); c1.sync()
rng = np.random.default_rng(0)
nodes = rng.permutation(N)[:V].astype(np.uint32)
nbrs = rng.integers(0, N, size=(V, n), dtype=np.uint32)
table = np.full(8 * 1024, 2 ** 64 - 1, dtype=np.uint64)  # empty edge set
dset = A.DevSet(c0, table, 1024, 0)
p = A.make_params(N, 10 * N, K, n)
d_nodes, d_nb = c0.from_host(nodes), c0.from_host(nbrs)
d_beta = c0.from_host(np.full(2 * K, 0.5, np.float32))
d_vec, d_sum = c0.buf(np.float32, V * K), c0.buf(np.float32, V)
pool = A.Rng(c0, V * 32, 42, 43)
for part, label in ((A.PhiOpts(A.MODE_WG, 32, 0, 0, 0, 2), "half of the slots (rank 0 of 2)"),
                    (A.PhiOpts(A.MODE_WG, 32, 0, 0), "all slots")):
    ts = []
    for it in range(8):
        c0.timer_start()
        c0.update_phi(p, part, d_beta, s0, dset, d_nodes, d_nb, V, it + 1, pool, d_vec, d_sum)
        ts.append(c0.timer_stop_ms())
    t = float(np.median(ts[3:]))
    slots = V / 2 if part.part_count == 2 else V
    remote = slots * n * 4 * K * 0.5
    print("N=%d K=%d %s: %.3f ms, remote rows %.3f GB -> %.1f GB/s inbound over NVLink (direct peer access)"
          % (N, K, label, t, remote / 1e9, remote / t / 1e6))
