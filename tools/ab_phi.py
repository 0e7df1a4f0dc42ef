#!/usr/bin/env python3
"""A/B timing of update_phi variants at the DBLP shape, one process for all K (development tool):
default dispatch against the switches of csrc/phi.cu (AMMSB_PHI_LATE_NOISE, AMMSB_PHI_EARLY_1024,
noise off).  The graph is built in HBM, so the whole sweep takes seconds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-ammsb-gpu_b200")]
import devgraph  # noqa: E402
import pyammsb as A  # noqa: E402

N, E, m, n = 317080, 1049866, 16384, 32
Ks = [int(k) for k in sys.argv[1:]] or [64, 128, 256, 512, 1024]
ctx = A.Ctx(0)
g = devgraph.DeviceGraph(ctx, N, E, 0.1)
V = m + 1
rng = np.random.default_rng(0)
d_nodes = ctx.from_host(rng.permutation(N)[:V].astype(np.uint32))
d_nb = ctx.buf(np.uint32, V * n)
npool = A.Rng(ctx, 2 * m * 2 * n, 56, 57)
ctx.neighbor_sample(npool, d_nodes, V, N, n, 32, d_nb)
ppool = A.Rng(ctx, 2 * m * 32, 42, 43)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ctx.sync()
    ts = []
    for _ in range(iters):
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop_ms())
    return float(np.median(ts))


print("K      variant                ms      GB/s algorithmic   of 6550")
for K in Ks:
    p = A.make_params(N, E, K, n)
    store = A.Store(ctx, N, K)
    store.init_pi()
    theta = rng.gamma(1.0, 1.0, 2 * K).astype(np.float32).reshape(K, 2)
    d_beta = ctx.from_host((theta / theta.sum(1, keepdims=True)).astype(np.float32).ravel())
    d_vec, d_sum = ctx.buf(np.float32, V * K), ctx.buf(np.float32, V)
    nbytes = V * ((n + 2) * 4 * K + n * 68 + 8)
    step = [0]

    def run(noise=True):
        step[0] += 1
        ctx.update_phi(p, A.PhiOpts(A.MODE_WG, 32, 0 if noise else 1, 0), d_beta, store, g.train, d_nodes, d_nb, V,
                       step[0], ppool, d_vec, d_sum)

    variants = [("default", None, True), ("noise off", None, False)]
    variants.insert(1, ("AMMSB_PHI_LATE_NOISE", "AMMSB_PHI_LATE_NOISE", True) if K <= 512 else
                    ("AMMSB_PHI_EARLY_1024", "AMMSB_PHI_EARLY_1024", True))
    for name, env, noise in variants:
        if env:
            os.environ[env] = "1"
        t = timeit(lambda: run(noise))
        if env:
            del os.environ[env]
        print("%-6d %-22s %7.4f %9.1f %14.1f%%" % (K, name, t, nbytes / t / 1e6, 100 * nbytes / t / 1e6 / 6550))
    for b in (d_beta, d_vec, d_sum):
        b.free()
    store.free()
