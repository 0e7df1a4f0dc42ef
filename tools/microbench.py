#!/usr/bin/env python3
"""Per-stage microbenchmark at a named shape (development tool; bench.py is the contract)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-ammsb-gpu_b200"), os.path.join(ROOT, "oracle"),
                os.path.join(ROOT, "tests")]
import pyammsb as A  # noqa: E402
import pyoracle  # noqa: E402
from util import make_edges, split_edges, fake_nonlinks  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=317080)
ap.add_argument("--E", type=int, default=1049866)
ap.add_argument("--K", type=int, default=1024)
ap.add_argument("--m", type=int, default=16384)
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--H", type=int, default=104986)
args = ap.parse_args()

N, E, K, m, n = args.N, args.E, args.K, args.m, args.n
orc = pyoracle.Oracle()
ctx = A.Ctx(0)
print(ctx.name(), "SMs", ctx.sm_count())
t0 = time.time()
keys = make_edges(N, E, 1)
train, held = split_edges(keys, 0.1)
tset = orc.set_build(train)
hset = orc.set_build(held)
print("graph+sets %.1fs" % (time.time() - t0))
p = A.make_params(N, E, K, n)
store = A.Store(ctx, N, K)
store.init_pi()
dts = A.DevSet(ctx, tset.table(), tset.num_bins, tset.prime_idx)
dhs = A.DevSet(ctx, hset.table(), hset.num_bins, hset.prime_idx)
V = m + 1
rng = np.random.default_rng(0)
nodes = rng.permutation(N)[:V].astype(np.uint32)
d_nodes = ctx.from_host(nodes)
d_nb = ctx.buf(np.uint32, V * n)
npool = A.Rng(ctx, 2 * m * 2 * n, 56, 57)
ppool = A.Rng(ctx, 2 * m * 32, 42, 43)
bpool = A.Rng(ctx, K, 44, 45)
theta = rng.gamma(1.0, 1.0, 2 * K).astype(np.float32)
d_theta = ctx.from_host(theta)
d_beta = ctx.from_host(orc.theta_to_beta(theta))
d_vec, d_sum = ctx.buf(np.float32, V * K), ctx.buf(np.float32, V)
u = nodes[0]
ev = rng.integers(0, N, size=m).astype(np.uint64)
edges = (np.minimum(ev, u).astype(np.uint64) << np.uint64(32)) | np.maximum(ev, u).astype(np.uint64)
d_edges = ctx.from_host(edges)
d_ts, d_g = ctx.buf(np.float32, K), ctx.buf(np.float32, 2 * K)
ws = ctx.buf(np.uint8, ctx.beta_workspace_bytes(K))
H = args.H
hn = (rng.integers(0, N, size=H).astype(np.uint64) << np.uint64(32)) | rng.integers(0, N, size=H).astype(np.uint64)
hedges = np.concatenate([held[:H // 2], hn[:H - min(H // 2, len(held))]])[:H]
H = len(hedges)
d_hedges = ctx.from_host(hedges)
d_ppx = ctx.buf(np.float32, H).zero()
pws = ctx.buf(np.uint8, ctx.perplexity_workspace_bytes())
opts = A.PhiOpts(A.MODE_WG, 32, 0, 0)
ctx.sync()


def timeit(name, fn, nbytes):
    for _ in range(3):
        fn()
    ctx.sync()
    ts = []
    for _ in range(args.iters):
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop_ms())
    t = float(np.median(ts))
    print("%-16s %9.3f ms  (min %.3f)  %8.1f GB/s algorithmic  (%.1f MB)" %
          (name, t, min(ts), nbytes / t / 1e6, nbytes / 1e6))
    return t


step = [0]


def f_ns():
    ctx.neighbor_sample(npool, d_nodes, V, N, n, 32, d_nb)


def f_phi():
    step[0] += 1
    ctx.update_phi(p, opts, d_beta, store, dts, d_nodes, d_nb, V, step[0], ppool, d_vec, d_sum)


def f_pi():
    ctx.update_pi(K, store, d_vec, d_sum, d_nodes, V)


def f_beta():
    ctx.update_beta(p, d_theta, d_beta, store, dts, d_edges, m, 2.0 * E / m, step[0] + 1, bpool, d_ts,
                    d_g, ws)


call = [0]


def f_ppx():
    call[0] += 1
    ctx.perplexity(p, store, d_beta, dhs, d_hedges, H, d_ppx, call[0], pws)


timeit("neighbor_sample", f_ns, V * (4 + 4 * n + 32))
timeit("update_phi", f_phi, V * ((n + 2) * 4 * K + n * 68 + 8))
timeit("update_pi", f_pi, V * (8 * K + 8))
timeit("update_beta", f_beta, m * (8 * K + 72) + 24 * K)
timeit("perplexity", f_ppx, H * (8 * K + 88) + 8 * K)
opts_nn = A.PhiOpts(A.MODE_WG, 32, 1, 0)


def f_phi_nn():
    ctx.update_phi(p, opts_nn, d_beta, store, dts, d_nodes, d_nb, V, 5, ppool, d_vec, d_sum)


timeit("update_phi(no noise)", f_phi_nn, V * ((n + 2) * 4 * K + n * 68 + 8))
print("launches", A.launch_count())
