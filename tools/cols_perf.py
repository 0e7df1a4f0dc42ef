#!/usr/bin/env python3
"""Timing of the column-sharded kernels on ONE GPU (development tool).

  emu   the whole G-rank job on one GPU (G emulated ranks, one cooperative launch): the HBM
        traffic is that of the one-GPU kernel, so the time beside ammsb_update_phi's says what the
        split into pieces + the exchange protocol cost (the "NVLink" is the local L2 here).
  loop  ONE rank's share of a G-GPU weak-scaled step (V = G*m + 1 slots, K/G columns) with the
        exchange waits switched off (AMMSB_COLS_LOOPBACK): the HBM/issue side of the production
        kernel without NVLink latency.
usage: cols_perf.py [K] [G] [m]   -- tuning through AMMSB_COLS_WARPS / _R / _D
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-ammsb-gpu_b200")]
import devgraph  # noqa: E402
import pyammsb as A  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
N, E, n = 317080, 1049866, 32
PEAK = 6550.0
ctx = A.Ctx(0)
g = devgraph.DeviceGraph(ctx, N, E, 0.1)
rng = np.random.default_rng(0)
p = A.make_params(N, E, K, n)
theta = rng.gamma(1.0, 1.0, 2 * K).astype(np.float32)
beta = (theta.reshape(K, 2) / theta.reshape(K, 2).sum(1, keepdims=True)).astype(np.float32).ravel()


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    ctx.sync()
    ts = []
    for _ in range(iters):
        ctx.timer_start()
        fn()
        ts.append(ctx.timer_stop_ms())
    return float(np.median(ts))


def minibatch(V):
    d_nodes = ctx.from_host(rng.permutation(N)[:V].astype(np.uint32))
    d_nb = ctx.buf(np.uint32, V * n)
    npool = A.Rng(ctx, V * 2 * n, 56, 57)
    ctx.neighbor_sample(npool, d_nodes, V, N, n, 32, d_nb)
    d_edges = ctx.from_host(rng.integers(0, N, size=V, dtype=np.uint64) << np.uint64(32) |
                            rng.integers(0, N, size=V, dtype=np.uint64))
    npool.free()
    return d_nodes, d_nb, d_edges


def bytes_phi(V, Kc):
    return V * ((n + 2) * 4 * Kc + n * 68 + 8)


def sweep(label, fn, nbytes, configs):
    for w, R, D in configs:
        os.environ.update(AMMSB_COLS_WARPS=str(w), AMMSB_COLS_R=str(R), AMMSB_COLS_D=str(D))
        try:
            t = timeit(fn)
            print("%-40s warps=%d R=%2d D=%d  %8.4f ms  %8.1f GB/s  %5.1f%% of %d" %
                  (label, w, R, D, t, nbytes / t / 1e6, 100 * nbytes / t / 1e6 / PEAK, PEAK), flush=True)
        except A.AmmsbError as e:
            print("%-40s warps=%d R=%2d D=%d  failed: %s" % (label, w, R, D, e), flush=True)


configs = [(6, 6, 3), (6, 6, 2), (5, 7, 3), (4, 10, 4), (4, 8, 3), (8, 4, 2), (8, 5, 2), (7, 5, 2)]
if os.environ.get("COLS_PERF_CONFIGS"):
    configs = [tuple(int(x) for x in c.split(",")) for c in os.environ["COLS_PERF_CONFIGS"].split(";")]
mode = os.environ.get("COLS_PERF_MODE", "emu,loop")

# ---- one GPU reference ----
V1 = m + 1
d_nodes, d_nb, d_edges = minibatch(V1)
store = A.Store(ctx, N, K)
store.init_pi()
d_beta = ctx.from_host(beta)
d_vec, d_sum = ctx.buf(np.float32, V1 * K), ctx.buf(np.float32, V1)
ppool = A.Rng(ctx, V1 * 32, 42, 43)
step = [0]


def one():
    step[0] += 1
    ctx.update_phi(p, A.PhiOpts(A.MODE_WG, 32, 0, 0), d_beta, store, g.train, d_nodes, d_nb, V1, step[0], ppool, d_vec,
                   d_sum)


t = timeit(one)
print("one GPU k_update_phi_fast   V=%d K=%d: %.4f ms  %.1f GB/s (%.1f%%)" %
      (V1, K, t, bytes_phi(V1, K) / t / 1e6, 100 * bytes_phi(V1, K) / t / 1e6 / PEAK), flush=True)
t = timeit(lambda: ctx.update_pi(K, store, d_vec, d_sum, d_nodes, V1))
print("one GPU k_update_pi: %.4f ms" % t)
for b in (d_vec, d_sum):
    b.free()
store.free()

if "emu" in mode:
    ranks = [A.Cols(ctx, N, K, G, r, n, V1, V1, 1) for r in range(G)]
    for a in ranks:
        for b in ranks:
            if a is not b:
                a.attach_local(b)
        a.init_pi()
        a.write_theta(theta, beta)
    pools = [A.Rng(ctx, V1 * 32, 42, 43) for _ in range(G)]
    bpools = [A.Rng(ctx, K, 44, 45) for _ in range(G)]

    def emu():
        step[0] += 1
        A.cols_update_phi(ctx, ranks, p, A.PhiOpts(A.MODE_WG, 32, 0, 0), g.train, d_nodes, d_nb, V1, step[0], pools)

    def emu_pi():
        emu()
        A.cols_update_pi(ctx, ranks, d_nodes, V1, step[0])

    def emu_beta():
        step[0] += 1
        A.cols_update_beta(ctx, ranks, p, g.train, d_edges, m, 2.0 * E / m, step[0], bpools)

    sweep("emulated %d ranks, V=%d" % (G, V1), emu, bytes_phi(V1, K), configs)
    t0 = timeit(emu)
    t1 = timeit(emu_pi)
    print("emulated update_pi: %.4f ms (phi %.4f, phi+pi %.4f)" % (t1 - t0, t0, t1))
    t = timeit(emu_beta)
    print("emulated update_beta (m=%d): %.4f ms  %.1f GB/s" % (m, t, (m * (8 * K + 72) + 24 * K) / t / 1e6))
    for a in ranks:
        a.check()
    for a in ranks + pools + bpools:
        a.free()

if "loop" in mode:
    os.environ["AMMSB_COLS_LOOPBACK"] = "1"
    Vg = G * m + 1
    for b in (d_nodes, d_nb, d_edges):
        b.free()
    d_nodes, d_nb, d_edges = minibatch(Vg)
    # only rank 0 computes and loopback never waits for a peer: every peer mailbox pointer is an
    # alias of the own mailbox (ammsb_cols_alias_self, a diagnostic entry point)
    r0 = A.Cols(ctx, N, K, G, 0, n, Vg, Vg, 1)
    r0.alias_self()
    r0.init_pi()
    r0.write_theta(theta, beta)
    pools = [A.Rng(ctx, min(Vg, 65535) * 32, 42, 43)]
    bpools = [A.Rng(ctx, K, 44, 45)]

    def loop():
        step[0] += 1
        A.cols_update_phi(ctx, [r0], p, A.PhiOpts(A.MODE_WG, 32, 0, 0), g.train, d_nodes, d_nb, Vg, step[0], pools)

    def loop_pi():
        loop()
        A.cols_update_pi(ctx, [r0], d_nodes, Vg, step[0])

    def loop_beta():
        step[0] += 1
        A.cols_update_beta(ctx, [r0], p, g.train, d_edges, G * m, 2.0 * E / m, step[0], bpools)

    for dbg in os.environ.get("COLS_PERF_DEBUGS", "0").split(","):
        os.environ["AMMSB_COLS_DEBUG"] = dbg
        for nb in os.environ.get("COLS_PERF_NBS", "2").split(","):
            os.environ["AMMSB_COLS_NB"] = nb
            sweep("1 of %d, V=%d loopback dbg=%s NB=%s" % (G, Vg, dbg, nb), loop, bytes_phi(Vg, K // G), configs)
    os.environ["AMMSB_COLS_DEBUG"] = "0"
    del os.environ["AMMSB_COLS_NB"]
    t0 = timeit(loop)
    t1 = timeit(loop_pi)
    print("loopback update_pi: %.4f ms (phi %.4f, phi+pi %.4f)" % (t1 - t0, t0, t1))
    t = timeit(loop_beta)
    print("loopback update_beta (m=%d): %.4f ms  %.1f GB/s" % (G * m, t, (G * m * (8 * K // G + 72) + 24 * K) / t / 1e6))
