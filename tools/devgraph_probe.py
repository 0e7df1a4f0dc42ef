"""Build a synthetic graph of a named shape in HBM (no model) and draw a few device mini-batches:
sizes, times and basic invariants.  python tools/devgraph_probe.py com-Friendster 131072"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mcmc-ammsb-gpu_b200"))
import devgraph  # noqa: E402
import pyammsb as A  # noqa: E402
import synth  # noqa: E402

shape = sys.argv[1] if len(sys.argv) > 1 else "com-Friendster"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
N, E, K, r = synth.SHAPES[shape]
ctx = A.Ctx(0)
t0 = time.time()
g = devgraph.DeviceGraph(ctx, N, E, r, seed=1, log=print)
print("total build %.1fs; degree mean %.1f max %d; H %d" % (time.time() - t0, g.degree.mean(), g.max_fan_out, g.H))
assert int(g.degree.sum()) == 2 * g.num_training
pairs = g.d_heldout_pairs.read()
links, fakes = pairs[:g.num_heldout_links], pairs[g.num_heldout_links:]
assert g.heldout.has(links[:100000]).all() and not g.heldout.has(fakes[:100000]).any()
assert not g.train.has(links[:100000]).any() and not g.train.has(fakes[:100000]).any()
smp = g.sampler(m)
d_e, d_n = ctx.buf(np.uint64, g.max_edges(m)), ctx.buf(np.uint32, g.max_nodes(m))
seed = C.c_uint(12345)
t0 = time.time()
for i in range(10):
    w, ne, nn = smp.sample(seed, d_e, d_n)
    ctx.sync()
    e, v = d_e.read(ne), d_n.read(nn)
    assert len(np.unique(e)) == ne and len(np.unique(v)) == nn
    if ne == m:
        assert not g.train.has(e).any() and not g.heldout.has(e).any()
    else:
        assert g.train.has(e).all() and ne == g.degree[v[0]]
    print("mini-batch %d: %d edges %d nodes weight %g" % (i, ne, nn, w))
print("10 mini-batches (with read-back) %.2fs" % (time.time() - t0))
