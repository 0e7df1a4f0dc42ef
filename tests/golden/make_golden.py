#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own code (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the development container only:
    python oracle/build_ref.py && python tests/golden/make_golden.py
The fixtures are committed; /root/reference is not needed to *check* them."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import pyoracle  # noqa: E402
from util import Problem, make_edges  # noqa: E402

ref = pyoracle.Oracle(path=os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so"))
L = ref.L


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


# ---- RNG known answers (random.cl.inc) ----
pool = ref.rng_pool(8, 42, 43)
u64 = ref.draw_u64(pool, 16)
pool = ref.rng_pool(8, 42, 43)
randn = ref.draw_randn(pool, 64)
state_after_randn = pool.copy()
pool = ref.rng_pool(8, 11, 113)
gamma = ref.draw_gamma(pool, 32, 1.0, 1.0)
pool = ref.rng_pool(8, 11, 113)
gamma_half = ref.draw_gamma(pool, 32, 0.5, 2.0)
save("rng.npz", u64=u64, randn=randn, state_after_randn=state_after_randn, gamma=gamma,
     gamma_half=gamma_half,
     rounded=np.array([ref.round_param(x) for x in (1 / 64, 1 / 1024, 0.0315, 1024, 0.5, 1e-7, 1.0)],
                      dtype=np.float32))

# ---- a small seeded problem: every operator once ----
prob = Problem(ref, 300, 64, 9000, 16, seed=7)
p = prob.p_orc
nodes = prob.minibatch_nodes(65, 5)
npool = ref.rng_pool(65 * 32, 56, 57)
neighbors, table = ref.neighbor_sample(npool, nodes, prob.N, prob.n, 32)
out = dict(N=prob.N, K=prob.K, E=prob.E, n=prob.n, train_edges=prob.train_edges,
           heldout_edges=prob.heldout_edges, n_heldout_links=len(prob.heldout_links),
           train_table=prob.train_set.table(), train_bins=prob.train_set.num_bins,
           train_prime=prob.train_set.prime_idx, heldout_table=prob.heldout_set.table(),
           heldout_bins=prob.heldout_set.num_bins, heldout_prime=prob.heldout_set.prime_idx,
           pi=prob.pi, phi=prob.phi, theta=prob.theta, beta=prob.beta, nodes=nodes,
           neighbors=neighbors, sampler_table=table, sampler_state=npool)
for mode, tag in ((pyoracle.MODE_WG, "wg"), (pyoracle.MODE_THREAD, "thread")):
    states = 65 * (32 if mode == pyoracle.MODE_WG else 1)
    for noise in (0, 1):
        pool = ref.rng_pool(states, 42, 43)
        v = ref.update_phi(mode, 32, p, prob.beta, prob.pi, prob.phi, prob.train_set, nodes, neighbors,
                           3, pool, disable_noise=not noise)
        out["phi_vec_%s_%d" % (tag, noise)] = v
        if noise:
            out["phi_state_%s" % tag] = pool
            pi2, phi2 = prob.pi.copy(), prob.phi.copy()
            ref.update_pi(mode, 32, prob.K, pi2, phi2, v, nodes)
            out["pi_after_%s" % tag] = pi2[nodes]
            out["phi_after_%s" % tag] = phi2[nodes]
edges = prob.minibatch_edges(48, 3)
theta, beta = prob.theta.copy(), prob.beta.copy()
bpool = ref.rng_pool(prob.K, 44, 45)
ts, g = ref.update_beta(pyoracle.MODE_THREAD, 32, p, theta, beta, prob.pi, prob.train_set, edges, 17.5, 4,
                        bpool)
out.update(mb_edges=edges, theta_sum=ts, grads=g, theta_after=theta, beta_after=beta, beta_state=bpool)
ppx = np.zeros(len(prob.heldout_edges), dtype=np.float32)
avgs, sums = [], []
for call in (1, 2, 3):
    a, s = ref.perplexity(pyoracle.MODE_THREAD, 32, p, prob.pi, prob.beta, prob.heldout_set,
                          prob.heldout_edges, ppx, call)
    avgs.append(a)
    sums.append(s)
out.update(ppx_avg=np.array(avgs), ppx_sums=np.array(sums), ppx_per_edge=ppx)
pi0, phi0 = ref.init_pi(200, 48)
out.update(init_pi=pi0, init_phi=phi0)
save("operators.npz", **out)

# ---- host logic of the reference: split, graph, mini-batch strategies ----
L.ref_generate_sets.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_double, C.c_uint, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]
L.ref_sampler_create.restype = C.c_void_p
L.ref_sampler_create.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                 C.c_uint64]
L.ref_sampler_max_fan_out.restype = C.c_uint64
L.ref_sampler_max_fan_out.argtypes = [C.c_void_p]
L.ref_sample.restype = C.c_float
L.ref_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

Ng, Eg, m = 500, 3000, 24
keys = make_edges(Ng, Eg, 3)
tr = np.zeros(Eg, dtype=np.uint64)
he = np.zeros(Eg, dtype=np.uint64)
ntr, nhe = C.c_uint64(0), C.c_uint64(0)
ok = L.ref_generate_sets(Ng, keys.ctypes.data_as(C.c_void_p), Eg, 0.1, 12345, tr.ctypes.data_as(C.c_void_p),
                         C.byref(ntr), he.ctypes.data_as(C.c_void_p), C.byref(nhe))
assert ok
tr, he = tr[:ntr.value], he[:nhe.value]
n_links = Eg - len(tr)
h = L.ref_sampler_create(Ng, Eg, tr.ctypes.data_as(C.c_void_p), len(tr), he.ctypes.data_as(C.c_void_p),
                         n_links, m)
host = dict(N=Ng, E=Eg, m=m, edges=keys, srand_seed=12345, heldout_ratio=0.1, training=tr, heldout=he,
            max_fan_out=L.ref_sampler_max_fan_out(h))
for strat, name in enumerate(["Node", "NodeLink", "NodeNonLink", "BFLink", "BFNonLink", "BF"]):
    seed = C.c_uint(1000 + strat)
    all_e, all_n, meta = [], [], []
    for it in range(6):
        eb = np.zeros(4096, dtype=np.uint64)
        nb = np.zeros(8192, dtype=np.uint32)
        ne, nn = C.c_uint64(0), C.c_uint64(0)
        w = L.ref_sample(h, strat, C.byref(seed), eb.ctypes.data_as(C.c_void_p), C.byref(ne),
                         nb.ctypes.data_as(C.c_void_p), C.byref(nn))
        all_e.append(eb[:ne.value])
        all_n.append(nb[:nn.value])
        meta.append((ne.value, nn.value, w, seed.value))
    host["mb_%s_edges" % name] = np.concatenate(all_e)
    host["mb_%s_nodes" % name] = np.concatenate(all_n)
    host["mb_%s_meta" % name] = np.array(meta, dtype=np.float64)
L.ref_sampler_destroy(h)
save("host.npz", **host)

# ---- SNAP text loader (data.cc:36-78): a small file with a 4-line header, repeated and reversed
# pairs, sparse vertex ids ----
L.ref_unique_edges_from_file.restype = C.c_int64
L.ref_unique_edges_from_file.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64]
rng = np.random.default_rng(11)
ids = rng.choice(100000, size=300, replace=False)
pairs = [(int(ids[a]), int(ids[b])) for a, b in rng.integers(0, 300, size=(2500, 2)) if a != b]
pairs += [(b, a) for a, b in pairs[:200]] + pairs[200:300]
snap_text = "# Undirected graph: synthetic\n# test fixture\n# Nodes: 300 Edges: %d\n# FromNodeId\tToNodeId\n" % len(pairs)
snap_text += "".join("%d\t%d\n" % p for p in pairs)
snap_path = "/tmp/ammsb_golden_snap.txt"
open(snap_path, "w").write(snap_text)
out_e = np.zeros(len(pairs), dtype=np.uint64)
nv = C.c_uint64(0)
n = L.ref_unique_edges_from_file(snap_path.encode(), 4321, C.byref(nv), out_e.ctypes.data_as(C.c_void_p), len(out_e))
assert n > 0
save("snap.npz", text=np.frombuffer(snap_text.encode(), dtype=np.uint8), srand_seed=4321, N=nv.value, edges=out_e[:n])
