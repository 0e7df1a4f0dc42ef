"""-m gpu tests of the graph inputs built on the device (csrc/graph.cu): cuckoo build, synthetic
edge lists, adjacency, and the Node mini-batch strategy -- against the host implementations
(libmcmc.so, themselves pinned to the reference's data.cc / cuckoo.cc / sample.cc by
tests/test_host.py) and the CPU oracle.  Everything here is integer work: the bar is bit-exact
(set equality where the device emits in another order, and that is said in the test)."""
import ctypes as C

import numpy as np
import pytest

import pyammsb as A
import pymcmc
import pyoracle
from util import RefSampler, make_edges

pytestmark = pytest.mark.gpu


def edge_keys(u, v):
    u, v = np.asarray(u, dtype=np.uint64), np.asarray(v, dtype=np.uint64)
    return (np.minimum(u, v) << np.uint64(32)) | np.maximum(u, v)


@pytest.mark.parametrize("n", [1, 7, 1000, 250000])
def test_set_build_membership_and_geometry(ctx, orc, n):
    """device-built cuckoo set: every key found, no phantom members, table geometry = the host's;
    the table image also answers correctly under the oracle's Set_HasEdge (the reference lookup)"""
    rng = np.random.default_rng(n)
    N = 100000
    keys = np.unique(edge_keys(rng.integers(0, N, 2 * n), rng.integers(0, N, 2 * n)))[:n]
    rng.shuffle(keys)
    s = A.BuiltSet(ctx, keys)
    assert s.num_bins == int(pymcmc.lib().mcmc_host_set_bins(len(keys)))
    probe = np.concatenate([keys, edge_keys(rng.integers(0, N, 50000), rng.integers(0, N, 50000))])
    want = np.isin(probe, keys).astype(np.uint8)
    assert np.array_equal(s.has(probe), want)
    table = s.table()
    stored = table[table != np.uint64(0xFFFFFFFFFFFFFFFF)]
    assert np.array_equal(np.unique(stored), np.sort(keys))  # nothing lost, nothing invented
    # the reference's lookup on the device-built image
    st = pyoracle.SetStruct(table.ctypes.data, int(s.num_bins), int(s.prime_idx), 0)
    got = np.zeros(len(probe), dtype=np.uint8)
    orc.L.orc_set_has_many(C.byref(st), probe.ctypes.data_as(C.c_void_p), C.c_uint64(len(probe)),
                             got.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, want)
    # built from a device buffer: same membership
    d = ctx.from_host(keys)
    s2 = A.BuiltSet(ctx, d, len(keys))
    assert np.array_equal(s2.has(probe), want)
    d.free(); s.free(); s2.free()


def test_set_build_duplicates_and_empty(ctx):
    keys = edge_keys([1, 1, 2, 2, 2], [5, 5, 9, 9, 9])
    s = A.BuiltSet(ctx, keys)
    assert s.has(edge_keys([1, 2, 3], [5, 9, 4])).tolist() == [1, 1, 0]
    s.free()
    e = A.BuiltSet(ctx, np.zeros(0, dtype=np.uint64))
    assert e.has(edge_keys([1], [2])).tolist() == [0]
    e.free()


def test_graph_generate_unique_canonical_deterministic(ctx):
    N, E = 5000, 60000
    bufs = [ctx.buf(np.uint64, E) for _ in range(3)]
    A.graph_generate(ctx, N, E, 1, bufs[0])
    A.graph_generate(ctx, N, E, 1, bufs[1])
    A.graph_generate(ctx, N, E, 2, bufs[2])
    a, b, c = (x.read() for x in bufs)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    u, v = (a >> np.uint64(32)).astype(np.int64), (a & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert (u < v).all() and (v < N).all()
    assert len(np.unique(a)) == E
    assert not np.array_equal(a, np.sort(a))  # shuffled (data.cc:62)
    # endpoints roughly uniform: every vertex decile holds its share of endpoints
    hist = np.histogram(np.concatenate([u, v]), bins=10, range=(0, N))[0]
    assert hist.min() > 0.9 * 2 * E / 10 and hist.max() < 1.1 * 2 * E / 10
    for x in bufs:
        x.free()


def test_graph_nonlinks_avoid_both_sets(ctx):
    N, E = 2000, 30000
    keys = make_edges(N, E, 5)
    a, b = A.BuiltSet(ctx, keys[:20000]), A.BuiltSet(ctx, keys[20000:])
    out = ctx.buf(np.uint64, 5000)
    A.graph_nonlinks(ctx, N, 5000, 3, a, b, out)
    got = out.read()
    assert len(np.unique(got)) == 5000
    assert not np.isin(got, keys).any()
    u, v = got >> np.uint64(32), got & np.uint64(0xFFFFFFFF)
    assert (u < v).all() and (v < N).all()
    A.graph_nonlinks(ctx, N, 5000, 3, a, None, out)
    got = out.read()
    assert not np.isin(got, keys[:20000]).any() and len(np.unique(got)) == 5000
    out.free(); a.free(); b.free()


def test_graph_csr_matches_host_adjacency(ctx):
    N, E = 3000, 25000
    keys = make_edges(N, E, 8)
    d_e, d_off, d_adj, d_deg = ctx.from_host(keys), ctx.buf(np.uint64, N + 1), ctx.buf(np.uint32, 2 * E), \
        ctx.buf(np.uint32, N)
    A.graph_csr(ctx, N, d_e, E, d_off, d_adj, d_deg)
    off, adj, deg = d_off.read(), d_adj.read(), d_deg.read()
    u, v = (keys >> np.uint64(32)).astype(np.int64), (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    src, dst = np.concatenate([u, v]), np.concatenate([v, u])
    order = np.lexsort((dst, src))
    assert np.array_equal(adj, dst[order].astype(np.uint32))
    assert np.array_equal(deg, np.bincount(src, minlength=N).astype(np.uint32))
    assert off[0] == 0 and off[-1] == 2 * E and np.array_equal(np.diff(off.astype(np.int64)), deg)
    for x in (d_e, d_off, d_adj, d_deg):
        x.free()


@pytest.mark.parametrize("N,E,m,built", [(20000, 120000, 4096, False), (3000, 40000, 2500, True),
                                         (500, 3000, 24, False), (317080, 1049866, 16384, True)])
def test_device_sampler_draws_the_host_strategys_minibatches(ctx, N, E, m, built):
    """the device Node strategy against the host strategy (sample.cc, pinned to the reference by
    tests/test_host.py) on the same seed, a run of mini-batches: the same edges, the same nodes
    (compared as sets: the device emits in draw order, the host in std::unordered_set order), the
    same weight, and the same rand_r state afterwards -- so the streams never drift apart.  The
    dense 3000-vertex graph makes the first candidate pass fall short (refusals + repeats)."""
    keys = make_edges(N, E, 9)
    cfg = pymcmc.Config(K=8, mini_batch_size=m, heldout_ratio=0.1, strategy="Node")
    cfg.set_graph(N, keys, srand_seed=5)
    tr, he = cfg.edges()
    # when the reference-derived library travelled with the repo: the reference's own sample.cc too
    ref = RefSampler(N, keys, 0.1, 5, m) if RefSampler.available() else None
    seed_r = C.c_uint(77)
    if built:
        n_links = len(he) // 2  # held-out links come first (data.cc:86-100); the set holds only those
        train, heldout = A.BuiltSet(ctx, tr), A.BuiltSet(ctx, he[:n_links])
    else:
        train = A.DevSet(ctx, *cfg.set_table(0))
        heldout = A.DevSet(ctx, *cfg.set_table(1))
    d_tr = ctx.from_host(tr)
    d_off, d_adj, d_deg = ctx.buf(np.uint64, N + 1), ctx.buf(np.uint32, 2 * len(tr)), ctx.buf(np.uint32, N)
    A.graph_csr(ctx, N, d_tr, len(tr), d_off, d_adj, d_deg)
    degree = d_deg.read()
    assert degree.max() == cfg.max_fan_out()
    smp = A.DeviceSampler(ctx, N, E, m, train, heldout, d_off, d_adj, degree)
    d_edges = ctx.buf(np.uint64, max(m, int(degree.max())))
    d_nodes = ctx.buf(np.uint32, max(m, int(degree.max())) + 1)
    seed_h, seed_d = C.c_uint(77), C.c_uint(77)
    kinds = set()
    for _ in range(16):
        w_h, e_h, n_h = cfg.sample("Node", seed_h)
        if ref is not None:
            w_r, e_r, n_r = ref.sample("Node", seed_r)
            assert np.array_equal(e_r, e_h) and np.array_equal(n_r, n_h) and seed_r.value == seed_h.value
        w_d, ne, nn = smp.sample(seed_d, d_edges, d_nodes)
        ctx.sync()
        e_d, n_d = d_edges.read(ne), d_nodes.read(nn)
        assert (ne, nn) == (len(e_h), len(n_h))
        assert np.array_equal(np.sort(e_d), np.sort(e_h))
        assert np.array_equal(np.sort(n_d), np.sort(n_h))
        assert len(np.unique(n_d)) == nn
        assert np.float32(w_d) == np.float32(w_h)
        assert seed_d.value == seed_h.value
        if ne == m:
            # a non-link mini-batch is emitted in draw order, which is the order in which the host
            # strategy inserts into its std::unordered_set: putting the device's edges through that
            # container's ordering gives the host's mini-batch exactly, order included
            assert np.array_equal(pymcmc.set_order(e_d, 8)[1], e_h)
            ends, cnt = np.unique(np.concatenate([e_d >> np.uint64(32), e_d & np.uint64(0xFFFFFFFF)]),
                                  return_counts=True)
            assert n_d[0] == ends[cnt.argmax()]  # the shared endpoint u comes first
        kinds.add(ne == m)
    assert kinds == {True, False}  # both link and non-link mini-batches were drawn
    for x in (d_tr, d_off, d_adj, d_deg, d_edges, d_nodes):
        x.free()
    smp.free(); train.free(); heldout.free(); cfg.close()
    if ref is not None:
        ref.close()


def host_order_csr(N, tr):
    """mcmc::Graph (data.cc:12-25): for every training edge (u, v) in list order, v joins u's
    neighbors and u joins v's -- the order sampleNodeLink walks"""
    u, v = (tr >> np.uint64(32)).astype(np.int64), (tr & np.uint64(0xffffffff)).astype(np.int64)
    ends = np.stack([u, v], axis=1).ravel()       # u0, v0, u1, v1, ...
    other = np.stack([v, u], axis=1).ravel()
    order = np.argsort(ends, kind="stable")        # per vertex, in order of appearance
    deg = np.bincount(ends, minlength=N)
    off = np.zeros(N + 1, np.uint64)
    off[1:] = np.cumsum(deg)
    return off, other[order].astype(np.uint32), deg.astype(np.uint32)


@pytest.mark.parametrize("n", [1, 5, 13, 14, 100, 1109, 1110, 2357, 2358, 5000, 40000])
def test_orderset_is_libstdcxx_iteration_order(ctx, n):
    """keys in insertion order -> std::unordered_set iteration order, against the host container
    (host/mcmc/std_order_set.h, itself tested against std::unordered_set) at sizes around the
    growth steps of _Prime_rehash_policy and across the one-CTA / radix-sort switch"""
    rng = np.random.default_rng(n)
    for wide in (False, True):
        keys = rng.choice(1 << 22, size=n, replace=False).astype(np.uint64)
        if wide:
            keys = (keys << np.uint64(32)) | rng.integers(0, 1 << 22, n).astype(np.uint64)
        want = pymcmc.unordered_set_order(keys)
        os_ = C.c_void_p()
        A._ck(A.lib().ammsb_orderset_create(ctx.h, max(n, 4), C.byref(os_)))
        d_in, d_out = ctx.from_host(keys), ctx.buf(np.uint64, n)
        A._ck(A.lib().ammsb_orderset_apply(os_, ctx.h, d_in.ptr, n, d_out.ptr))
        ctx.sync()
        assert np.array_equal(d_out.read(), want)
        A.lib().ammsb_orderset_destroy(os_)
        d_in.free(); d_out.free()


@pytest.mark.parametrize("N,E,m", [(3000, 60000, 500), (20000, 100000, 4096), (317080, 1049866, 16384)])
def test_device_node_strategy_is_bit_identical_to_the_host_strategy(ctx, N, E, m):
    """with the reference's emission order computed on the device (csrc/orderset.cu) and the
    adjacency in the host Graph's order, the device Node strategy + node extraction equals
    sampleNode (sample.cc:253-302) + ExtractNodesFromMiniBatch (learner.cc:162-173) element by
    element, mini-batch after mini-batch, with the same rand_r state afterwards"""
    keys = make_edges(N, E, 9)
    cfg = pymcmc.Config(K=8, mini_batch_size=m, heldout_ratio=0.1, strategy="Node")
    cfg.set_graph(N, keys, srand_seed=5)
    tr, he = cfg.edges()
    ref = RefSampler(N, keys, 0.1, 5, m) if RefSampler.available() and N <= 20000 else None
    train = A.DevSet(ctx, *cfg.set_table(0))
    heldout = A.DevSet(ctx, *cfg.set_table(1))
    off, adj, deg = host_order_csr(N, tr)
    assert deg.max() == cfg.max_fan_out()
    d_off, d_adj = ctx.from_host(off), ctx.from_host(adj)
    smp = A.DeviceSampler(ctx, N, E, m, train, heldout, d_off, d_adj, deg, exact_order=True)
    cap = max(m, int(deg.max()))
    d_edges, d_nodes = ctx.buf(np.uint64, cap), ctx.buf(np.uint32, 2 * cap + 2)
    seed_h, seed_d, seed_r = C.c_uint(77), C.c_uint(77), C.c_uint(77)
    kinds = set()
    for _ in range(12):
        w_h, e_h, n_h = cfg.sample("Node", seed_h)
        if ref is not None:
            w_r, e_r, n_r = ref.sample("Node", seed_r)
            assert np.array_equal(e_r, e_h) and np.array_equal(n_r, n_h)
        w_d, ne, nn = smp.sample(seed_d, d_edges, d_nodes)
        ctx.sync()
        assert (ne, nn) == (len(e_h), len(n_h))
        assert np.array_equal(d_edges.read(ne), e_h), "edges are not in the reference's order"
        assert np.array_equal(d_nodes.read(nn), n_h), "nodes are not in the reference's order"
        assert np.float32(w_d) == np.float32(w_h) and seed_d.value == seed_h.value
        kinds.add(ne == m)
    assert kinds == {True, False}
    for b in (d_off, d_adj, d_edges, d_nodes):
        b.free()
    smp.free(); train.free(); heldout.free()
    if ref is not None:
        ref.close()
    cfg.close()
