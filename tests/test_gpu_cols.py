"""-m gpu tests of the column-sharded multi-GPU layout (csrc/cols.cu) on ONE GPU: the G ranks are
emulated as one cooperative launch over all ranks' shards and mailboxes (B200_PROFILING.md: ranks
that wait on one another must not be separate launches on one GPU), so the complete exchange
protocol -- peer stores into mailboxes, self-validating words, re-arming, step parity -- runs as in
production.  Bars: update_phi / update_pi bit-identical to the one-GPU kernels of phi.cu (and so
independent of G), RNG pool states bit-identical, update_beta / perplexity against the oracle
within the fp32 tolerance of the one-GPU tests and identical on every rank."""
import os

import numpy as np
import pytest

import pyammsb as A
from test_gpu_parity import dev_params, dev_set, dev_store, link_heavy_problem, RTOL
from util import Problem, rel_err

pytestmark = pytest.mark.gpu


class Ranks:
    """G emulated ranks of the column layout on one device, loaded with a Problem's state"""

    def __init__(self, ctx, prob, G, Vcap, Ecap=0, Hcap=0):
        self.ctx, self.G = ctx, G
        self.r = [A.Cols(ctx, prob.N, prob.K, G, g, prob.n, Vcap, max(Ecap, 1), max(Hcap, 1)) for g in range(G)]
        for a in self.r:
            for b in self.r:
                if a is not b:
                    a.attach_local(b)
            a.write_pi(prob.pi)
            a.write_phi(prob.phi)
            a.write_theta(prob.theta, prob.beta)

    def pi(self):
        out = np.full((self.r[0].N, self.r[0].K), np.nan, np.float32)
        for a in self.r:
            a.read_pi(out)
        return out

    def phi_vec(self, V):
        out = np.full((V, self.r[0].K), np.nan, np.float32)
        for a in self.r:
            a.read_phi_vec(out)
        return out

    def check(self):
        for a in self.r:
            a.check()

    def free(self):
        for a in self.r:
            a.free()


def merged_pool(pools, G, wg=32):
    """RNG pool as one GPU would leave it: state of lane l from rank l % G"""
    states = [p.get_state() for p in pools]
    out = states[0].copy()
    lane = np.arange(out.shape[0]) % wg
    for g in range(1, G):
        out[lane % G == g] = states[g][lane % G == g]
    return out


@pytest.mark.parametrize("G", [2, 4, 8])
def test_cols_rows_roundtrip(ctx, orc, G):
    prob = Problem(orc, 300, 256, 3000, 8)
    rk = Ranks(ctx, prob, G, 64)
    assert np.array_equal(rk.pi(), prob.pi)
    for a in rk.r:
        assert np.array_equal(a.read_phi(), prob.phi)
        th, be = a.read_theta()
        assert np.array_equal(th, prob.theta) and np.array_equal(be, prob.beta)
    rk.free()


def one_gpu_phi_pi(ctx, prob, nodes, neighbors, step, noise):
    """ammsb_update_phi + ammsb_update_pi with the one-warp-per-slot kernel"""
    V, K = len(nodes), prob.K
    st = dev_store(ctx, prob)
    dset = dev_set(ctx, prob.train_set)
    units = min(V, 65535)
    pool = A.Rng(ctx, units * 32, 42, 43)
    d_nodes, d_nb, d_beta = ctx.from_host(nodes), ctx.from_host(neighbors), ctx.from_host(prob.beta)
    d_vec, d_sum = ctx.buf(np.float32, V * K), ctx.buf(np.float32, V)
    os.environ["AMMSB_PHI_NOSPLIT"] = "1"  # few slots: still the neighbor-by-neighbor association
    try:
        ctx.update_phi(dev_params(prob.p_orc), A.PhiOpts(A.MODE_WG, 32, 0 if noise else 1, 0), d_beta, st, dset,
                       d_nodes, d_nb, V, step, pool, d_vec, d_sum)
    finally:
        del os.environ["AMMSB_PHI_NOSPLIT"]
    vec = d_vec.read().reshape(V, K)
    ctx.update_pi(K, st, d_vec, d_sum, d_nodes, V)
    out = dict(vec=vec, pi=st.read_pi(), phi=st.read_phi(), pool=pool.get_state())
    for b in (d_nodes, d_nb, d_beta, d_vec, d_sum):
        b.free()
    pool.free(); dset.free(); st.free()
    return out


def cols_phi_pi(ctx, prob, G, nodes, neighbors, step, noise, rk=None, pools=None):
    V = len(nodes)
    own = rk is None
    if own:
        rk = Ranks(ctx, prob, G, V)
        pools = [A.Rng(ctx, min(V, 65535) * 32, 42, 43) for _ in range(G)]
    dset = dev_set(ctx, prob.train_set)
    d_nodes, d_nb = ctx.from_host(nodes), ctx.from_host(neighbors)
    A.cols_update_phi(ctx, rk.r, dev_params(prob.p_orc), A.PhiOpts(A.MODE_WG, 32, 0 if noise else 1, 0), dset, d_nodes,
                      d_nb, V, step, pools)
    ctx.sync()
    rk.check()
    vec = rk.phi_vec(V)
    A.cols_update_pi(ctx, rk.r, d_nodes, V, step)
    ctx.sync()
    rk.check()
    out = dict(vec=vec, pi=rk.pi(), phis=[a.read_phi() for a in rk.r], pool=merged_pool(pools, G))
    d_nodes.free(); d_nb.free(); dset.free()
    if own:
        for p in pools:
            p.free()
        rk.free()
    return out


def random_neighbors(prob, nodes, seed):
    rng = np.random.default_rng(seed)
    nb = rng.integers(0, prob.N, size=(len(nodes), prob.n), dtype=np.uint32)
    # make a share of the pairs training links so that both y branches are taken
    tr = prob.train_edges
    lo, hi = (tr >> np.uint64(32)).astype(np.uint32), (tr & np.uint64(0xffffffff)).astype(np.uint32)
    pos = {int(v): i for i, v in enumerate(nodes)}
    for a, b in zip(lo[:4000], hi[:4000]):
        for u, v in ((a, b), (b, a)):
            i = pos.get(int(u))
            if i is not None:
                nb[i, rng.integers(0, prob.n)] = v
    return np.ascontiguousarray(nb)


@pytest.mark.parametrize("K,G", [(1024, 8), (1024, 4), (1024, 2), (512, 8), (512, 2), (256, 4), (128, 8), (128, 2)])
@pytest.mark.parametrize("noise", [False, True])
def test_cols_phi_pi_bit_identical_to_one_gpu(ctx, orc, K, G, noise):
    prob = link_heavy_problem(orc, 700, K, 32, seed=K + G)
    for V in (257, 37):  # 37: a last group with idle sub-groups
        nodes = prob.minibatch_nodes(V, 5)
        nb = random_neighbors(prob, nodes, 9)
        one = one_gpu_phi_pi(ctx, prob, nodes, nb, 3, noise)
        col = cols_phi_pi(ctx, prob, G, nodes, nb, 3, noise)
        assert np.array_equal(col["vec"], one["vec"]), \
            "phi_vec differs: %d of %d elements, max rel %g" % ((col["vec"] != one["vec"]).sum(), one["vec"].size,
                                                               rel_err(col["vec"], one["vec"]).max())
        assert np.array_equal(col["pi"], one["pi"])
        for ph in col["phis"]:
            assert np.array_equal(ph, one["phi"])
        if noise:
            assert np.array_equal(col["pool"], one["pool"]), "RNG pool state differs from the one-GPU launch"


@pytest.mark.parametrize("n", [8, 16, 40, 64])
def test_cols_phi_neighbor_counts(ctx, orc, n):
    """n below the ring depth, n not a multiple of 32 (a short last segment), two full segments"""
    prob = link_heavy_problem(orc, 500, 256, n, seed=n)
    nodes = prob.minibatch_nodes(200, 5)
    nb = random_neighbors(prob, nodes, 3)
    one = one_gpu_phi_pi(ctx, prob, nodes, nb, 4, True)
    for G in (8, 4):
        col = cols_phi_pi(ctx, prob, G, nodes, nb, 4, True)
        assert np.array_equal(col["vec"], one["vec"]) and np.array_equal(col["pi"], one["pi"])
        assert np.array_equal(col["pool"], one["pool"])


def test_cols_phi_more_slots_than_units(ctx, orc):
    """V > 65535: a unit (and its RNG state) serves several slots, phi.cc:740-747"""
    N, K, n, V = 90000, 128, 8, 70001
    prob = Problem(orc, N, K, 200000, n)
    nodes = prob.minibatch_nodes(V, 5)
    nb = random_neighbors(prob, nodes, 3)
    one = one_gpu_phi_pi(ctx, prob, nodes, nb, 2, True)
    col = cols_phi_pi(ctx, prob, 8, nodes, nb, 2, True)
    assert np.array_equal(col["vec"], one["vec"]) and np.array_equal(col["pi"], one["pi"])
    assert np.array_equal(col["pool"], one["pool"])


def test_cols_steps_alternate_mailbox_halves(ctx, orc):
    """four consecutive iterations (both mailbox halves used twice, words re-armed in between) with
    a different mini-batch each, against the one-GPU kernels fed the same evolving state"""
    K, G, V = 512, 8, 300
    prob = link_heavy_problem(orc, 900, K, 32, seed=3)
    rk = Ranks(ctx, prob, G, V)
    pools = [A.Rng(ctx, V * 32, 42, 43) for _ in range(G)]
    pool_one = None
    for step in range(1, 5):
        nodes = prob.minibatch_nodes(V, 100 + step)
        nb = random_neighbors(prob, nodes, 200 + step)
        col = cols_phi_pi(ctx, prob, G, nodes, nb, step, True, rk=rk, pools=pools)
        # one GPU, same state: prob.pi / prob.phi are advanced below
        st = dev_store(ctx, prob)
        dset = dev_set(ctx, prob.train_set)
        if pool_one is None:
            pool_one = A.Rng(ctx, V * 32, 42, 43)
        d_nodes, d_nb, d_beta = ctx.from_host(nodes), ctx.from_host(nb), ctx.from_host(prob.beta)
        d_vec, d_sum = ctx.buf(np.float32, V * K), ctx.buf(np.float32, V)
        ctx.update_phi(dev_params(prob.p_orc), A.PhiOpts(A.MODE_WG, 32, 0, 0), d_beta, st, dset, d_nodes, d_nb, V, step,
                       pool_one, d_vec, d_sum)
        ctx.update_pi(K, st, d_vec, d_sum, d_nodes, V)
        prob.pi, prob.phi = st.read_pi(), st.read_phi()
        assert np.array_equal(col["pi"], prob.pi), "step %d" % step
        assert np.array_equal(col["pool"], pool_one.get_state())
        for b in (d_nodes, d_nb, d_beta, d_vec, d_sum):
            b.free()
        dset.free(); st.free()
    pool_one.free()
    for p in pools:
        p.free()
    rk.free()


@pytest.mark.parametrize("K,G,m", [(1024, 8, 300), (512, 4, 129), (256, 2, 64), (128, 8, 5)])
def test_cols_update_beta_vs_oracle(ctx, orc, K, G, m):
    prob = link_heavy_problem(orc, 500, K, 8)
    edges = prob.minibatch_edges(m, 3)
    scale, step = 17.5, 4
    theta_o, beta_o = prob.theta.copy(), prob.beta.copy()
    opool = orc.rng_pool(K, 44, 45)
    orc.update_beta(A.MODE_WG, 32, prob.p_orc, theta_o, beta_o, prob.pi, prob.train_set, edges, scale, step, opool)
    rk = Ranks(ctx, prob, G, 64, Ecap=m)
    pools = [A.Rng(ctx, K, 44, 45) for _ in range(G)]
    dset = dev_set(ctx, prob.train_set)
    d_edges = ctx.from_host(edges)
    A.cols_update_beta(ctx, rk.r, dev_params(prob.p_orc), dset, d_edges, m, scale, step, pools)
    ctx.sync()
    rk.check()
    th, be = rk.r[0].read_theta()
    for a in rk.r[1:]:  # every rank holds the published values of every column
        t2, b2 = a.read_theta()
        assert np.array_equal(t2, th) and np.array_equal(b2, be)
    # state k of the beta pool is advanced by the rank that owns column k
    owner = A.cols_owner(np.arange(K), G)
    st = np.stack([p.get_state() for p in pools])
    assert np.array_equal(st[owner, np.arange(K)], opool)
    et, eb = rel_err(th, theta_o), rel_err(be, beta_o)
    cond = np.abs(th.astype(np.float64) - theta_o) / (np.abs(theta_o) + prob.theta + np.abs(theta_o - prob.theta))
    print(f"K={K} G={G} m={m}: theta max rel {et.max():.3e} beta {eb.max():.3e} cond {cond.max():.3e}")
    assert cond.max() < 1e-6
    assert float((et > RTOL).mean()) < 5e-3 and np.median(et) < 1e-6
    assert float((eb > RTOL).mean()) < 5e-3 and np.median(eb) < 1e-6
    d_edges.free(); dset.free()
    for p in pools:
        p.free()
    rk.free()


@pytest.mark.parametrize("K,G", [(1024, 8), (256, 4), (128, 2)])
def test_cols_perplexity_vs_oracle(ctx, orc, K, G):
    prob = Problem(orc, 800, K, 6000, 8, heldout_ratio=0.2)
    H = len(prob.heldout_edges)
    ppx_o = np.zeros(H, dtype=np.float32)
    rk = Ranks(ctx, prob, G, 64, Hcap=H)
    dset = dev_set(ctx, prob.heldout_set)
    d_edges = ctx.from_host(prob.heldout_edges)
    for call in (1, 2, 3):
        avg_o, sums_o = orc.perplexity(A.MODE_THREAD, 32, prob.p_orc, prob.pi, prob.beta, prob.heldout_set,
                                       prob.heldout_edges, ppx_o, call)
        avg, sums = A.cols_perplexity(ctx, rk.r, dev_params(prob.p_orc), dset, d_edges, H, call)
        rk.check()
        for g in range(G):
            assert sums[g][2] == sums_o[2] == len(prob.heldout_links)
            assert sums[g][3] == sums_o[3] == len(prob.heldout_nonlinks)
            assert abs(avg[g] - avg_o) / abs(avg_o) < 1e-5
            assert avg[g] == avg[0]  # every rank finishes every pair from the same sums
    d_edges.free(); dset.free()
    rk.free()


@pytest.mark.parametrize("G,N,K", [(8, 300, 128), (4, 1000, 256)])
def test_cols_init_pi_equals_one_gpu_init(ctx, orc, G, N, K):
    st = A.Store(ctx, N, K)
    st.init_pi(1.0, 1.0)
    want_pi, want_phi = st.read_pi(), st.read_phi()
    st.free()
    prob = Problem(orc, N, K, 4 * N, 8)
    rk = Ranks(ctx, prob, G, 64)
    for a in rk.r:
        a.init_pi(1.0, 1.0)
    assert np.array_equal(rk.pi(), want_pi)
    for a in rk.r:
        assert np.array_equal(a.read_phi(), want_phi)
    rk.free()


@pytest.mark.parametrize("G,V,n", [(8, 300, 32), (4, 1000, 16), (2, 70001, 8)])
def test_cols_neighbor_sampler_equals_one_gpu_sampler(ctx, orc, G, V, n):
    """the sampler partitioned by state ownership (work-item gid on rank gid % G) delivers the lists
    of ammsb_neighbor_sample to every rank's mailbox; update_phi reads them there"""
    N, K = max(2 * V, 800), 128
    prob = Problem(orc, N, K, 6 * N, n)
    nodes = prob.minibatch_nodes(V, 5)
    d_nodes = ctx.from_host(nodes)
    pool1 = A.Rng(ctx, max(V, 64) * 2 * n, 56, 57)
    d_nb = ctx.buf(np.uint32, V * n)
    ctx.neighbor_sample(pool1, d_nodes, V, N, n, 32, d_nb)
    want = d_nb.read().reshape(V, n)
    rk = Ranks(ctx, prob, G, V)
    npools = [A.Rng(ctx, max(V, 64) * 2 * n, 56, 57) for _ in range(G)]
    step = 5
    A.cols_neighbor_sample(ctx, rk.r, d_nodes, V, 32, step, npools)
    # sampler state gid is advanced by rank gid % G only
    st = np.stack([p.get_state() for p in npools])
    gid = np.arange(st.shape[1])
    assert np.array_equal(st[gid % G, gid], pool1.get_state())
    # update_phi from the delivered lists == update_phi from the explicit ones (and == one GPU)
    one = one_gpu_phi_pi(ctx, prob, nodes, want, step, True)
    pools = [A.Rng(ctx, min(V, 65535) * 32, 42, 43) for _ in range(G)]
    dset = dev_set(ctx, prob.train_set)
    A.cols_update_phi(ctx, rk.r, dev_params(prob.p_orc), A.PhiOpts(A.MODE_WG, 32, 0, 0), dset, d_nodes, None, V, step, pools)
    ctx.sync()
    rk.check()
    assert np.array_equal(rk.phi_vec(V), one["vec"])
    for b in (d_nodes, d_nb):
        b.free()
    for p in pools + npools + [pool1]:
        p.free()
    dset.free(); rk.free()
