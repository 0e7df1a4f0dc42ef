"""CPU tests (-m "not gpu"): pin the oracle.

  1. oracle restatement == golden vectors generated from the reference's own code
     (tests/golden/*.npz, made by tests/golden/make_golden.py from oracle/_ref);
  2. oracle restatement == oracle/_ref live, when the reference-derived library is present
     (this container; it also travels to the GPU box);
  3. the exact properties the reference's own unit tests assert."""
import ctypes as C
import os

import numpy as np
import pytest

import pyoracle
from util import Problem, make_edges, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference); golden vectors still apply")
    return pyoracle.Oracle(path=REF_SO)


def params_from(orc, g):
    return orc.make_params(int(g["N"]), int(g["E"]), int(g["K"]), int(g["n"]))


class GoldSet:
    """golden cuckoo table wrapped like an OracleSet"""

    def __init__(self, orc, table, bins, prime):
        self.table_arr = np.ascontiguousarray(table, dtype=np.uint64)
        self.s = pyoracle.SetStruct(self.table_arr.ctypes.data, int(bins), int(prime), 0)
        self.orc, self.num_bins, self.prime_idx = orc, int(bins), int(prime)


# ------------------------------------------------------------------ golden ----

def test_golden_rng(orc):
    g = np.load(os.path.join(GOLD, "rng.npz"))
    pool = orc.rng_pool(8, 42, 43)
    assert np.array_equal(orc.draw_u64(pool, 16), g["u64"])
    assert g["u64"][0, :4].tolist() == [352324268, 360712944, 2955487616344221, 5981343618259954]
    pool = orc.rng_pool(8, 42, 43)
    assert np.array_equal(orc.draw_randn(pool, 64), g["randn"])
    assert np.array_equal(pool, g["state_after_randn"])
    pool = orc.rng_pool(8, 11, 113)
    assert np.array_equal(orc.draw_gamma(pool, 32, 1.0, 1.0), g["gamma"])
    pool = orc.rng_pool(8, 11, 113)
    assert np.array_equal(orc.draw_gamma(pool, 32, 0.5, 2.0), g["gamma_half"])
    got = np.array([orc.round_param(x) for x in (1 / 64, 1 / 1024, 0.0315, 1024, 0.5, 1e-7, 1.0)],
                   dtype=np.float32)
    assert np.array_equal(got, g["rounded"])


def test_golden_operators(orc):
    g = np.load(os.path.join(GOLD, "operators.npz"))
    p = params_from(orc, g)
    K, n, N = int(g["K"]), int(g["n"]), int(g["N"])
    # cuckoo build reproduces the reference's table bit for bit
    ts = orc.set_build(g["train_edges"])
    assert np.array_equal(ts.table(), g["train_table"])
    assert (ts.num_bins, ts.prime_idx) == (int(g["train_bins"]), int(g["train_prime"]))
    nl = int(g["n_heldout_links"])
    hs = orc.set_build(g["heldout_edges"][:nl])
    assert np.array_equal(hs.table(), g["heldout_table"])
    # sampler
    nodes = g["nodes"]
    npool = orc.rng_pool(len(nodes) * 2 * n, 56, 57)
    nb, tab = orc.neighbor_sample(npool, nodes, N, n, 32)
    assert np.array_equal(nb, g["neighbors"]) and np.array_equal(tab, g["sampler_table"])
    assert np.array_equal(npool, g["sampler_state"])
    pi, phi, beta = g["pi"], g["phi"], g["beta"]
    for mode, tag in ((pyoracle.MODE_WG, "wg"), (pyoracle.MODE_THREAD, "thread")):
        states = len(nodes) * (32 if mode == pyoracle.MODE_WG else 1)
        for noise in (0, 1):
            pool = orc.rng_pool(states, 42, 43)
            v = orc.update_phi(mode, 32, p, beta, pi, phi, ts, nodes, nb, 3, pool, disable_noise=not noise)
            assert np.array_equal(v, g["phi_vec_%s_%d" % (tag, noise)]), (tag, noise)
            if noise:
                assert np.array_equal(pool, g["phi_state_%s" % tag])
                pi2, phi2 = pi.copy(), phi.copy()
                orc.update_pi(mode, 32, K, pi2, phi2, v, nodes)
                assert np.array_equal(pi2[nodes], g["pi_after_%s" % tag])
                assert np.array_equal(phi2[nodes], g["phi_after_%s" % tag])
    theta, b2 = g["theta"].copy(), beta.copy()
    bpool = orc.rng_pool(K, 44, 45)
    tsum, grads = orc.update_beta(pyoracle.MODE_THREAD, 32, p, theta, b2, pi, ts, g["mb_edges"], 17.5, 4, bpool)
    assert np.array_equal(tsum, g["theta_sum"]) and np.array_equal(grads, g["grads"])
    assert np.array_equal(theta, g["theta_after"]) and np.array_equal(b2, g["beta_after"])
    assert np.array_equal(bpool, g["beta_state"])
    ppx = np.zeros(len(g["heldout_edges"]), dtype=np.float32)
    for i, call in enumerate((1, 2, 3)):
        a, s = orc.perplexity(pyoracle.MODE_THREAD, 32, p, pi, beta, hs, g["heldout_edges"], ppx, call)
        assert a == g["ppx_avg"][i] and np.array_equal(s, g["ppx_sums"][i])
    assert np.array_equal(ppx, g["ppx_per_edge"])
    pi0, phi0 = orc.init_pi(200, 48)
    assert np.array_equal(pi0, g["init_pi"]) and np.array_equal(phi0, g["init_phi"])


# -------------------------------------------------------------- live vs _ref ----

@pytest.mark.parametrize("N,K,E,n,V", [(600, 96, 24000, 32, 257), (400, 33, 9000, 7, 100),
                                       (1500, 128, 30000, 16, 300)])
def test_oracle_equals_reference_code(orc, ref, N, K, E, n, V):
    prob = Problem(orc, N, K, E, n, seed=N)
    p = prob.p_orc
    s2 = ref.set_build(prob.train_edges)
    assert np.array_equal(prob.train_set.table(), s2.table())
    assert (prob.train_set.prime_idx, prob.train_set.count) == (s2.prime_idx, s2.count)
    probe = np.concatenate([prob.train_edges, prob.heldout_edges])
    assert np.array_equal(prob.train_set.has(probe), s2.has(probe))
    nodes = prob.minibatch_nodes(V, 11)
    n1, n2 = orc.rng_pool(V * 2 * n, 56, 57), ref.rng_pool(V * 2 * n, 56, 57)
    nb1, h1 = orc.neighbor_sample(n1, nodes, N, n, 32)
    nb2, h2 = ref.neighbor_sample(n2, nodes, N, n, 32)
    assert np.array_equal(nb1, nb2) and np.array_equal(h1, h2) and np.array_equal(n1, n2)
    for mode in (pyoracle.MODE_THREAD, pyoracle.MODE_WG):
        for noise in (False, True):
            q1, q2 = orc.rng_pool(V * 32, 42, 43), ref.rng_pool(V * 32, 42, 43)
            v1 = orc.update_phi(mode, 32, p, prob.beta, prob.pi, prob.phi, prob.train_set, nodes, nb1, 3, q1,
                                not noise)
            v2 = ref.update_phi(mode, 32, p, prob.beta, prob.pi, prob.phi, s2, nodes, nb1, 3, q2, not noise)
            assert np.array_equal(v1, v2) and np.array_equal(q1, q2)
        pa, fa, pb, fb = prob.pi.copy(), prob.phi.copy(), prob.pi.copy(), prob.phi.copy()
        orc.update_pi(mode, 32, K, pa, fa, v1, nodes)
        ref.update_pi(mode, 32, K, pb, fb, v1, nodes)
        assert np.array_equal(pa, pb) and np.array_equal(fa, fb)
    edges = prob.minibatch_edges(64, 3)
    th1, be1, th2, be2 = prob.theta.copy(), prob.beta.copy(), prob.theta.copy(), prob.beta.copy()
    b1, b2 = orc.rng_pool(K, 44, 45), ref.rng_pool(K, 44, 45)
    r1 = orc.update_beta(pyoracle.MODE_THREAD, 32, p, th1, be1, prob.pi, prob.train_set, edges, 17.5, 4, b1)
    r2 = ref.update_beta(pyoracle.MODE_THREAD, 32, p, th2, be2, prob.pi, s2, edges, 17.5, 4, b2)
    assert all(np.array_equal(x, y) for x, y in zip(r1, r2))
    assert np.array_equal(th1, th2) and np.array_equal(be1, be2) and np.array_equal(b1, b2)
    hs2 = ref.set_build(prob.heldout_links)
    for mode in (pyoracle.MODE_THREAD, pyoracle.MODE_WG):
        x1 = np.zeros(len(prob.heldout_edges), np.float32)
        x2 = x1.copy()
        for call in (1, 2):
            a1 = orc.perplexity(mode, 32, p, prob.pi, prob.beta, prob.heldout_set, prob.heldout_edges, x1, call)
            a2 = ref.perplexity(mode, 32, p, prob.pi, prob.beta, hs2, prob.heldout_edges, x2, call)
            assert a1[0] == a2[0] and np.array_equal(a1[1], a2[1])
        assert np.array_equal(x1, x2)
    i1, i2 = orc.init_pi(257, K), ref.init_pi(257, K)
    assert np.array_equal(i1[0], i2[0]) and np.array_equal(i1[1], i2[1])


def test_oracle_rng_equals_reference_code(orc, ref):
    for sx, sy in ((42, 43), (11, 113), (0, 1), (2**63, 2**64 - 5)):
        p1, p2 = orc.rng_pool(32, sx, sy), ref.rng_pool(32, sx, sy)
        assert np.array_equal(p1, p2)
        assert np.array_equal(orc.draw_u64(p1, 40), ref.draw_u64(p2, 40))
        assert np.array_equal(orc.draw_randn(p1, 3000), ref.draw_randn(p2, 3000))
        for a, b in ((1.0, 1.0), (0.3, 1.5), (2.5, 0.1)):
            assert np.array_equal(orc.draw_gamma(p1, 200, a, b), ref.draw_gamma(p2, 200, a, b))
        assert np.array_equal(p1, p2)
    for x in (1 / 3, 1 / 64, 0.0315, 1e-7, 123456.789, 1 / 1024):
        assert orc.round_param(x) == ref.round_param(x)


# ---------------------------------------- the reference's own test properties ----

def test_cuckoo_property(orc):
    # cuckoo-test.cc:29-43: first half+1 of random keys inserted; those found, the rest not
    rng = np.random.default_rng(0)
    keys = np.unique(rng.integers(0, 2**63, size=200_000, dtype=np.uint64))
    rng.shuffle(keys)
    half = len(keys) // 2 + 1
    s = orc.set_build(keys[:half])
    got = s.has(keys)
    assert got[:half].all() and not got[half:].any()
    assert s.num_bins == 1 + int(np.ceil(1.15 * half / 8))  # cuckoo.cc:100-101


def test_rng_pool_property(orc):
    # random-test.cc:58-63
    pool = orc.rng_pool(1000, 42, 43)
    i = np.arange(1000, dtype=np.uint64)
    assert np.array_equal(pool[:, 0], 42 + i) and np.array_equal(pool[:, 1], 43 + i)


def test_sampler_properties(orc):
    # wg-sample-test.cc:43-68: packed == non-sentinel table entries in table order; no duplicates
    N, n, V = 12000, 20, 4096
    nodes = np.random.default_rng(1).integers(0, N, size=V).astype(np.uint32)
    pool = orc.rng_pool(V * 2 * n, 3337, 54351)
    packed, table = orc.neighbor_sample(pool, nodes, N, n, 32)
    for i in range(0, V, 37):
        row = table[i]
        assert np.array_equal(row[row != N], packed[i])
        assert (row == N).sum() == 2 * n - n
        assert len(np.unique(packed[i])) == n
    assert (packed != nodes[:, None]).all()


@pytest.mark.parametrize("length", [1, 7, 32, 100, 1000, 11331])
def test_wg_sum_and_normalize_properties(orc, length):
    # wg-sum-test.cc:26-44 (exact uint sums), wg-normalize-test.cc:29-47
    for wg in (2, 4, 16, 32, 64, 96, 113):
        x = (np.arange(length, dtype=np.uint32) + 1)
        got = orc.L.orc_wg_sum_u32(x.ctypes.data_as(C.c_void_p), length, wg)
        assert got == (length * (length + 1) // 2) % 2**32
        f = x.astype(np.float32)
        s = orc.L.orc_wg_sum_f32(f.ctypes.data_as(C.c_void_p), length, wg)
        g = f.copy()
        orc.L.orc_wg_normalize_f32(g.ctypes.data_as(C.c_void_p), length, wg)
        assert np.array_equal(g, f / np.float32(s))
        assert abs(float(g.sum(dtype=np.float64)) - 1.0) < 1e-4


def test_thread_and_wg_variants_agree_loosely(orc):
    # wg-phi-test.cc:134-141 (2 %), wg-beta-test.cc:130-139, wg-perplexity-test.cc:103-107
    prob = Problem(orc, 500, 128, 20000, 4, seed=9)
    nodes = np.arange(prob.N, dtype=np.uint32)
    pool = orc.rng_pool(prob.N * 8, 56, 57)
    nb, _ = orc.neighbor_sample(pool, nodes, prob.N, 4, 32)
    a = orc.update_phi(pyoracle.MODE_THREAD, 32, prob.p_orc, prob.beta, prob.pi, prob.phi, prob.train_set,
                       nodes, nb, 1, None, True)
    b = orc.update_phi(pyoracle.MODE_WG, 32, prob.p_orc, prob.beta, prob.pi, prob.phi, prob.train_set, nodes,
                       nb, 1, None, True)
    assert rel_err(a, b).max() < 0.02


def test_c_abi_exports_every_declared_symbol():
    import re
    hdr = open(os.path.join(ROOT, "include", "ammsb.h")).read()
    names = sorted(set(re.findall(r"\b(ammsb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    so = os.path.join(ROOT, "mcmc-ammsb-gpu_b200", "libammsb.so")
    assert os.path.exists(so), "libammsb.so not built -- run __graft_entry__.build()"
    lib = C.CDLL(so)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
