"""Shared synthetic-problem helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                      "libref_oracle.so")


class RefSampler:
    """The reference's own data.cc split and sample.cc strategies (compiled in place into
    oracle/_ref; test infrastructure): GenerateSetsFromEdges on `keys`, then mini-batches."""

    STRATEGIES = ["Node", "NodeLink", "NodeNonLink", "BFLink", "BFNonLink", "BF"]

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, N, keys, heldout_ratio, srand_seed, m):
        L = self.L = C.CDLL(REF_SO)
        L.ref_generate_sets.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_double, C.c_uint, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_sampler_create.restype = C.c_void_p
        L.ref_sampler_create.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                         C.c_uint64]
        L.ref_sampler_destroy.argtypes = [C.c_void_p]
        L.ref_sample.restype = C.c_float
        L.ref_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        E = len(keys)
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        tr, he = np.zeros(E, dtype=np.uint64), np.zeros(E, dtype=np.uint64)
        ntr, nhe = C.c_uint64(0), C.c_uint64(0)
        if not L.ref_generate_sets(N, vp(keys), E, heldout_ratio, srand_seed, vp(tr), C.byref(ntr), vp(he),
                                   C.byref(nhe)):
            raise RuntimeError("reference GenerateSetsFromEdges failed")
        self.training, self.heldout = tr[:ntr.value].copy(), he[:nhe.value].copy()
        self.m = m
        self.h = L.ref_sampler_create(N, E, vp(self.training), len(self.training), vp(self.heldout),
                                      E - len(self.training), m)

    def sample(self, strategy, seed):
        """(weight, edges, nodes) of the reference strategy; `seed` (c_uint) advances in place"""
        s = self.STRATEGIES.index(strategy)
        eb = np.zeros(8 * self.m + 65536, dtype=np.uint64)
        nb = np.zeros(16 * self.m + 131072, dtype=np.uint32)
        ne, nn = C.c_uint64(0), C.c_uint64(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        w = self.L.ref_sample(self.h, s, C.byref(seed), vp(eb), C.byref(ne), vp(nb), C.byref(nn))
        return float(w), eb[:ne.value].copy(), nb[:nn.value].copy()

    def close(self):
        if self.h:
            self.L.ref_sampler_destroy(self.h)
            self.h = None


def make_edges(N, E, seed):
    """E unique undirected edges u<v as canonical keys (types.h:66-74), shuffled."""
    rng = np.random.default_rng(seed)
    keys = np.zeros(0, dtype=np.uint64)
    while len(keys) < E:
        u = rng.integers(0, N, size=2 * E, dtype=np.uint64)
        v = rng.integers(0, N, size=2 * E, dtype=np.uint64)
        m = u != v
        lo, hi = np.minimum(u[m], v[m]), np.maximum(u[m], v[m])
        keys = np.unique(np.concatenate([keys, (lo << np.uint64(32)) | hi]))
    keys = keys[rng.permutation(len(keys))[:E]]
    return keys


def split_edges(keys, heldout_ratio):
    """data.cc:86-99: held-out links = first E - ceil((1 - r/2) E) entries."""
    E = len(keys)
    training_len = int(np.ceil((1 - heldout_ratio / 2) * E))
    h = E - training_len
    return keys[h:], keys[:h]


def fake_nonlinks(N, count, forbidden, seed):
    rng = np.random.default_rng(seed)
    forb = set(int(x) for x in forbidden)
    out = []
    while len(out) < count:
        u, v = int(rng.integers(0, N)), int(rng.integers(0, N))
        if u == v:
            continue
        k = (min(u, v) << 32) | max(u, v)
        if k in forb:
            continue
        forb.add(k)
        out.append(k)
    return np.array(out, dtype=np.uint64)


def random_pi(N, K, seed):
    rng = np.random.default_rng(seed)
    g = rng.gamma(1.0, 1.0, size=(N, K)).astype(np.float32)
    phi = g.sum(axis=1, dtype=np.float32)
    pi = (g / phi[:, None]).astype(np.float32)
    return np.ascontiguousarray(pi), np.ascontiguousarray(phi)


def random_theta(K, seed):
    rng = np.random.default_rng(seed)
    return rng.gamma(1.0, 1.0, size=2 * K).astype(np.float32) + np.float32(0.05)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


class Problem:
    """A small a-MMSB state shared by oracle and device: graph, sets, pi/phi, theta/beta."""

    def __init__(self, orc, N, K, E, n, seed=1, heldout_ratio=0.1, **hyper):
        self.N, self.K, self.E, self.n = N, K, E, n
        keys = make_edges(N, E, seed)
        self.train_edges, self.heldout_links = split_edges(keys, heldout_ratio)
        # the reference's two hashes are correlated for some table sizes (cuckoo.cc:199-209:
        # both depend on k mod gcd-factors of N_), so a build can legitimately fail; shrink the
        # held-out part until both sets build, as a user of the reference would re-split
        for _ in range(64):
            try:
                self.train_set = orc.set_build(self.train_edges)
                self.heldout_set = orc.set_build(self.heldout_links)
                break
            except RuntimeError:
                self.train_edges = np.concatenate([self.train_edges, self.heldout_links[-1:]])
                self.heldout_links = self.heldout_links[:-1]
        else:
            raise RuntimeError("could not build cuckoo sets")
        self.heldout_nonlinks = fake_nonlinks(N, len(self.heldout_links), keys, seed + 1)
        self.heldout_edges = np.concatenate([self.heldout_links, self.heldout_nonlinks])
        self.pi, self.phi = random_pi(N, K, seed + 2)
        self.theta = random_theta(K, seed + 3)
        self.beta = orc.theta_to_beta(self.theta)
        self.p_orc = orc.make_params(N, E, K, n, **hyper)

    def minibatch_nodes(self, V, seed):
        rng = np.random.default_rng(seed)
        return rng.permutation(self.N)[:V].astype(np.uint32)

    def minibatch_edges(self, m, seed, link_fraction=0.5):
        """mixed link / non-link mini-batch (canonical keys)"""
        rng = np.random.default_rng(seed)
        nl = int(m * link_fraction)
        links = self.train_edges[rng.permutation(len(self.train_edges))[:nl]]
        non = fake_nonlinks(self.N, m - nl, np.concatenate([self.train_edges, self.heldout_links]),
                            seed + 7)
        e = np.concatenate([links, non])
        return e[rng.permutation(len(e))]
