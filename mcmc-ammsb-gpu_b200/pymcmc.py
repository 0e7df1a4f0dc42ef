"""ctypes binding of libmcmc.so -- the C++ host API (mcmc::Config / Learner / data /
sampling strategies), through the C wrappers of host/capi.cc.  Harness-side only."""
import ctypes as C
import os

import numpy as np

import pyammsb

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcmc.so")

STRATEGIES = ["Node", "NodeLink", "NodeNonLink", "BFLink", "BFNonLink", "BF"]
PHI_MODES = ["THREAD", "WG-NAIVE", "WG-SHARED", "WG-GEN"]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise pyammsb.AmmsbError("libmcmc.so not built: run __graft_entry__.build()")
        pyammsb.lib()  # dependency, resolved through rpath as well
        L = C.CDLL(LIB_PATH)
        L.mcmc_last_error.restype = C.c_char_p
        L.mcmc_config_create.restype = C.c_void_p
        L.mcmc_config_destroy.argtypes = [C.c_void_p]
        L.mcmc_config_set.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.mcmc_config_set_seed.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_uint64]
        L.mcmc_config_set_graph.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint]
        for f in ("num_training", "num_heldout", "max_fan_out", "max_nodes", "max_edges"):
            getattr(L, "mcmc_config_" + f).restype = C.c_uint64
            getattr(L, "mcmc_config_" + f).argtypes = [C.c_void_p]
        L.mcmc_config_get_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mcmc_config_print.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.mcmc_config_params.argtypes = [C.c_void_p, C.c_void_p]
        L.mcmc_config_set_info.restype = C.c_uint64
        L.mcmc_config_set_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.mcmc_config_set_table.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mcmc_sample.restype = C.c_float
        L.mcmc_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
        L.mcmc_host_set_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.mcmc_host_set_bins.restype = C.c_uint64
        L.mcmc_host_set_bins.argtypes = [C.c_uint64]
        L.mcmc_test_set_order.restype = C.c_uint64
        L.mcmc_test_set_order.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.mcmc_init_theta_host.argtypes = [C.c_uint32, C.c_float, C.c_float, C.c_void_p]
        L.mcmc_learner_create.restype = C.c_void_p
        L.mcmc_learner_create.argtypes = [C.c_void_p, C.c_int]
        L.mcmc_learner_destroy.argtypes = [C.c_void_p]
        L.mcmc_learner_run.argtypes = [C.c_void_p, C.c_uint32]
        L.mcmc_learner_heldout_perplexity.argtypes = [C.c_void_p, C.c_void_p]
        L.mcmc_learner_print_stats.argtypes = [C.c_void_p]
        L.mcmc_learner_training_perplexity.argtypes = [C.c_void_p, C.c_void_p]
        L.mcmc_learner_train_ppx_edges.restype = C.c_uint64
        L.mcmc_learner_train_ppx_edges.argtypes = [C.c_void_p, C.c_void_p]
        L.mcmc_learner_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.mcmc_learner_edges_processed.restype = C.c_uint64
        L.mcmc_learner_mirror_beta.argtypes = [C.c_void_p, C.c_void_p]
        L.mcmc_learner_h2d_bytes.restype = C.c_uint64
        L.mcmc_learner_h2d_bytes.argtypes = [C.c_void_p]
        L.mcmc_learner_edges_processed.argtypes = [C.c_void_p]
        L.mcmc_learner_peek.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_void_p]
        L.mcmc_learner_serialize.argtypes = [C.c_void_p, C.c_char_p]
        L.mcmc_learner_parse.argtypes = [C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise pyammsb.AmmsbError(lib().mcmc_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def host_set_build(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    bins = int(lib().mcmc_host_set_bins(len(keys)))
    table = np.zeros(8 * bins, dtype=np.uint64)
    b, p, s = C.c_uint64(0), C.c_uint32(0), C.c_uint64(0)
    ok = lib().mcmc_host_set_build(_p(keys), len(keys), _p(table), table.size, C.byref(b), C.byref(p), C.byref(s))
    return bool(ok), table, b.value, p.value, s.value


def set_order(keys, width):
    """(order of std::unordered_set, order of StdOrderSet) for one insert sequence"""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    a, b = np.zeros(len(keys), np.uint64), np.zeros(len(keys), np.uint64)
    n = lib().mcmc_test_set_order(_p(keys), len(keys), width, _p(a), _p(b))
    if n == 2 ** 64 - 1:
        raise AssertionError("StdOrderSet disagrees with std::unordered_set on membership")
    return a[:n], b[:n]


def unordered_set_order(keys):
    """iteration order of std::unordered_set<uint64_t> fed `keys` in order (the real container)"""
    return set_order(keys, 8)[0]


def init_theta_host(K, eta0=1.0, eta1=1.0):
    out = np.zeros(2 * K, dtype=np.float32)
    lib().mcmc_init_theta_host(K, eta0, eta1, _p(out))
    return out


class Config:
    """mcmc::Config with the CLI defaults of the reference's main.cc:50-70."""

    def __init__(self, cli_defaults=True, **kw):
        self.h = C.c_void_p(lib().mcmc_config_create())
        if cli_defaults:
            self.set(alpha=0)
            self.set_seed("beta_seed", 44, 45)
            self.set_seed("neighbor_seed", 56, 57)
        self.set(**kw)

    def set(self, **kw):
        for k, v in kw.items():
            if k == "strategy" and isinstance(v, str):
                v = STRATEGIES.index(v)
            if k == "phi_mode" and isinstance(v, str):
                v = PHI_MODES.index(v)
            _ck(lib().mcmc_config_set(self.h, k.encode(), float(v)))
        return self

    def set_seed(self, name, x, y):
        _ck(lib().mcmc_config_set_seed(self.h, name.encode(), x, y))

    def set_graph(self, N, edges, srand_seed=1):
        edges = np.ascontiguousarray(edges, dtype=np.uint64)
        _ck(lib().mcmc_config_set_graph(self.h, N, _p(edges), len(edges), srand_seed))
        self.N = N

    def edges(self):
        tr = np.zeros(lib().mcmc_config_num_training(self.h), dtype=np.uint64)
        he = np.zeros(lib().mcmc_config_num_heldout(self.h), dtype=np.uint64)
        lib().mcmc_config_get_edges(self.h, _p(tr), _p(he))
        return tr, he

    def max_fan_out(self):
        return int(lib().mcmc_config_max_fan_out(self.h))

    def max_nodes(self):
        return int(lib().mcmc_config_max_nodes(self.h))

    def max_edges(self):
        return int(lib().mcmc_config_max_edges(self.h))

    def params(self):
        p = pyammsb.Params()
        lib().mcmc_config_params(self.h, C.byref(p))
        return p

    def set_table(self, which):
        b, p = C.c_uint64(0), C.c_uint32(0)
        cap = lib().mcmc_config_set_info(self.h, which, C.byref(b), C.byref(p))
        t = np.zeros(cap, dtype=np.uint64)
        lib().mcmc_config_set_table(self.h, which, _p(t))
        return t, b.value, p.value

    def sample(self, strategy, seed):
        """one host mini-batch; `seed` is a ctypes c_uint that is advanced in place"""
        if isinstance(strategy, str):
            strategy = STRATEGIES.index(strategy)
        eb = np.zeros(max(self.max_edges(), 1), dtype=np.uint64)
        nb = np.zeros(max(self.max_nodes(), 2), dtype=np.uint32)
        ne, nn = C.c_uint64(0), C.c_uint64(0)
        w = lib().mcmc_sample(self.h, strategy, C.byref(seed), _p(eb), C.byref(ne), _p(nb), C.byref(nn))
        return float(w), eb[:ne.value].copy(), nb[:nn.value].copy()

    def __str__(self):
        b = C.create_string_buffer(4096)
        lib().mcmc_config_print(self.h, b, 4096)
        return b.value.decode()

    def close(self):
        if self.h:
            lib().mcmc_config_destroy(self.h)
            self.h = None


class ShardedLearner:
    """mcmc::ShardedLearner: the iteration over several GPUs of one box (column-sharded pi); ranks that
    share a device are computed by one launch, so devices=[0, 0] runs the two-rank protocol on one GPU"""

    def __init__(self, cfg, devices):
        self.cfg, self.world = cfg, len(devices)
        lib().mcmc_sharded_create.restype = C.c_void_p
        lib().mcmc_sharded_edges_processed.restype = C.c_uint64
        d = (C.c_int * len(devices))(*devices)
        h = lib().mcmc_sharded_create(cfg.h, d, len(devices))
        if not h:
            raise pyammsb.AmmsbError(lib().mcmc_last_error().decode())
        self.h = C.c_void_p(h)

    def run(self, iters):
        _ck(lib().mcmc_sharded_run(self.h, iters))

    def heldout_perplexity(self):
        out = C.c_float(0)
        _ck(lib().mcmc_sharded_heldout_perplexity(self.h, C.byref(out)))
        return out.value

    def edges_processed(self):
        return int(lib().mcmc_sharded_edges_processed(self.h))

    def read(self, N, K):
        """pi [N, K], phi [N], beta [world, 2K], theta [world, 2K] (every rank's copy)"""
        pi, phi = np.zeros((N, K), np.float32), np.zeros(N, np.float32)
        beta, theta = np.zeros((self.world, 2 * K), np.float32), np.zeros((self.world, 2 * K), np.float32)
        _ck(lib().mcmc_sharded_read(self.h, _p(pi), _p(phi), _p(beta), _p(theta), C.c_uint64(N), C.c_uint64(K)))
        return pi, phi, beta, theta

    def close(self):
        if self.h is not None:
            lib().mcmc_sharded_destroy(self.h)
            self.h = None


class Learner:
    def __init__(self, cfg, device=0):
        self.cfg = cfg
        h = lib().mcmc_learner_create(cfg.h, device)
        if not h:
            raise pyammsb.AmmsbError(lib().mcmc_last_error().decode())
        self.h = C.c_void_p(h)

    def run(self, iters):
        _ck(lib().mcmc_learner_run(self.h, iters))

    def heldout_perplexity(self):
        out = C.c_float(0)
        _ck(lib().mcmc_learner_heldout_perplexity(self.h, C.byref(out)))
        return out.value

    def training_perplexity(self):
        out = C.c_float(0)
        _ck(lib().mcmc_learner_training_perplexity(self.h, C.byref(out)))
        return out.value

    def training_perplexity_edges(self):
        n = lib().mcmc_learner_train_ppx_edges(self.h, None)
        e = np.zeros(n, dtype=np.uint64)
        lib().mcmc_learner_train_ppx_edges(self.h, _p(e))
        return e

    def print_stats(self):
        _ck(lib().mcmc_learner_print_stats(self.h))

    def read(self, N, K, pi=True):
        pi_a = np.zeros((N, K), dtype=np.float32) if pi else None
        phi, beta, theta = np.zeros(N, np.float32), np.zeros(2 * K, np.float32), np.zeros(2 * K, np.float32)
        _ck(lib().mcmc_learner_read(self.h, _p(pi_a) if pi else None, _p(phi), _p(beta), _p(theta), N))
        return pi_a, phi, beta, theta

    def read_beta(self, K, out=None):
        """device -> host read of beta [2K] (the per-iteration result a caller polls)"""
        beta = np.empty(2 * K, np.float32) if out is None else out
        _ck(lib().mcmc_learner_read(self.h, None, None, _p(beta), None, 0))
        return beta

    def mirror_beta(self, pinned_ptr):
        """every iteration of run() ends with a D2H copy of beta[2K] into this pinned buffer"""
        _ck(lib().mcmc_learner_mirror_beta(self.h, C.c_void_p(pinned_ptr)))

    def h2d_bytes(self):
        return int(lib().mcmc_learner_h2d_bytes(self.h))

    def edges_processed(self):
        return int(lib().mcmc_learner_edges_processed(self.h))

    def peek(self, n):
        eb = np.zeros(self.cfg.max_edges(), dtype=np.uint64)
        nb = np.zeros(self.cfg.max_nodes(), dtype=np.uint32)
        nbr = np.zeros(self.cfg.max_nodes() * n, dtype=np.uint32)
        ne, nn = C.c_uint64(0), C.c_uint64(0)
        w = C.c_float(0)
        _ck(lib().mcmc_learner_peek(self.h, _p(eb), C.byref(ne), _p(nb), C.byref(nn), _p(nbr), n, C.byref(w)))
        return eb[:ne.value].copy(), nb[:nn.value].copy(), nbr[:nn.value * n].reshape(-1, n).copy(), w.value

    def serialize(self, path):
        _ck(lib().mcmc_learner_serialize(self.h, path.encode()))

    def parse(self, path):
        _ck(lib().mcmc_learner_parse(self.h, path.encode()))

    def close(self):
        if self.h:
            lib().mcmc_learner_destroy(self.h)
            self.h = None
