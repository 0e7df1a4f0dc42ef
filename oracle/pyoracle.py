"""ctypes binding of the CPU oracle (oracle/liboracle*.so).  TEST INFRASTRUCTURE ONLY:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MODE_THREAD, MODE_WG = 0, 1


class RngState(C.Structure):
    _fields_ = [("x", C.c_uint64), ("y", C.c_uint64)]


class Params(C.Structure):
    _fields_ = [("N", C.c_uint64), ("E", C.c_uint64), ("K", C.c_uint32),
                ("num_neighbors", C.c_uint32), ("alpha", C.c_float), ("a", C.c_float),
                ("b", C.c_float), ("c", C.c_float), ("epsilon", C.c_float),
                ("eta0", C.c_float), ("eta1", C.c_float)]


class SetStruct(C.Structure):
    _fields_ = [("table", C.c_void_p), ("num_bins", C.c_uint64), ("prime_idx", C.c_uint32),
                ("count", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One loaded oracle library (single-threaded checker by default, OpenMP build for the
    cpu_baseline)."""

    def __init__(self, omp=False, path=None):
        if path is None:
            path = os.path.join(_HERE, "liboracle_omp.so" if omp else "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.orc_rand.restype = C.c_uint64
        L.orc_random.restype = C.c_float
        L.orc_randn.restype = C.c_float
        L.orc_rand_gamma.restype = C.c_float
        L.orc_rand_gamma.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_round_param.restype = C.c_float
        L.orc_round_param.argtypes = [C.c_float]
        L.orc_eps_t.restype = C.c_float
        L.orc_eps_t.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_set_bins_for.restype = C.c_uint64
        L.orc_set_bins_for.argtypes = [C.c_uint64]
        L.orc_set_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_set_has.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_set_has_many.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_rng_init.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_neighbor_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_update_phi.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                     C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_update_pi.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_update_beta.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_float,
                                      C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_perplexity.restype = C.c_double
        L.orc_perplexity.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                     C.c_void_p]
        L.orc_init_pi.argtypes = [C.c_uint64, C.c_uint32, C.c_float, C.c_float, C.c_void_p,
                                  C.c_void_p]
        L.orc_theta_to_beta.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_wg_sum_f32.restype = C.c_float
        L.orc_wg_sum_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_wg_sum_u32.restype = C.c_uint32
        L.orc_wg_sum_u32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_wg_normalize_f32.restype = C.c_float
        L.orc_wg_normalize_f32.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]

    # ---- params ----
    def round_param(self, f):
        return float(self.L.orc_round_param(f))

    def make_params(self, N, E, K, n, alpha=None, a=0.0315, b=1024.0, c=0.5, epsilon=1e-7,
                    eta0=1.0, eta1=1.0):
        if alpha is None or alpha == 0:
            alpha = float(np.float32(1.0) / np.float32(K))  # main.cc:153
        r = self.round_param
        return Params(N, E, K, n, r(alpha), r(a), r(b), r(c), r(epsilon), r(eta0), r(eta1))

    def eps_t(self, p, step):
        return float(self.L.orc_eps_t(C.byref(p), step))

    # ---- rng ----
    def rng_pool(self, n, sx, sy):
        pool = np.zeros((n, 2), dtype=np.uint64)
        self.L.orc_rng_init(_p(pool), n, sx, sy)
        return pool

    def draw_u64(self, pool, draws):
        out = np.zeros((pool.shape[0], draws), dtype=np.uint64)
        for i in range(pool.shape[0]):
            st = C.c_void_p(pool.ctypes.data + 16 * i)
            for d in range(draws):
                out[i, d] = self.L.orc_rand(st)
        return out

    def draw_randn(self, pool, draws):
        out = np.zeros((pool.shape[0], draws), dtype=np.float32)
        for i in range(pool.shape[0]):
            st = C.c_void_p(pool.ctypes.data + 16 * i)
            for d in range(draws):
                out[i, d] = self.L.orc_randn(st)
        return out

    def draw_gamma(self, pool, draws, a, b):
        out = np.zeros((pool.shape[0], draws), dtype=np.float32)
        for i in range(pool.shape[0]):
            st = C.c_void_p(pool.ctypes.data + 16 * i)
            for d in range(draws):
                out[i, d] = self.L.orc_rand_gamma(st, a, b)
        return out

    # ---- cuckoo ----
    def set_build(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        s = SetStruct()
        ok = self.L.orc_set_build(_p(keys), len(keys), C.byref(s))
        if not ok:
            raise RuntimeError("cuckoo build failed")
        return OracleSet(self, s)

    # ---- operators ----
    def neighbor_sample(self, pool, nodes, N, n, wg=32):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        V = len(nodes)
        h = np.zeros((V, 2 * n), dtype=np.uint32)
        out = np.zeros((V, n), dtype=np.uint32)
        self.L.orc_neighbor_sample(_p(pool), _p(nodes), V, N, n, wg, _p(h), _p(out))
        return out, h

    def update_phi(self, mode, wg, p, beta, pi, phi, train, nodes, neighbors, step, pool,
                   disable_noise=False):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        neighbors = np.ascontiguousarray(neighbors, dtype=np.uint32)
        V = len(nodes)
        phi_vec = np.zeros((V, p.K), dtype=np.float32)
        self.L.orc_update_phi(mode, wg, C.byref(p), _p(beta), _p(pi), _p(phi), C.byref(train.s),
                              _p(nodes), _p(neighbors), V, step,
                              _p(pool) if pool is not None else None, int(disable_noise),
                              _p(phi_vec))
        return phi_vec

    def update_pi(self, mode, wg, K, pi, phi, phi_vec, nodes):
        nodes = np.ascontiguousarray(nodes, dtype=np.uint32)
        self.L.orc_update_pi(mode, wg, K, _p(pi), _p(phi), _p(phi_vec), _p(nodes), len(nodes))

    def update_beta(self, mode, wg, p, theta, beta, pi, train, edges, scale, step, pool):
        edges = np.ascontiguousarray(edges, dtype=np.uint64)
        theta_sum = np.zeros(p.K, dtype=np.float32)
        grads = np.zeros(2 * p.K, dtype=np.float32)
        self.L.orc_update_beta(mode, wg, C.byref(p), _p(theta), _p(beta), _p(pi),
                               C.byref(train.s), _p(edges), len(edges), scale, step, _p(pool),
                               _p(theta_sum), _p(grads))
        return theta_sum, grads

    def perplexity(self, mode, wg, p, pi, beta, heldout, edges, ppx_per_edge, call_count):
        edges = np.ascontiguousarray(edges, dtype=np.uint64)
        sums = np.zeros(4, dtype=np.float64)
        avg = self.L.orc_perplexity(mode, wg, C.byref(p), _p(pi), _p(beta), C.byref(heldout.s),
                                    _p(edges), len(edges), _p(ppx_per_edge), call_count, _p(sums))
        return float(avg), sums

    def init_pi(self, N, K, eta0=1.0, eta1=1.0):
        pi = np.zeros((N, K), dtype=np.float32)
        phi = np.zeros(N, dtype=np.float32)
        self.L.orc_init_pi(N, K, eta0, eta1, _p(pi), _p(phi))
        return pi, phi

    def theta_to_beta(self, theta):
        beta = np.zeros_like(theta)
        self.L.orc_theta_to_beta(len(theta) // 2, _p(theta), _p(beta))
        return beta


class OracleSet:
    def __init__(self, orc, s):
        self.orc, self.s = orc, s
        self.num_bins = int(s.num_bins)
        self.prime_idx = int(s.prime_idx)
        self.count = int(s.count)

    def table(self):
        n = 2 * 4 * self.num_bins
        return np.ctypeslib.as_array(C.cast(self.s.table, C.POINTER(C.c_uint64)), shape=(n,)).copy()

    def has(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(len(keys), dtype=np.uint8)
        self.orc.L.orc_set_has_many(C.byref(self.s), _p(keys), len(keys), _p(out))
        return out

    def __del__(self):
        try:
            self.orc.L.orc_set_free(C.byref(self.s))
        except Exception:
            pass
