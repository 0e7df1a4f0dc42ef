"""ctypes binding of libammsb.so (the C ABI in include/ammsb.h).

Harness-side only: tests, bench.py and __graft_entry__ drive the CUDA path through
this module.  It fails loudly when the library is missing or no GPU is present --
there is no CPU fallback anywhere in the product path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libammsb.so")

MODE_THREAD, MODE_WG = 0, 1


class Params(C.Structure):
    _fields_ = [("N", C.c_uint64), ("E", C.c_uint64), ("K", C.c_uint32),
                ("num_neighbors", C.c_uint32), ("alpha", C.c_float), ("a", C.c_float),
                ("b", C.c_float), ("c", C.c_float), ("epsilon", C.c_float),
                ("eta0", C.c_float), ("eta1", C.c_float)]


class PhiOpts(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("wg", C.c_uint32), ("disable_noise", C.c_uint32),
                ("strict", C.c_uint32), ("part_index", C.c_uint32), ("part_count", C.c_uint32)]


class AmmsbError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AmmsbError("libammsb.so not built: run __graft_entry__.build() "
                             "(make -C mcmc-ammsb-gpu_b200/csrc)")
        L = C.CDLL(LIB_PATH)
        L.ammsb_last_error.restype = C.c_char_p
        L.ammsb_version.restype = C.c_char_p
        L.ammsb_round_param.restype = C.c_float
        L.ammsb_round_param.argtypes = [C.c_float]
        L.ammsb_eps_t.restype = C.c_float
        L.ammsb_eps_t.argtypes = [C.c_void_p, C.c_uint32]
        # byte counts are size_t: without a prototype ctypes passes a Python int as a 32-bit int
        L.ammsb_malloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.ammsb_free.argtypes = [C.c_void_p, C.c_void_p]
        L.ammsb_memset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
        for f in (L.ammsb_h2d, L.ammsb_d2h, L.ammsb_d2d, L.ammsb_h2d_async, L.ammsb_d2h_async):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise AmmsbError(lib().ammsb_last_error().decode())


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return a


def round_param(f):
    return float(lib().ammsb_round_param(f))


def make_params(N, E, K, n, alpha=None, a=0.0315, b=1024.0, c=0.5, epsilon=1e-7, eta0=1.0,
                eta1=1.0):
    if alpha is None or alpha == 0:
        alpha = float(np.float32(1.0) / np.float32(K))  # main.cc:153
    r = round_param
    return Params(N, E, K, n, r(alpha), r(a), r(b), r(c), r(epsilon), r(eta0), r(eta1))


def device_count():
    n = C.c_int(0)
    rc = lib().ammsb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def launch_count():
    n = C.c_uint64(0)
    lib().ammsb_launch_count(C.byref(n))
    return n.value


class DevBuf:
    """clcuda::Buffer<T> stand-in: a typed device allocation."""

    def __init__(self, ctx, dtype, count):
        self.ctx, self.dtype, self.count = ctx, np.dtype(dtype), int(count)
        self.nbytes = self.dtype.itemsize * self.count
        ptr = C.c_void_p()
        _ck(lib().ammsb_malloc(ctx.h, max(self.nbytes, 16), C.byref(ptr)))
        self.ptr = ptr

    def write(self, arr, offset=0):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert offset + arr.size <= self.count
        _ck(lib().ammsb_h2d(self.ctx.h, C.c_void_p(self.ptr.value + offset * self.dtype.itemsize),
                            _p(arr), arr.nbytes))
        return self

    def read(self, count=None, offset=0):
        count = self.count - offset if count is None else count
        out = np.empty(count, dtype=self.dtype)
        _ck(lib().ammsb_d2h(self.ctx.h, _p(out),
                            C.c_void_p(self.ptr.value + offset * self.dtype.itemsize), out.nbytes))
        return out

    def zero(self):
        _ck(lib().ammsb_memset(self.ctx.h, self.ptr, 0, self.nbytes))
        return self

    def free(self):
        if self.ptr is not None:
            lib().ammsb_free(self.ctx.h, self.ptr)
            self.ptr = None


class Ctx:
    def __init__(self, device=0):
        h = C.c_void_p()
        _ck(lib().ammsb_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def sync(self):
        _ck(lib().ammsb_ctx_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        _ck(lib().ammsb_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sm_count(self):
        n = C.c_int(0)
        _ck(lib().ammsb_ctx_sm_count(self.h, C.byref(n)))
        return n.value

    def name(self):
        b = C.create_string_buffer(256)
        _ck(lib().ammsb_ctx_device_name(self.h, b, 256))
        return b.value.decode()

    def buf(self, dtype, count):
        return DevBuf(self, dtype, count)

    def from_host(self, arr):
        arr = np.ascontiguousarray(arr)
        return DevBuf(self, arr.dtype, arr.size).write(arr.ravel())

    def timer_start(self):
        _ck(lib().ammsb_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float(0)
        _ck(lib().ammsb_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    # ---- operators ----
    def neighbor_sample(self, pool, d_nodes, V, N, n, wg, d_neighbors, d_hash=None):
        _ck(lib().ammsb_neighbor_sample(self.h, pool.h, d_nodes.ptr, V, N, n, wg, d_neighbors.ptr,
                                        d_hash.ptr if d_hash is not None else None))

    def update_phi(self, p, opts, d_beta, store, train, d_nodes, d_neighbors, V, step, pool,
                   d_phi_vec, d_phi_sum):
        _ck(lib().ammsb_update_phi(self.h, C.byref(p), C.byref(opts), d_beta.ptr, store.h, train.h,
                                   d_nodes.ptr, d_neighbors.ptr, V, step,
                                   pool.h if pool is not None else None, d_phi_vec.ptr,
                                   d_phi_sum.ptr))

    def update_pi(self, K, store, d_phi_vec, d_phi_sum, d_nodes, V):
        _ck(lib().ammsb_update_pi(self.h, K, store.h, d_phi_vec.ptr,
                                  d_phi_sum.ptr if d_phi_sum is not None else None, d_nodes.ptr, V))

    def update_pi_part(self, K, store, d_phi_vec, d_phi_sum, d_nodes, V, opts):
        _ck(lib().ammsb_update_pi_part(self.h, K, store.h, d_phi_vec.ptr,
                                       d_phi_sum.ptr if d_phi_sum is not None else None, d_nodes.ptr, V,
                                       C.byref(opts)))

    def beta_workspace_bytes(self, K):
        n = C.c_size_t(0)
        _ck(lib().ammsb_beta_workspace_bytes(self.h, K, C.byref(n)))
        return n.value

    def beta_grads(self, p, d_theta, d_beta, store, train, d_edges, E_mb, d_theta_sum, d_grads, ws):
        _ck(lib().ammsb_beta_grads(self.h, C.byref(p), d_theta.ptr, d_beta.ptr, store.h, train.h,
                                   d_edges.ptr, E_mb, d_theta_sum.ptr, d_grads.ptr, ws.ptr,
                                   C.c_size_t(ws.nbytes)))

    def update_theta(self, p, d_theta, d_beta, d_grads, scale, step, pool):
        _ck(lib().ammsb_update_theta(self.h, C.byref(p), d_theta.ptr, d_beta.ptr, d_grads.ptr,
                                     C.c_float(scale), step, pool.h))

    def update_beta(self, p, d_theta, d_beta, store, train, d_edges, E_mb, scale, step, pool,
                    d_theta_sum, d_grads, ws):
        _ck(lib().ammsb_update_beta(self.h, C.byref(p), d_theta.ptr, d_beta.ptr, store.h, train.h,
                                    d_edges.ptr, E_mb, C.c_float(scale), step, pool.h,
                                    d_theta_sum.ptr, d_grads.ptr, ws.ptr, C.c_size_t(ws.nbytes)))

    def perplexity_workspace_bytes(self):
        n = C.c_size_t(0)
        _ck(lib().ammsb_perplexity_workspace_bytes(self.h, C.byref(n)))
        return n.value

    def perplexity(self, p, store, d_beta, heldout, d_edges, H, d_ppx, call_count, ws):
        sums = (C.c_double * 4)()
        avg = C.c_double(0)
        _ck(lib().ammsb_perplexity(self.h, C.byref(p), store.h, d_beta.ptr, heldout.h, d_edges.ptr,
                                   H, d_ppx.ptr, call_count, sums, C.byref(avg), ws.ptr,
                                   C.c_size_t(ws.nbytes)))
        return avg.value, np.array(list(sums))

    def perplexity_partial(self, p, store, d_beta, heldout, d_edges, H, d_ppx, call_count, d_sums,
                           ws):
        _ck(lib().ammsb_perplexity_partial(self.h, C.byref(p), store.h, d_beta.ptr, heldout.h,
                                           d_edges.ptr, H, d_ppx.ptr, call_count, d_sums.ptr,
                                           ws.ptr, C.c_size_t(ws.nbytes)))

    def row_sum(self, d_in, rows, length, d_out):
        _ck(lib().ammsb_row_sum(self.h, d_in.ptr, rows, length, d_out.ptr))

    def row_normalize(self, d_inout, rows, length, d_sum):
        _ck(lib().ammsb_row_normalize(self.h, d_inout.ptr, rows, length,
                                      d_sum.ptr if d_sum is not None else None))

    def close(self):
        if self.h is not None:
            lib().ammsb_ctx_destroy(self.h)
            self.h = None


class Rng:
    def __init__(self, ctx, n, sx, sy):
        h = C.c_void_p()
        _ck(lib().ammsb_rng_create(ctx.h, C.c_uint64(n), C.c_uint64(sx), C.c_uint64(sy), C.byref(h)))
        self.h, self.n, self.ctx = h, n, ctx

    def get_state(self):
        out = np.empty((self.n, 2), dtype=np.uint64)
        _ck(lib().ammsb_rng_get_state(self.h, _p(out)))
        return out

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.uint64)
        assert st.shape == (self.n, 2)
        _ck(lib().ammsb_rng_set_state(self.h, _p(st)))

    def draw_u64(self, draws):
        out = np.empty((self.n, draws), dtype=np.uint64)
        _ck(lib().ammsb_rng_draw_u64(self.h, draws, _p(out)))
        return out

    def draw_randn(self, draws):
        out = np.empty((self.n, draws), dtype=np.float32)
        _ck(lib().ammsb_rng_draw_randn(self.h, draws, _p(out)))
        return out

    def draw_gamma(self, draws, a, b):
        out = np.empty((self.n, draws), dtype=np.float32)
        _ck(lib().ammsb_rng_draw_gamma(self.h, draws, C.c_float(a), C.c_float(b), _p(out)))
        return out

    def free(self):
        if self.h is not None:
            lib().ammsb_rng_destroy(self.h)
            self.h = None


class DevSet:
    def __init__(self, ctx, table, num_bins, prime_idx):
        table = np.ascontiguousarray(table, dtype=np.uint64)
        assert table.size == 8 * num_bins
        h = C.c_void_p()
        _ck(lib().ammsb_set_create(ctx.h, _p(table), C.c_uint64(num_bins), prime_idx, C.byref(h)))
        self.h, self.ctx = h, ctx

    def has(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(len(keys), dtype=np.uint8)
        _ck(lib().ammsb_set_has(self.h, _p(keys), C.c_uint64(len(keys)), _p(out)))
        return out

    def free(self):
        if self.h is not None:
            lib().ammsb_set_destroy(self.h)
            self.h = None


class BuiltSet(DevSet):
    """cuckoo set built on the device from a key list (host array or device buffer)"""

    def __init__(self, ctx, keys, n=None):
        h = C.c_void_p()
        if isinstance(keys, np.ndarray):
            keys = np.ascontiguousarray(keys, dtype=np.uint64)
            _ck(lib().ammsb_set_build(ctx.h, _p(keys), C.c_uint64(len(keys)), C.byref(h)))
        else:
            _ck(lib().ammsb_set_build_device(ctx.h, keys.ptr, C.c_uint64(n), C.byref(h)))
        self.h, self.ctx = h, ctx
        bins, prime = C.c_uint64(0), C.c_uint32(0)
        _ck(lib().ammsb_set_info(h, C.byref(bins), C.byref(prime)))
        self.num_bins, self.prime_idx = bins.value, prime.value

    def table(self):
        out = np.empty(8 * self.num_bins, dtype=np.uint64)
        _ck(lib().ammsb_set_read_table(self.h, _p(out)))
        return out


def graph_generate(ctx, N, E, seed, d_edges):
    _ck(lib().ammsb_graph_generate(ctx.h, C.c_uint64(N), C.c_uint64(E), C.c_uint64(seed), d_edges.ptr))


def graph_nonlinks(ctx, N, count, seed, set_a, set_b, d_out):
    _ck(lib().ammsb_graph_nonlinks(ctx.h, C.c_uint64(N), C.c_uint64(count), C.c_uint64(seed), set_a.h,
                                   set_b.h if set_b is not None else None, d_out.ptr))


def graph_csr(ctx, N, d_edges, E, d_offsets, d_adj, d_degree=None):
    _ck(lib().ammsb_graph_csr(ctx.h, C.c_uint64(N), d_edges.ptr, C.c_uint64(E), d_offsets.ptr, d_adj.ptr,
                              d_degree.ptr if d_degree is not None else None))


_libc = None


def rand_r(seed):
    """glibc rand_r on a ctypes c_uint (advanced in place)"""
    global _libc
    if _libc is None:
        _libc = C.CDLL(None)
        _libc.rand_r.argtypes = [C.POINTER(C.c_uint)]
        _libc.rand_r.restype = C.c_int
    return _libc.rand_r(C.byref(seed))


class DeviceSampler:
    """sampleNode (sample.cc:295-302) with the mini-batch produced on the device: the host draws
    the coin and the vertex with rand_r exactly as the strategy does, the device examines the
    candidate stream.  Same edges and nodes as the host strategy for the same seed (in draw
    order), same seed afterwards."""

    def __init__(self, ctx, N, E, m, train, heldout, d_offsets, d_adj, degree, exact_order=False):
        """exact_order: emit edges and nodes in the reference's std::unordered_set order
        (csrc/orderset.cu); with d_adj in the host Graph's order the mini-batch is then bit-identical
        to the host strategy's"""
        h = C.c_void_p()
        _ck(lib().ammsb_sampler_create(ctx.h, C.c_uint64(N), m, C.byref(h)))
        self.h, self.ctx, self.N, self.E, self.m = h, ctx, N, E, m
        self.train, self.heldout, self.d_offsets, self.d_adj, self.degree = train, heldout, d_offsets, d_adj, degree
        self.os = None
        if exact_order:
            o = C.c_void_p()
            _ck(lib().ammsb_orderset_create(ctx.h, 2 * max(m, int(np.max(degree)) if len(degree) else 1) + 2, C.byref(o)))
            self.os = o

    def _finish(self, ctx, d_edges, d_nodes, E_mb):
        nn = C.c_uint32(0)
        _ck(lib().ammsb_minibatch_finish(self.os, ctx.h, d_edges.ptr, E_mb, d_nodes.ptr, C.byref(nn)))
        return nn.value

    def sample(self, seed, d_edges, d_nodes, ctx=None):
        """one mini-batch into d_edges / d_nodes; returns (weight, E_mb, V)"""
        ctx = ctx or self.ctx
        if rand_r(seed) % 2:  # sampleNodeLink: unseen vertices until one has training neighbors
            while True:
                u = rand_r(seed) % self.N
                d = int(self.degree[u])
                if d > 0:
                    break
            _ck(lib().ammsb_minibatch_link(ctx.h, u, d, self.d_offsets.ptr, self.d_adj.ptr, d_edges.ptr, d_nodes.ptr))
            if self.os is not None:
                return float(np.float32(self.N)), d, self._finish(ctx, d_edges, d_nodes, d)
            return float(np.float32(self.N)), d, d + 1
        u = rand_r(seed) % self.N
        ne, nn = C.c_uint32(0), C.c_uint32(0)
        _ck(lib().ammsb_minibatch_nonlink(self.h, ctx.h, u, C.byref(seed), self.train.h,
                                          self.heldout.h if self.heldout is not None else None, d_edges.ptr,
                                          d_nodes.ptr, C.byref(ne), C.byref(nn)))
        if self.os is not None:
            nn = C.c_uint32(self._finish(ctx, d_edges, d_nodes, ne.value))
        return float(np.float32(2 * self.E) / np.float32(self.m)), ne.value, nn.value

    def free(self):
        if self.h is not None:
            lib().ammsb_sampler_destroy(self.h)
            self.h = None
        if self.os is not None:
            lib().ammsb_orderset_destroy(self.os)
            self.os = None


class Peer:
    """cross-GPU barrier / rank-ordered all-reduce over NVLink peer memory (csrc/peer.cu)"""

    def __init__(self, ctx, world, rank, slot_bytes):
        h = C.c_void_p()
        _ck(lib().ammsb_peer_create(ctx.h, world, rank, C.c_size_t(slot_bytes), C.byref(h)))
        self.h, self.ctx, self.world, self.rank = h, ctx, world, rank

    def export_fd(self):
        fd = C.c_int(-1)
        _ck(lib().ammsb_peer_export_fd(self.h, C.byref(fd)))
        return fd.value

    def attach_fd(self, peer_rank, fd):
        _ck(lib().ammsb_peer_attach_fd(self.h, peer_rank, fd))

    def barrier(self, ctx=None):
        _ck(lib().ammsb_peer_barrier((ctx or self.ctx).h, self.h))

    def allreduce_f32(self, d_buf, count, ctx=None):
        _ck(lib().ammsb_peer_allreduce_f32((ctx or self.ctx).h, self.h, d_buf.ptr, count))

    def allreduce_f64(self, d_buf, count, ctx=None):
        _ck(lib().ammsb_peer_allreduce_f64((ctx or self.ctx).h, self.h, d_buf.ptr, count))

    def check(self):
        """raise if a wait inside an exchange kernel gave up (a peer died or never launched)"""
        ep = C.c_uint32(0)
        _ck(lib().ammsb_peer_check(self.h, C.byref(ep)))
        if ep.value:
            raise AmmsbError("rank %d: peer exchange number %d timed out" % (self.rank, ep.value))

    def free(self):
        if self.h is not None:
            lib().ammsb_peer_destroy(self.h)
            self.h = None


class Store:
    def __init__(self, ctx, N, K, num_shards=1, shard_id=0, shareable=False):
        h = C.c_void_p()
        create = lib().ammsb_store_create_shareable if shareable else lib().ammsb_store_create
        _ck(create(ctx.h, C.c_uint64(N), K, num_shards, shard_id, C.byref(h)))
        self.h, self.ctx, self.N, self.K = h, ctx, N, K
        self.num_shards, self.shard_id = num_shards, shard_id
        a, b = C.c_uint64(0), C.c_uint64(0)
        _ck(lib().ammsb_store_rows(h, C.byref(a), C.byref(b)))
        self.first_row, self.local_rows = a.value, b.value

    def init_pi(self, eta0=1.0, eta1=1.0):
        _ck(lib().ammsb_store_init_pi(self.h, C.c_float(eta0), C.c_float(eta1)))

    def write_pi(self, arr, row0=None):
        row0 = self.first_row if row0 is None else row0
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1, self.K)
        _ck(lib().ammsb_store_write_pi(self.h, C.c_uint64(row0), C.c_uint64(arr.shape[0]), _p(arr)))

    def read_pi(self, row0=None, nrows=None):
        row0 = self.first_row if row0 is None else row0
        nrows = self.local_rows if nrows is None else nrows
        out = np.empty((nrows, self.K), dtype=np.float32)
        _ck(lib().ammsb_store_read_pi(self.h, C.c_uint64(row0), C.c_uint64(nrows), _p(out)))
        return out

    def write_phi(self, arr, row0=None):
        row0 = self.first_row if row0 is None else row0
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        _ck(lib().ammsb_store_write_phi(self.h, C.c_uint64(row0), C.c_uint64(arr.size), _p(arr)))

    def read_phi(self, row0=None, nrows=None):
        row0 = self.first_row if row0 is None else row0
        nrows = self.local_rows if nrows is None else nrows
        out = np.empty(nrows, dtype=np.float32)
        _ck(lib().ammsb_store_read_phi(self.h, C.c_uint64(row0), C.c_uint64(nrows), _p(out)))
        return out

    def export_handles(self):
        a = (C.c_uint8 * 64)()
        b = (C.c_uint8 * 64)()
        _ck(lib().ammsb_store_export(self.h, a, b))
        return bytes(a), bytes(b)

    def attach(self, shard, pi_handle, phi_handle):
        a = (C.c_uint8 * 64).from_buffer_copy(pi_handle)
        b = (C.c_uint8 * 64).from_buffer_copy(phi_handle)
        _ck(lib().ammsb_store_attach(self.h, shard, a, b))

    def export_fds(self):
        a, b = C.c_int(-1), C.c_int(-1)
        _ck(lib().ammsb_store_export_fd(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def attach_fds(self, shard, pi_fd, phi_fd):
        _ck(lib().ammsb_store_attach_fd(self.h, shard, pi_fd, phi_fd))

    def add_mirror_fds(self, pi_fd, phi_fd):
        _ck(lib().ammsb_store_add_mirror_fd(self.h, pi_fd, phi_fd))

    def add_mirror(self, pi_handle, phi_handle):
        a = (C.c_uint8 * 64).from_buffer_copy(pi_handle)
        b = (C.c_uint8 * 64).from_buffer_copy(phi_handle)
        _ck(lib().ammsb_store_add_mirror(self.h, a, b))

    def add_mirror_local(self, peer):
        _ck(lib().ammsb_store_add_mirror_local(self.h, peer.h))

    def attach_local(self, shard, peer):
        _ck(lib().ammsb_store_attach_local(self.h, shard, peer.h))

    def local_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        _ck(lib().ammsb_store_local_ptrs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def free(self):
        if self.h is not None:
            lib().ammsb_store_destroy(self.h)
            self.h = None


# ------------------------------------------------------------ column-sharded layout ----
def cols_local_index(k, G):
    """local float index of global column k on its owner rank (k % 32) % G -- csrc/cols.cu"""
    k = np.asarray(k)
    l, i = k & 31, k >> 5
    return ((i >> 2) * (32 // G) + l // G) * 4 + (i & 3)


def cols_owner(k, G):
    return (np.asarray(k) & 31) % G


class Cols:
    """one rank's share of the column-sharded pi (csrc/cols.cu): the columns of the reference
    work-items l = rank (mod world), as [N][K/world], plus the mailbox the peers write into"""

    def __init__(self, ctx, N, K, world, rank, n, max_nodes, max_edges, max_pairs):
        h = C.c_void_p()
        _ck(lib().ammsb_cols_create(ctx.h, C.c_uint64(N), K, world, rank, n, max_nodes, max_edges,
                                    C.c_uint64(max_pairs), C.byref(h)))
        self.h, self.ctx, self.N, self.K, self.world, self.rank = h, ctx, N, K, world, rank

    def export_fd(self):
        fd = C.c_int(-1)
        _ck(lib().ammsb_cols_export_fd(self.h, C.byref(fd)))
        return fd.value

    def attach_fd(self, peer_rank, fd):
        _ck(lib().ammsb_cols_attach_fd(self.h, peer_rank, fd))

    def attach_local(self, peer):
        _ck(lib().ammsb_cols_attach_local(self.h, peer.rank, peer.h))

    def alias_self(self):
        """timing diagnostics only (AMMSB_COLS_LOOPBACK): every unattached peer mailbox = the own one"""
        _ck(lib().ammsb_cols_alias_self(self.h))

    def mailbox_bytes(self):
        n = C.c_size_t(0)
        _ck(lib().ammsb_cols_mailbox_bytes(self.h, C.byref(n)))
        return n.value

    def init_pi(self, eta0=1.0, eta1=1.0):
        _ck(lib().ammsb_cols_init_pi(self.h, C.c_float(eta0), C.c_float(eta1)))

    def write_pi(self, arr, row0=0):
        arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1, self.K)
        _ck(lib().ammsb_cols_write_pi(self.h, C.c_uint64(row0), C.c_uint64(arr.shape[0]), _p(arr)))

    def read_pi(self, out, row0=0):
        """fills the columns this rank owns into out[nrows, K] (other columns untouched)"""
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape[1] == self.K
        _ck(lib().ammsb_cols_read_pi(self.h, C.c_uint64(row0), C.c_uint64(out.shape[0]), _p(out)))
        return out

    def write_phi(self, arr, row0=0):
        arr = np.ascontiguousarray(arr, dtype=np.float32)
        _ck(lib().ammsb_cols_write_phi(self.h, C.c_uint64(row0), C.c_uint64(arr.size), _p(arr)))

    def read_phi(self, row0=0, nrows=None):
        out = np.empty(self.N if nrows is None else nrows, dtype=np.float32)
        _ck(lib().ammsb_cols_read_phi(self.h, C.c_uint64(row0), C.c_uint64(out.size), _p(out)))
        return out

    def write_theta(self, theta, beta):
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        beta = np.ascontiguousarray(beta, dtype=np.float32)
        assert theta.size == 2 * self.K and beta.size == 2 * self.K
        _ck(lib().ammsb_cols_write_theta(self.h, _p(theta), _p(beta)))

    def read_theta(self):
        theta, beta = np.empty(2 * self.K, np.float32), np.empty(2 * self.K, np.float32)
        _ck(lib().ammsb_cols_read_theta(self.h, _p(theta), _p(beta)))
        return theta, beta

    def beta_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        _ck(lib().ammsb_cols_beta_ptr(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def read_phi_vec(self, out):
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape[1] == self.K
        _ck(lib().ammsb_cols_read_phi_vec(self.h, out.shape[0], _p(out)))
        return out

    def check(self):
        t = C.c_uint32(0)
        _ck(lib().ammsb_cols_check(self.h, C.byref(t)))
        if t.value:
            raise AmmsbError("rank %d: a column-layout exchange wait timed out" % self.rank)

    def free(self):
        if self.h is not None:
            lib().ammsb_cols_destroy(self.h)
            self.h = None


def _handles(objs):
    return (C.c_void_p * len(objs))(*[o.h if o is not None else None for o in objs])


def cols_neighbor_sample(ctx, ranks, d_nodes, V, wg, step, pools):
    _ck(lib().ammsb_cols_neighbor_sample(ctx.h, _handles(ranks), len(ranks), d_nodes.ptr, V, wg, step, _handles(pools)))


def cols_update_phi(ctx, ranks, p, opts, train, d_nodes, d_neighbors, V, step, pools):
    """d_neighbors None: the lists delivered by cols_neighbor_sample for the same step"""
    _ck(lib().ammsb_cols_update_phi(ctx.h, _handles(ranks), len(ranks), C.byref(p), C.byref(opts), train.h,
                                    d_nodes.ptr, d_neighbors.ptr if d_neighbors is not None else None, V, step,
                                    _handles(pools)))


def cols_update_pi(ctx, ranks, d_nodes, V, step):
    _ck(lib().ammsb_cols_update_pi(ctx.h, _handles(ranks), len(ranks), d_nodes.ptr, V, step))


def cols_update_beta(ctx, ranks, p, train, d_edges, E_mb, scale, step, pools):
    _ck(lib().ammsb_cols_update_beta(ctx.h, _handles(ranks), len(ranks), C.byref(p), train.h, d_edges.ptr, E_mb,
                                     C.c_float(scale), step, _handles(pools)))


def cols_perplexity(ctx, ranks, p, heldout, d_edges, H, call_count, want=True):
    nv = len(ranks)
    sums, avg = np.zeros((nv, 4), np.float64), np.zeros(nv, np.float64)
    _ck(lib().ammsb_cols_perplexity(ctx.h, _handles(ranks), nv, C.byref(p), heldout.h, d_edges.ptr, H, call_count,
                                    _p(sums) if want else None, _p(avg) if want else None))
    return avg, sums
