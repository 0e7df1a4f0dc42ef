"""devgraph.py -- the data side of a run built in HBM: synthetic edge list, training / held-out
split, cuckoo edge sets, held-out pairs, adjacency, and the Node mini-batch strategy.

The reference does all of this on the host (main.cc:101-154 -> data.cc:80-128, cuckoo.cc:117-197,
Graph data.cc:12-25, sample.cc:253-302).  At the com-Friendster shape (N = 65.6 M, E = 1.8 G) a
host build needs minutes and ~50 GB per process; on the device it is seconds.  The split follows
GenerateSetsFromEdges: with the edge list in its (pseudo-random) order, the first
E - ceil((1 - r/2) E) edges are the held-out links, the rest the training edges, and as many
non-links (in neither set) are appended to the held-out pairs.  The cuckoo sets have the host's
geometry and hash functions, so every kernel looks them up as usual.
"""
import ctypes as C
import math

import numpy as np

import pyammsb as A


class _View:
    """a device pointer with the .ptr of a pyammsb buffer"""

    def __init__(self, ptr):
        self.ptr = C.c_void_p(ptr)


def split_sizes(E, heldout_ratio):
    """(training edges, held-out links) of GenerateSetsFromEdges (data.cc:86-88):
    training = ceil((1 - r/2) E) in double, r being Config::heldout_ratio, a float"""
    r = float(np.float32(heldout_ratio))
    training_len = int(math.ceil((1 - r / 2) * E))
    return training_len, int(E) - training_len


class DeviceGraph:
    def __init__(self, ctx, N, E, heldout_ratio, seed=1, log=lambda *a: None):
        import time
        t0 = time.time()
        self.ctx, self.N, self.E = ctx, int(N), int(E)
        self.num_training, self.num_heldout_links = split_sizes(E, heldout_ratio)
        edges = ctx.buf(np.uint64, self.E)
        A.graph_generate(ctx, self.N, self.E, seed, edges)
        links = _View(edges.ptr.value)
        training = _View(edges.ptr.value + 8 * self.num_heldout_links)
        log("device graph: %d edges generated in %.1fs" % (self.E, time.time() - t0))
        t0 = time.time()
        self.train = A.BuiltSet(ctx, training, self.num_training)
        self.heldout = A.BuiltSet(ctx, links, self.num_heldout_links)
        log("device graph: cuckoo sets (%d + %d keys) built in %.1fs" %
            (self.num_training, self.num_heldout_links, time.time() - t0))
        t0 = time.time()
        # held-out pairs: the links, then as many non-links (data.cc:110-126)
        self.H = 2 * self.num_heldout_links
        self.d_heldout_pairs = ctx.buf(np.uint64, max(self.H, 1))
        if self.num_heldout_links:
            A._ck(A.lib().ammsb_d2d(ctx.h, self.d_heldout_pairs.ptr, links.ptr, C.c_size_t(8 * self.num_heldout_links)))
            A.graph_nonlinks(ctx, self.N, self.num_heldout_links, seed + 1, self.train, self.heldout,
                             _View(self.d_heldout_pairs.ptr.value + 8 * self.num_heldout_links))
        # adjacency of the training graph (mcmc::Graph) for the link mini-batches
        self.d_offsets = ctx.buf(np.uint64, self.N + 1)
        self.d_adj = ctx.buf(np.uint32, max(2 * self.num_training, 1))
        d_degree = ctx.buf(np.uint32, self.N)
        A.graph_csr(ctx, self.N, training, self.num_training, self.d_offsets, self.d_adj, d_degree)
        self.degree = d_degree.read()
        d_degree.free()
        edges.free()
        self.max_fan_out = int(self.degree.max()) if self.N else 0
        log("device graph: held-out pairs + adjacency in %.1fs (max fan-out %d)" % (time.time() - t0, self.max_fan_out))

    # capacities the reference allocates with (sample.cc:86-97,129-131)
    def max_nodes(self, m):
        return max(2 * m, 1 + self.max_fan_out)

    def max_edges(self, m):
        return max(m, self.max_fan_out)

    def sampler(self, m, ctx=None):
        return A.DeviceSampler(ctx or self.ctx, self.N, self.E, m, self.train, self.heldout, self.d_offsets,
                               self.d_adj, self.degree)

    def free(self):
        for b in (self.d_heldout_pairs, self.d_offsets, self.d_adj):
            b.free()
        self.train.free()
        self.heldout.free()


def host_order_csr(N, training):
    """mcmc::Graph (data.cc:12-25) as arrays: for every training edge (u, v) in list order, v joins
    u's neighbors and u joins v's -- the order sampleNodeLink (sample.cc:253-272) walks them in"""
    u = (training >> np.uint64(32)).astype(np.int64)
    v = (training & np.uint64(0xffffffff)).astype(np.int64)
    ends = np.stack([u, v], axis=1).ravel()
    other = np.stack([v, u], axis=1).ravel()
    order = np.argsort(ends, kind="stable")  # per vertex, in order of appearance
    deg = np.bincount(ends, minlength=N)
    off = np.zeros(N + 1, np.uint64)
    off[1:] = np.cumsum(deg)
    return off, other[order].astype(np.uint32), deg.astype(np.uint32)


class UploadedGraph:
    """A HOST-built graph (a pymcmc.Config after set_graph: the reference's own split, cuckoo
    tables and Graph) put on the device once, with the interface of DeviceGraph.  The mini-batches
    are then drawn on the GPU by the Node strategy in the reference's order (csrc/graph.cu +
    csrc/orderset.cu): the same edges and nodes, element for element, as the host strategy
    (sample.cc:253-302 + learner.cc:162-173) draws from the same seed -- the host only draws the
    coin and the vertex.  This takes the host sampler out of the iteration loop."""

    def __init__(self, ctx, cfg, log=lambda *a: None):
        import time
        t0 = time.time()
        p = cfg.params()
        self.ctx, self.N, self.E = ctx, int(p.N), int(p.E)
        tr, he = cfg.edges()
        self.num_training, self.num_heldout_links = len(tr), len(he) // 2
        self.train = A.DevSet(ctx, *cfg.set_table(0))
        self.heldout = A.DevSet(ctx, *cfg.set_table(1))
        self.H = len(he)
        self.d_heldout_pairs = ctx.from_host(he) if self.H else ctx.buf(np.uint64, 1)
        off, adj, deg = host_order_csr(self.N, tr)
        self.d_offsets, self.d_adj, self.degree = ctx.from_host(off), ctx.from_host(adj), deg
        self.max_fan_out = int(deg.max()) if self.N else 0
        assert self.max_fan_out == cfg.max_fan_out()
        self._max_nodes, self._max_edges = cfg.max_nodes(), cfg.max_edges()
        log("host graph on the device (sets, held-out pairs, adjacency in Graph order): %.1fs" % (time.time() - t0))

    def max_nodes(self, m):
        return max(2 * m, 1 + self.max_fan_out, self._max_nodes)

    def max_edges(self, m):
        return max(m, self.max_fan_out, self._max_edges)

    def sampler(self, m, ctx=None):
        return A.DeviceSampler(ctx or self.ctx, self.N, self.E, m, self.train, self.heldout, self.d_offsets,
                               self.d_adj, self.degree, exact_order=True)

    def free(self):
        for b in (self.d_heldout_pairs, self.d_offsets, self.d_adj):
            b.free()
        self.train.free()
        self.heldout.free()
