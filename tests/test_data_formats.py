"""CPU tests of the data formats either side of the hot path: the SNAP
text loader against the reference's own data.cc (golden fixture generated from it, and live when
oracle/_ref is present), and the gzip dataset dump of main.cc:109-143 against the byte layout the
reference writes."""
import ctypes as C
import gzip
import os
import struct

import numpy as np
import pytest

import pymcmc
from util import REF_SO, make_edges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_snap(lib, fn, path, seed, cap):
    f = getattr(lib, fn)
    f.restype = C.c_int64
    f.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint64]
    out, nv = np.zeros(cap, dtype=np.uint64), C.c_uint64(0)
    n = f(str(path).encode(), seed, C.byref(nv), out.ctypes.data_as(C.c_void_p), cap)
    assert n >= 0, n
    return nv.value, out[:n]


def test_snap_loader_matches_golden(tmp_path):
    """skip 4 header lines, renumber vertices in std::unordered_set order, sort, de-duplicate,
    std::random_shuffle with libc rand() -- the edge list, order included, of the reference"""
    g = np.load(os.path.join(GOLD, "snap.npz"))
    path = tmp_path / "graph.txt"
    path.write_bytes(g["text"].tobytes())
    N, edges = load_snap(pymcmc.lib(), "mcmc_unique_edges_from_file", path, int(g["srand_seed"]), 10000)
    assert N == int(g["N"]) == 300
    assert np.array_equal(edges, g["edges"])
    assert len(np.unique(edges)) == len(edges)


@pytest.mark.parametrize("seed", [1, 99])
def test_snap_loader_matches_reference_code(tmp_path, seed):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference); the golden fixture still applies")
    rng = np.random.default_rng(seed)
    ids = rng.choice(10 ** 7, size=5000, replace=False)
    pairs = ids[rng.integers(0, 5000, size=(40000, 2))]
    path = tmp_path / "g.txt"
    with open(path, "w") as f:
        f.write("# a\n# b\n# c\n# d\n")
        for a, b in pairs:
            f.write("%d %d\n" % (a, b))
    N_r, e_r = load_snap(C.CDLL(REF_SO), "ref_unique_edges_from_file", path, seed, 50000)
    N_h, e_h = load_snap(pymcmc.lib(), "mcmc_unique_edges_from_file", path, seed, 50000)
    assert N_h == N_r and np.array_equal(e_h, e_r)


def test_dataset_dump_is_the_references_gzip_layout(tmp_path):
    """main.cc:109-143: gzip stream of u64 N, f32 heldout_ratio, u64 num_edges, u64 edges[]"""
    L = pymcmc.lib()
    N, r = 4000, np.float32(0.05)
    edges = make_edges(N, 30000, 2)
    path = str(tmp_path / "d.gz")
    assert L.mcmc_dump_dataset(path.encode(), C.c_uint64(N), C.c_float(r), edges.ctypes.data_as(C.c_void_p),
                               C.c_uint64(len(edges))) == 0
    raw = gzip.open(path, "rb").read()
    assert raw[:20] == struct.pack("<QfQ", N, r, len(edges))
    assert np.array_equal(np.frombuffer(raw[20:], dtype=np.uint64), edges)
    # and a file written the reference's way loads back
    other = str(tmp_path / "e.gz")
    with gzip.open(other, "wb") as f:
        f.write(struct.pack("<QfQ", N + 1, np.float32(0.25), len(edges) - 7) + edges[:-7].tobytes())
    L.mcmc_load_dataset.restype = C.c_int64
    out, n_, r_ = np.zeros(len(edges), dtype=np.uint64), C.c_uint64(0), C.c_float(0)
    n = L.mcmc_load_dataset(other.encode(), C.byref(n_), C.byref(r_), out.ctypes.data_as(C.c_void_p),
                            C.c_uint64(len(out)))
    assert n == len(edges) - 7 and n_.value == N + 1 and np.float32(r_.value) == np.float32(0.25)
    assert np.array_equal(out[:n], edges[:-7])
    assert L.mcmc_load_dataset(str(tmp_path / "missing.gz").encode(), C.byref(n_), C.byref(r_),
                               out.ctypes.data_as(C.c_void_p), C.c_uint64(len(out))) == -1


def test_graph_adjacency_is_symmetric_and_complete():
    """data-test.cc:27-53 of the reference: every edge appears in both endpoints' neighbor lists,
    nothing else does; lists are in edge-list order (the order sampleNodeLink emits in) and
    MaxFanOut is the largest list"""
    N, E = 800, 6000
    keys = make_edges(N, E, 4)
    cfg = pymcmc.Config(K=8, mini_batch_size=8, heldout_ratio=0.1)
    cfg.set_graph(N, keys)
    tr, he = cfg.edges()
    L = pymcmc.lib()
    L.mcmc_config_neighbors.restype = C.c_int64
    for which, edges in ((0, tr), (1, he)):
        want = [[] for _ in range(N)]
        for e in edges:
            u, v = int(e) >> 32, int(e) & 0xFFFFFFFF
            want[u].append(v)
            want[v].append(u)
        fan = 0
        for u in range(N):
            buf = np.zeros(256, dtype=np.uint32)
            d = L.mcmc_config_neighbors(cfg.h, which, C.c_uint32(u), buf.ctypes.data_as(C.c_void_p), C.c_uint64(256))
            assert d == len(want[u]) and buf[:d].tolist() == want[u]
            fan = max(fan, d)
        if which == 0:
            assert cfg.max_fan_out() == fan
    cfg.close()
