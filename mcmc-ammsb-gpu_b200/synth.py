"""Synthetic graphs of the shapes BASELINE.json names (there is no network for the SNAP files).

Edges are unique undirected pairs u < v with uniform endpoints, encoded as the reference's
64-bit keys (types.h:66-74: (min << 32) | max) and shuffled -- i.e. what
GetUniqueEdgesFromFile (data.cc:36-78) hands to GenerateSetsFromEdges."""
import numpy as np

SHAPES = {
    # name: (N, E, K, heldout_ratio)   -- SURVEY.md section 8 config table
    "ca-GrQc": (5242, 14496, 64, 0.01),
    "com-DBLP": (317080, 1049866, 1024, 0.10),
    "com-LiveJournal": (3997962, 34681189, 1024, 0.01),
    "com-Friendster": (65608366, 1806067135, 512, 0.01),
    # Friendster-sized pi store (N x K = 134 GB) over a graph with a reduced edge count that a
    # host can build in a minute: exercises the partitioned store and its NVLink gathers at the
    # Friendster scale without the 1.8 G-edge host-side graph build
    "com-Friendster-store": (65608366, 20000000, 512, 0.01),
    # one GPU's eighth of the Friendster shape (the per-GPU footprint of the 8-GPU run: 16.8 GB of
    # pi, a 2 GB cuckoo table) -- with the graph built in HBM (bench.py --graph device)
    "com-Friendster-eighth": (8201046, 225758392, 512, 0.01),
}


def make_edges(N, E, seed=1):
    rng = np.random.default_rng(seed)
    keys = np.zeros(0, dtype=np.uint64)
    while len(keys) < E:
        want = int((E - len(keys)) * 1.1) + 1024
        u = rng.integers(0, N, size=want, dtype=np.uint64)
        v = rng.integers(0, N, size=want, dtype=np.uint64)
        m = u != v
        lo, hi = np.minimum(u[m], v[m]), np.maximum(u[m], v[m])
        keys = np.unique(np.concatenate([keys, (lo << np.uint64(32)) | hi]))
    return keys[rng.permutation(len(keys))[:E]]
