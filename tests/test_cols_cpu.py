"""CPU tests of the host-side logic of the column-sharded layout (csrc/cols.cu)."""
import numpy as np

import pyammsb as A


def test_cols_index_map_is_a_bijection():
    """every global column has exactly one (owner rank, local index); a rank's local indices are
    0 .. K/G-1; a reference lane's columns stay on one rank in increasing order of k"""
    for G in (2, 4, 8):
        for K in (128, 256, 512, 1024):
            k = np.arange(K)
            owner, loc = A.cols_owner(k, G), A.cols_local_index(k, G)
            for g in range(G):
                assert sorted(loc[owner == g].tolist()) == list(range(K // G))
            for lane in range(32):
                ks = k[k % 32 == lane]
                assert len(set(owner[ks])) == 1 and owner[ks][0] == lane % G
                # float4 q of the lane holds its columns 4q .. 4q+3 contiguously
                assert np.array_equal(loc[ks] % 4, np.arange(len(ks)) % 4)


def test_wg_sum_tree_splits_into_rank_partials():
    """sum.cc:20-42 for 32 lanes: strides 16..G combine lanes of equal l % G (one GPU), strides
    G/2..1 combine the G per-GPU partials -- the identity the column layout rests on (fp32, exact)"""
    rng = np.random.default_rng(0)
    for G in (2, 4, 8):
        for _ in range(50):
            aux = rng.standard_normal(32).astype(np.float32) * np.float32(10.0) ** rng.integers(-6, 6, 32).astype(np.float32)
            ref = aux.copy()
            p2 = 16
            while p2 > 0:
                ref[:p2] = ref[:p2] + ref[p2:2 * p2]
                p2 //= 2
            part = np.zeros(G, np.float32)
            for g in range(G):
                v = aux[g::G].copy()  # lanes g, g+G, ... : local lane li = l // G
                h = len(v) // 2
                while h > 0:
                    v[:h] = v[:h] + v[h:2 * h]
                    h //= 2
                part[g] = v[0]
            h = G // 2
            while h > 0:
                part[:h] = part[:h] + part[h:2 * h]
                h //= 2
            assert part[0] == ref[0]


def test_neighbor_sampler_divide_free_identities():
    """the arithmetic ns_draw_slot (csrc/common.cuh) replaces generate_random_int's `%` with
    (sample.cc:23-46): the 64-bit remainder by N as Lemire's multiply form with 128 fractional bits,
    `x % capacity` as the 32-bit multiply form, and the probe sequence
    (l1 + q (1 + 2 capacity)) % capacity as l1, l1 + 1, ... -- exact for every input"""
    import random
    rnd = random.Random(5)
    M64, M128 = (1 << 64) - 1, (1 << 128) - 1
    for N in (2, 3, 33, 317080, 3997962, 65608366, (1 << 32) - 5):
        m = (M128 // N + 1) & M128  # ceil(2^128 / N), as NsGeom holds it (n_m_hi, n_m_lo)
        for _ in range(2000):
            a = rnd.getrandbits(64) if rnd.random() < 0.8 else rnd.choice([0, 1, N - 1, N, N + 1, M64])
            low = (m * a) & M128                 # fractional part of a / N
            assert (low * N) >> 128 == a % N
    for cap in (16, 64, 66, 100, 128, 254):
        cm = (M64 // cap + 1) & M64            # ceil(2^64 / capacity)
        for _ in range(2000):
            h = rnd.getrandbits(32)
            assert (((cm * h) & M64) * cap) >> 64 == h % cap
            l1, q = h % cap, rnd.randrange(cap)
            assert (l1 + q * (1 + (cap << 1))) % cap == (l1 + q) % cap
