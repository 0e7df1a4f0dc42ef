"""CPU tests (gloo, world_size 2) of the host-side logic of the multi-GPU path (dist.py): the
static work plan, the handle exchange, and the two reductions -- with the oracle standing in
for the kernels, so the arithmetic of "partial results per rank + all-reduce == one device"
is checked without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("mcmc-ammsb-gpu_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

import dist as D  # noqa: E402


def test_slot_partition_covers_every_slot_once():
    for V in (1, 7, 33, 4097, 16385, 70000):
        for world in (1, 2, 4, 8):
            r = D.slot_ranks(V, world)
            assert r.min() >= 0 and r.max() < world
            units = D.phi_units(V)
            # a unit (and so its RNG state) always belongs to one rank, whatever V is
            for u in {0, min(1, units - 1), units - 1}:
                assert len(set(r[u::units].tolist())) == 1 and r[u] == u % world
            counts = np.bincount(r, minlength=world)
            assert counts.sum() == V and counts.max() - counts.min() <= (V + units - 1) // units


def test_chunks_are_a_balanced_partition():
    for total in (0, 1, 5, 16384, 104986):
        for world in (1, 2, 3, 8):
            spans = [D.chunk(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import pyoracle
    from util import Problem
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    orc = pyoracle.Oracle()
    prob = Problem(orc, 400, 32, 4000, 8, seed=5)  # same seed -> every rank builds the same problem
    # handle exchange
    got = D.exchange((rank, bytes([rank]) * 64), world)
    assert [g[0] for g in got] == list(range(world)) and got[rank][1] == bytes([rank]) * 64

    # file-descriptor exchange (how the shareable pi shards travel between processes)
    import tempfile
    f = tempfile.TemporaryFile()
    f.write(b"rank%d" % rank)
    f.flush()
    peers = D.exchange_fds(rank, world, [f.fileno()])
    assert sorted(peers) == [r for r in range(world) if r != rank]
    for peer, fds in peers.items():
        assert os.pread(fds[0], 16, 0) == b"rank%d" % peer
        os.close(fds[0])

    # beta gradient: per-rank chunk + all-reduce == whole mini-batch on one device
    edges = prob.minibatch_edges(101, 3)
    scale, step = 12.5, 3

    def grads_of(e):
        th, be = prob.theta.copy(), prob.beta.copy()
        _, g = orc.update_beta(pyoracle.MODE_WG, 32, prob.p_orc, th, be, prob.pi, prob.train_set, e, scale, step,
                               orc.rng_pool(prob.K, 44, 45))
        return g, th
    lo, hi = D.chunk(len(edges), rank, world)
    g_part, _ = grads_of(edges[lo:hi])
    t = torch.from_numpy(g_part.copy())
    tdist.all_reduce(t)
    g_full, _ = grads_of(edges)
    err = np.abs(t.numpy() - g_full) / (np.abs(g_full) + 1e-12)
    assert np.median(err) < 1e-6 and err.max() < 1e-3, err.max()

    # perplexity: per-rank chunk of held-out pairs + all-reduce of the 4 sums
    H = len(prob.heldout_edges)
    lo, hi = D.chunk(H, rank, world)
    _, sums = orc.perplexity(pyoracle.MODE_WG, 32, prob.p_orc, prob.pi, prob.beta, prob.heldout_set,
                             prob.heldout_edges[lo:hi], np.zeros(hi - lo, np.float32), 1)
    t = torch.from_numpy(sums.copy())
    tdist.all_reduce(t)
    avg_full, sums_full = orc.perplexity(pyoracle.MODE_WG, 32, prob.p_orc, prob.pi, prob.beta, prob.heldout_set,
                                         prob.heldout_edges, np.zeros(H, np.float32), 1)
    s = t.numpy()
    assert s[2] == sums_full[2] and s[3] == sums_full[3]
    avg = -(s[0] + s[1]) / (s[2] + s[3])
    assert abs(avg - avg_full) <= 1e-6 * abs(avg_full)

    # update_phi: the slots of each rank, run separately, reproduce the single-device result
    # bit for bit (RNG state is owned by unit, units are owned by ranks)
    V = 37
    nodes = prob.minibatch_nodes(V, 2)
    nbrs, _ = orc.neighbor_sample(orc.rng_pool(V * 16, 56, 57), nodes, prob.N, prob.n, 32)
    full = orc.update_phi(pyoracle.MODE_WG, 32, prob.p_orc, prob.beta, prob.pi, prob.phi, prob.train_set, nodes, nbrs,
                          1, orc.rng_pool(V * 32, 42, 43))
    mine = np.nonzero(D.slot_ranks(V, world) == rank)[0]
    pool = orc.rng_pool(V * 32, 42, 43)
    part = np.zeros_like(full)
    for sidx in mine:  # one slot at a time, with the slot's own unit state
        sub = pool[sidx * 32:(sidx + 1) * 32].copy()
        part[sidx] = orc.update_phi(pyoracle.MODE_WG, 32, prob.p_orc, prob.beta, prob.pi, prob.phi, prob.train_set,
                                    nodes[sidx:sidx + 1], nbrs[sidx:sidx + 1], 1, sub)[0]
    t = torch.from_numpy(part)
    tdist.all_reduce(t)
    assert np.array_equal(t.numpy(), full)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    tdist.destroy_process_group()


def test_world2_gloo_reductions_match_single_device(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def test_bench_layout_plan_is_a_function_of_the_arguments():
    """bench.sharded_plan: column shards exactly when a GPU's piece of a row is at most 512 bytes
    (the shape the column update_phi keeps 12+ slots per SM in flight at), else a copy per GPU while
    it fits, else node partitions; both arms of the bench derive the same plan from the arguments"""
    import types
    sys.path.insert(0, ROOT)
    import bench
    def plan(world, K, N, n=32, store="auto", E=10 ** 6):
        args = types.SimpleNamespace(store=store, collectives="peer", graph="auto")
        return bench.sharded_plan(args, dict(N=N, K=K, n=n, E=E), world)
    assert plan(8, 1024, 317080)[0] == "columns"
    assert plan(4, 1024, 317080)[0] == "replicated"
    assert plan(2, 1024, 317080)[0] == "replicated"
    assert plan(8, 512, 65608366, E=1806067135) == ("columns", "peer", "device")
    assert plan(4, 512, 65608366)[0] == "columns"
    assert plan(2, 512, 65608366)[0] == "partitioned"      # 134 GB of pi: no copy per GPU
    assert plan(8, 1024, 317080, n=64)[0] == "replicated"  # the slot-at-a-time kernel stages one chunk of 32 neighbors
    assert plan(8, 1024, 317080, store="partitioned")[0] == "partitioned"


def test_host_order_csr_is_the_graph_adjacency():
    """devgraph.host_order_csr: mcmc::Graph's neighbor lists (data.cc:12-25), in its order"""
    sys.path.insert(0, os.path.join(ROOT, "mcmc-ammsb-gpu_b200"))
    try:
        import devgraph
    except OSError:
        pytest.skip("libammsb.so not built")
    tr = np.array([(3 << 32) | 5, (1 << 32) | 3, (0 << 32) | 5, (3 << 32) | 4], dtype=np.uint64)
    off, adj, deg = devgraph.host_order_csr(6, tr)
    lists = [adj[int(off[u]):int(off[u + 1])].tolist() for u in range(6)]
    assert lists == [[5], [3], [], [5, 1, 4], [3], [3, 0]]
    assert deg.tolist() == [1, 1, 0, 3, 1, 2]
