#!/usr/bin/env python3
"""torchrun entry (one rank per GPU): the sharded driver (dist.ShardedLearner) at world_size G
against the same driver at world_size 1 run by every rank on its own GPU.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_dist_gpu.py
Checks: after one iteration pi/phi are bit-identical (the Langevin stream is owned by units, the
initial beta is the same), theta/beta agree within the reduction-order tolerance; after 12
iterations everything agrees within tolerance; perplexity matches."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("mcmc-ammsb-gpu_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402

import dist as D  # noqa: E402
import pymcmc  # noqa: E402
from util import make_edges, rel_err  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, K, n, m = 3000, 256, 16, 512
    cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=n, heldout_ratio=0.1, strategy="Node")
    cfg.set_graph(N, make_edges(N, 30000, 3))
    mode = sys.argv[1] if len(sys.argv) > 1 else "partitioned"
    coll = sys.argv[2] if len(sys.argv) > 2 else "peer"
    if mode == "columns":
        # the column kernels sum the gradient neighbor by neighbor like the one-warp-per-slot kernel;
        # the one-GPU run must not switch to the CTA-per-slot kernel for small mini-batches
        os.environ["AMMSB_PHI_NOSPLIT"] = "1"
    sharded = D.ShardedLearner(cfg, rank, world, local, prefetch=False, store_mode=mode, collectives=coll)
    single = D.ShardedLearner(cfg, 0, 1, local, prefetch=False)
    lo, hi = sharded.local_rows()

    def compare(tag, exact_pi):
        tdist.barrier()
        torch.cuda.synchronize()
        a, b = sharded.read_local_pi(), single.read_local_pi()[lo:hi]
        if mode == "columns":  # all rows, the columns this rank owns
            own = ~np.isnan(a)
            assert own.sum() == a.size // world
            a, b = a[own], b[own]
        if exact_pi:
            assert np.array_equal(a, b), "%s: pi differs" % tag
        e = rel_err(a, b)
        assert np.median(e) < 1e-6 and (e > 1e-5).mean() < 5e-3, (tag, e.max())
        e = rel_err(sharded.read_beta(), single.read_beta())
        assert np.median(e) < 1e-6 and (e > 1e-5).mean() < 2e-2, (tag, "beta", e.max())

    p0, q0 = sharded.heldout_perplexity(), single.heldout_perplexity()
    assert abs(p0 - q0) <= 1e-5 * q0, (p0, q0)
    sizes = set()
    for it in range(12):
        wgt, edges, nodes = sharded.next_minibatch()
        w2, e2, n2 = single.next_minibatch()
        assert np.array_equal(edges, e2) and np.array_equal(nodes, n2) and wgt == w2
        sizes.add(len(nodes))
        for L in (sharded, single):
            d_nodes = L.ctx.from_host(nodes)
            d_edges = L.ctx.from_host(edges)
            L.device_step(d_nodes, d_edges, len(nodes), len(edges), wgt, it % L.STREAMS)
            L.stream.synchronize()
            d_nodes.free(); d_edges.free()
        compare("iteration %d" % it, exact_pi=(it == 0))
    assert len(sizes) > 1
    p1, q1 = sharded.heldout_perplexity(), single.heldout_perplexity()
    assert abs(p1 - q1) <= 1e-3 * q1, (p1, q1)
    # the host path (sampler threads + pinned staging) runs too
    e2e = D.ShardedLearner(cfg, rank, world, local, prefetch=True, store_mode=mode, collectives=coll)
    # replicas of theta/beta are bit-identical across ranks (rank-ordered sum / same NCCL result)
    mine = torch.from_numpy(sharded.read_beta()).cuda()
    ref = mine.clone()
    tdist.broadcast(ref, 0)
    assert torch.equal(mine, ref), "beta replicas differ across ranks"
    e2e.run(6)
    assert np.isfinite(e2e.heldout_perplexity())
    tdist.barrier()
    if rank == 0:
        print("dist gpu check ok (%s): world %d, perplexity %.5f -> %.5f (single %.5f -> %.5f)" % (mode, world, p0, p1, q0, q1))
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
