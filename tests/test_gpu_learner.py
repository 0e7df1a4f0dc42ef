"""GPU tests of the C++ host API (mcmc::Learner over the C ABI) against the CPU oracle:
per-iteration parity, free-running perplexity trajectory, checkpoint/resume determinism
(serialize-test.cc:90-134), all six mini-batch strategies, and the CLI binary."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import pyammsb as A
import pymcmc
from util import make_edges, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5  # north_star: pi/beta within 1e-5 relative per iteration
PPX_TOL = 1e-3  # north_star: perplexity trajectory within 1e-3 after N iterations


def make_cfg(N=1200, E=9000, K=64, m=64, n=16, seed=4, **kw):
    cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=n, heldout_ratio=0.1, **kw)
    cfg.set_graph(N, make_edges(N, E, seed))
    return cfg


class OracleLearner:
    """The reference schedule (learner.cc:214-250) restated over the oracle operators; the
    mini-batches (edges, nodes, weight) are the ones the device Learner is about to consume,
    everything else -- neighbor sampling, phi/pi, beta/theta, perplexity -- is the oracle's."""

    def __init__(self, orc, cfg, lrn, N, K, n):
        self.orc, self.cfg, self.N, self.K, self.n = orc, cfg, N, K, n
        self.p = orc.make_params(N, len(cfg.edges()[0]) + 0, K, n)
        self.p.E = cfg.params().E
        tr, he = cfg.edges()
        self.train_set = orc.set_build(tr)
        nl = len(he) // 2
        self.heldout_set = orc.set_build(he[:nl])
        self.heldout_edges = he
        self.pi, self.phi, self.beta, self.theta = lrn.read(N, K)
        self.phi_pool = orc.rng_pool(cfg.max_nodes() * 32, 42, 43)
        self.beta_pool = orc.rng_pool(K, 44, 45)
        self.nb_pools = [orc.rng_pool(cfg.max_nodes() * 2 * n, 56, 57) for _ in range(2)]
        self.phase = 0
        self.step = 0
        self.ppx_per_edge = np.zeros(len(he), dtype=np.float32)
        self.ppx_calls = 0

    def iterate(self, edges, nodes, weight):
        nbrs = self.iterate_phi_pi(nodes)
        self.iterate_beta(edges, weight)
        return nbrs

    def iterate_phi_pi(self, nodes):
        """neighbor sampling, update_phi, update_pi of the next iteration"""
        o = self.orc
        nbrs, _ = o.neighbor_sample(self.nb_pools[self.phase], nodes, self.N, self.n, 32)
        self.step += 1
        vec = o.update_phi(A.MODE_WG, 32, self.p, self.beta, self.pi, self.phi, self.train_set, nodes, nbrs,
                           self.step, self.phi_pool)
        o.update_pi(A.MODE_WG, 32, self.K, self.pi, self.phi, vec, nodes)
        return nbrs

    def iterate_beta(self, edges, weight):
        """update_beta of the iteration iterate_phi_pi began (reads self.pi: a test may put the
        device's rows there first, so that the stage is judged from identical input)"""
        self.orc.update_beta(A.MODE_WG, 32, self.p, self.theta, self.beta, self.pi, self.train_set, edges, weight,
                             self.step, self.beta_pool)
        self.phase = 1 - self.phase

    def perplexity(self):
        self.ppx_calls += 1
        avg, _ = self.orc.perplexity(A.MODE_WG, 32, self.p, self.pi, self.beta, self.heldout_set,
                                     self.heldout_edges, self.ppx_per_edge, self.ppx_calls)
        return float(np.exp(np.float32(avg)))


def close_enough(got, want, what, frac=5e-3):
    """rtol 1e-5 on all but the ill-conditioned elements (cancellation in the Langevin update;
    the reference's own THREAD and WG variants differ by more there, test_gpu_parity.py); those
    are bounded in absolute terms: 2e-6 of the largest element of their row / vector"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    e = rel_err(got, want)
    assert np.median(e) < 1e-6, (what, np.median(e))
    assert float((e > RTOL).mean()) < frac, (what, float((e > RTOL).mean()), e.max())
    scale = np.abs(want).max(axis=-1, keepdims=True)
    assert (np.abs(got - want) / scale).max() < 2e-6, (what, (np.abs(got - want) / scale).max())


@pytest.mark.parametrize("strategy", ["Node", "BF"])
def test_learner_iterations_match_oracle(ctx, orc, strategy):
    N, K, n = 1200, 64, 16
    cfg = make_cfg(N=N, K=K, n=n, strategy=strategy)
    lrn = pymcmc.Learner(cfg, 0)
    ol = OracleLearner(orc, cfg, lrn, N, K, n)
    # initial state is the reference's: libstdc++ gamma stream for theta, device gamma for pi
    pi0, phi0 = orc.init_pi(N, K)
    close_enough(ol.pi, pi0, "init pi")
    assert np.array_equal(ol.theta, pymcmc.init_theta_host(K))
    got, want = lrn.heldout_perplexity(), ol.perplexity()  # one call each: it is a running mean
    assert abs(got - want) <= PPX_TOL * want
    sizes = []
    for it in range(40):
        edges, nodes, nbrs, weight = lrn.peek(n)
        sizes.append(len(nodes))
        want_nbrs = ol.iterate(edges, nodes, weight)
        assert np.array_equal(nbrs, want_nbrs), "iteration %d: neighbor ids differ" % it
        lrn.run(1)
        if it < 12:
            pi, phi, beta, theta = lrn.read(N, K)
            # per-iteration gate: compare, then continue the oracle from the device state so
            # that every iteration is judged on its own (iterations 12.. run free)
            close_enough(pi[nodes], ol.pi[nodes], "pi it %d" % it)
            close_enough(phi[nodes], ol.phi[nodes], "phi it %d" % it)
            close_enough(theta, ol.theta, "theta it %d" % it, frac=2e-2)
            close_enough(beta, ol.beta, "beta it %d" % it, frac=2e-2)
            untouched = np.ones(N, bool)
            untouched[nodes] = False
            assert np.array_equal(pi[untouched], ol.pi[untouched])
            ol.pi, ol.phi, ol.beta, ol.theta = pi, phi, beta, theta
    assert len(set(sizes)) > 1 or strategy != "Node"  # both link and non-link mini-batches occurred
    assert lrn.edges_processed() > 0
    got, want = lrn.heldout_perplexity(), ol.perplexity()
    print("perplexity after 40 free-running iterations: device %.6f oracle %.6f" % (got, want))
    assert abs(got - want) <= PPX_TOL * want
    lrn.close()
    cfg.close()


def test_learner_perplexity_trajectory(ctx, orc):
    """free-running (never re-synchronised) 200 iterations: perplexity within 1e-3 throughout"""
    N, K, n = 800, 32, 8
    cfg = make_cfg(N=N, E=6000, K=K, m=32, n=n, seed=9)
    lrn = pymcmc.Learner(cfg, 0)
    ol = OracleLearner(orc, cfg, lrn, N, K, n)
    worst = 0.0
    for block in range(10):
        for _ in range(20):
            edges, nodes, _, weight = lrn.peek(n)
            ol.iterate(edges, nodes, weight)
            lrn.run(1)
        got, want = lrn.heldout_perplexity(), ol.perplexity()
        worst = max(worst, abs(got - want) / want)
    print("worst perplexity deviation over 200 iterations: %.3e (last %.6f vs %.6f)" % (worst, got, want))
    assert worst <= PPX_TOL
    lrn.close()
    cfg.close()


def test_learner_checkpoint_resume_is_bit_exact(ctx):
    """serialize-test.cc:90-134 EndToEnd: 10 iterations, Serialize, 10 more -> ppx; a fresh
    Learner Parse()s the file, runs 10 -> ppx2; ASSERT_EQ(ppx, ppx2)."""
    cfg = make_cfg(N=1024, E=1024 * 4, K=32, m=32, n=8, seed=12)
    path = os.path.join(tempfile.mkdtemp(), "ckpt.bin")
    lrn = pymcmc.Learner(cfg, 0)
    lrn.run(10)
    lrn.serialize(path)
    lrn.run(10)
    ppx = lrn.heldout_perplexity()
    state = lrn.read(1024, 32)
    lrn.close()
    lrn2 = pymcmc.Learner(cfg, 0)
    lrn2.parse(path)
    lrn2.run(10)
    ppx2 = lrn2.heldout_perplexity()
    state2 = lrn2.read(1024, 32)
    assert ppx == ppx2
    for a, b in zip(state, state2):
        assert np.array_equal(a, b)
    lrn2.close()
    cfg.close()


@pytest.mark.parametrize("strategy", pymcmc.STRATEGIES)
def test_learner_runs_every_strategy(ctx, strategy):
    cfg = make_cfg(N=600, E=5000, K=32, m=32, n=8, seed=2, strategy=strategy)
    lrn = pymcmc.Learner(cfg, 0)
    p0 = lrn.heldout_perplexity()
    lrn.run(25)
    p1 = lrn.heldout_perplexity()
    pi, phi, beta, theta = lrn.read(600, 32)
    assert np.isfinite(p0) and np.isfinite(p1)
    assert np.allclose(pi.sum(axis=1), 1.0, atol=1e-4) and (pi > 0).all() and (phi > 0).all()
    assert np.allclose(beta[0::2] + beta[1::2], 1.0, atol=1e-5)
    lrn.close()
    cfg.close()


def test_learner_phi_modes_agree(ctx):
    """wg-phi-test.cc:116-158: THREAD vs WG-NAIVE state within 2% with noise disabled."""
    outs = []
    for mode in ("THREAD", "WG-NAIVE"):
        cfg = make_cfg(N=700, E=5000, K=64, m=64, n=8, seed=5, phi_mode=mode, phi_disable_noise=1)
        lrn = pymcmc.Learner(cfg, 0)
        for _ in range(3):
            lrn.peek(8)
            lrn.run(1)
        outs.append(lrn.read(700, 64))
        lrn.close()
        cfg.close()
    assert rel_err(outs[0][0], outs[1][0]).max() < 0.02
    assert rel_err(outs[0][1], outs[1][1]).max() < 0.02


def test_cli_runs_on_a_snap_file(ctx, tmp_path):
    """main.cc: SNAP text in (4 header lines), ppx[...] lines out"""
    edges = make_edges(500, 4000, 3)
    f = tmp_path / "g.txt"
    with open(f, "w") as out:
        out.write("# a\n# b\n# c\n# d\n")
        for e in edges:
            out.write("%d\t%d\n" % (int(e) >> 32, int(e) & 0xffffffff))
    exe = os.path.join(ROOT, "mcmc-ammsb-gpu_b200", "ammsb-main")
    r = subprocess.run([exe, "-f", str(f), "-k", "16", "-m", "16", "-n", "8", "-x", "30", "-i", "10"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (r.stdout + r.stderr).count("ppx[") >= 3


def test_cli_devices_and_device_sampler(ctx, tmp_path):
    """ammsb-main --devices 0,0 (mcmc::ShardedLearner, two ranks on one GPU) and --device-sampler 1 print
    the same kind of ppx[...] trace as the plain run; the device-sampler run prints the SAME trace
    (its mini-batches are the host strategy's)"""
    edges = make_edges(1500, 12000, 3)
    f = tmp_path / "g.txt"
    with open(f, "w") as out:
        out.write("# a\n# b\n# c\n# d\n")
        for e in edges:
            out.write("%d\t%d\n" % (int(e) >> 32, int(e) & 0xffffffff))
    exe = os.path.join(ROOT, "mcmc-ammsb-gpu_b200", "ammsb-main")
    base = [exe, "-f", str(f), "-k", "128", "-m", "64", "-n", "16", "-x", "30", "-i", "10"]
    traces = {}
    for name, extra in (("plain", []), ("device-sampler", ["--device-sampler", "1"]), ("devices", ["--devices", "0,0"])):
        r = subprocess.run(base + extra, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (name, r.stderr[-2000:])
        traces[name] = [ln for ln in (r.stdout + r.stderr).splitlines() if ln.startswith("ppx[")]
        assert len(traces[name]) >= 4, (name, r.stderr[-2000:])
    assert traces["device-sampler"] == traces["plain"]
    last = lambda t: float(t[-1].split("=")[1])
    assert abs(last(traces["devices"]) - last(traces["plain"])) <= 1e-2 * last(traces["plain"])


def test_training_perplexity_option(ctx, orc):
    """MCMC_CALC_TRAIN_PPX (learner.cc:47-75,204-212) as a run-time switch"""
    N, K, n = 900, 32, 8
    cfg = make_cfg(N=N, E=7000, K=K, m=32, n=n, seed=6, calc_train_ppx=1, training_ppx_ratio=0.05)
    lrn = pymcmc.Learner(cfg, 0)
    lrn.run(5)
    edges = lrn.training_perplexity_edges()
    tr, he = cfg.edges()
    n_links = int(0.05 * len(tr))
    assert np.array_equal(edges[:n_links], tr[:n_links]) and len(edges) > n_links
    got = lrn.training_perplexity()
    pi, phi, beta, theta = lrn.read(N, K)
    ts = orc.set_build(tr)
    p = orc.make_params(N, int(cfg.params().E), K, n)
    avg, sums = orc.perplexity(A.MODE_WG, 32, p, pi, beta, ts, edges, np.zeros(len(edges), np.float32), 1)
    assert sums[2] == n_links
    assert abs(got - float(np.exp(np.float32(avg)))) <= 1e-5 * got
    lrn.close()
    cfg.close()


def test_learner_grqc_shape_matches_oracle(ctx, orc):
    """BASELINE.json configs[0]: synthetic ca-GrQc-shaped graph (N=5242, E=14496), K=64,
    stratified-random-node mini-batches with the reference CLI defaults m=32, n=32, r=0.01"""
    N, E, K, m, n = 5242, 14496, 64, 32, 32
    cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=n, heldout_ratio=0.01, strategy="Node")
    cfg.set_graph(N, make_edges(N, E, 1))
    assert len(cfg.edges()[1]) == 2 * (E - int(np.ceil((1 - 0.01 / 2) * E)))  # data.cc:86-99
    lrn = pymcmc.Learner(cfg, 0)
    ol = OracleLearner(orc, cfg, lrn, N, K, n)
    got, want = lrn.heldout_perplexity(), ol.perplexity()
    assert abs(got - want) <= PPX_TOL * want
    for it in range(60):
        edges, nodes, nbrs, weight = lrn.peek(n)
        assert np.array_equal(nbrs, ol.iterate(edges, nodes, weight))
        lrn.run(1)
    pi, phi, beta, theta = lrn.read(N, K)
    e = rel_err(pi, ol.pi)
    assert np.median(e) == 0 or np.median(e) < 1e-6  # most rows untouched, touched rows close
    got, want = lrn.heldout_perplexity(), ol.perplexity()
    print("ca-GrQc shape, 60 free-running iterations: device %.6f oracle %.6f" % (got, want))
    assert abs(got - want) <= PPX_TOL * want
    lrn.close()
    cfg.close()


@pytest.mark.parametrize("strategy", ["Node", "NodeLink", "NodeNonLink"])
def test_learner_device_sampler_is_the_host_strategy(ctx, strategy):
    """Config::device_sampler: mcmc::Learner::Run with the mini-batches drawn on the device
    (csrc/graph.cu + csrc/orderset.cu) consumes the host strategy's mini-batches element for
    element (sample.cc:253-302 + learner.cc:162-173) -- edges, nodes, weight, sampled neighbors --
    and therefore reaches the same state bit for bit"""
    N, K, n = 3000, 64, 16
    runs = []
    for dev in (0, 1):
        cfg = make_cfg(N=N, E=40000, K=K, m=256, n=n, seed=6, strategy=strategy, device_sampler=dev)
        lrn = pymcmc.Learner(cfg, 0)
        mbs = []
        for _ in range(10):
            mbs.append(lrn.peek(n))
            lrn.run(1)
        lrn.run(7)  # several mini-batches in flight
        runs.append((mbs, lrn.read(N, K), lrn.heldout_perplexity(), lrn.edges_processed()))
        lrn.close()
        cfg.close()
    (mb_h, st_h, ppx_h, ne_h), (mb_d, st_d, ppx_d, ne_d) = runs
    sizes = set()
    for (e_h, v_h, nb_h, w_h), (e_d, v_d, nb_d, w_d) in zip(mb_h, mb_d):
        assert np.array_equal(e_h, e_d) and np.array_equal(v_h, v_d), "device mini-batch differs from the host strategy's"
        assert np.array_equal(nb_h, nb_d) and w_h == w_d
        sizes.add(len(e_h) == 256)
    if strategy == "Node":
        assert sizes == {True, False}  # link and non-link mini-batches occurred
    for a, b in zip(st_h, st_d):
        assert np.array_equal(a, b)
    assert ppx_h == ppx_d and ne_h == ne_d


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_learner_matches_the_one_gpu_learner(ctx, world):
    """mcmc::ShardedLearner (column-sharded pi, host/mcmc/sharded_learner.cc) with all ranks emulated
    on one device -- the very kernels and mailbox protocol of a multi-GPU run -- against mcmc::Learner:
    after one iteration pi is the one-GPU result bit for bit (update_phi / update_pi complete their
    sums in the reference's tree order), theta / beta agree on every rank and with one GPU within the
    fp32 association of the gradient sum; free-running, the perplexity stays within 1e-3"""
    N, K, n = 2400, 128, 16
    os.environ["AMMSB_PHI_NOSPLIT"] = "1"  # one association of the gradient sum on both sides
    try:
        cfg = make_cfg(N=N, E=30000, K=K, m=128, n=n, seed=8, strategy="Node")
        one = pymcmc.Learner(cfg, 0)
        cfg2 = make_cfg(N=N, E=30000, K=K, m=128, n=n, seed=8, strategy="Node")  # srand again: same sampler seeds
        many = pymcmc.ShardedLearner(cfg2, [0] * world)
        pi0, phi0, beta0, theta0 = one.read(N, K)
        spi, sphi, sbeta, stheta = many.read(N, K)
        assert np.array_equal(spi, pi0) and np.array_equal(sphi, phi0)
        assert all(np.array_equal(stheta[r], theta0) and np.array_equal(sbeta[r], beta0) for r in range(world))
        one.run(1)
        many.run(1)
        pi1, phi1, beta1, theta1 = one.read(N, K)
        spi, sphi, sbeta, stheta = many.read(N, K)
        assert (pi1 != pi0).any()
        assert np.array_equal(spi, pi1), "column-sharded update_phi / update_pi differ from the one-GPU kernels"
        assert np.array_equal(sphi, phi1)
        for r in range(1, world):
            assert np.array_equal(stheta[r], stheta[0]) and np.array_equal(sbeta[r], sbeta[0])
        close_enough(stheta[0], theta1, "theta", frac=2e-2)
        close_enough(sbeta[0], beta1, "beta", frac=2e-2)
        one.run(40)
        many.run(40)
        assert many.edges_processed() == one.edges_processed()
        got, want = many.heldout_perplexity(), one.heldout_perplexity()
        print("world %d: perplexity after 41 iterations %.6f, one GPU %.6f" % (world, got, want))
        assert abs(got - want) <= PPX_TOL * want
        spi, sphi, sbeta, stheta = many.read(N, K)
        for r in range(1, world):
            assert np.array_equal(sbeta[r], sbeta[0])
        one.close(); many.close(); cfg.close(); cfg2.close()
    finally:
        del os.environ["AMMSB_PHI_NOSPLIT"]
