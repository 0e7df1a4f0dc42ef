"""-m gpu: the CUDA path against the committed golden vectors that were generated from the
reference's own code (tests/golden/make_golden.py).  No oracle library is involved here."""
import os

import numpy as np
import pytest

import pyammsb as A
from util import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "operators.npz"))


def params(g):
    return A.make_params(int(g["N"]), int(g["E"]), int(g["K"]), int(g["n"]))


def test_golden_rng_streams(ctx):
    g = np.load(os.path.join(GOLD, "rng.npz"))
    r = A.Rng(ctx, 8, 42, 43)
    assert np.array_equal(r.draw_u64(16), g["u64"])
    r.free()
    r = A.Rng(ctx, 8, 42, 43)
    got = r.draw_randn(64)
    assert np.array_equal(r.get_state(), g["state_after_randn"])
    assert rel_err(got, g["randn"]).max() < 1e-6
    r.free()
    r = A.Rng(ctx, 8, 11, 113)
    assert rel_err(r.draw_gamma(32, 1.0, 1.0), g["gamma"]).max() < 1e-6
    r.free()
    got = np.array([A.round_param(x) for x in (1 / 64, 1 / 1024, 0.0315, 1024, 0.5, 1e-7, 1.0)], dtype=np.float32)
    assert np.array_equal(got, g["rounded"])


def test_golden_membership_and_sampler(ctx, g):
    ts = A.DevSet(ctx, g["train_table"], int(g["train_bins"]), int(g["train_prime"]))
    assert ts.has(g["train_edges"]).all()
    assert not ts.has(g["heldout_edges"]).any()
    nodes, n, N = g["nodes"], int(g["n"]), int(g["N"])
    r = A.Rng(ctx, len(nodes) * 2 * n, 56, 57)
    d_nodes = ctx.from_host(nodes)
    d_out, d_hash = ctx.buf(np.uint32, nodes.size * n), ctx.buf(np.uint32, nodes.size * 2 * n)
    ctx.neighbor_sample(r, d_nodes, len(nodes), N, n, 32, d_out, d_hash)
    assert np.array_equal(d_out.read().reshape(-1, n), g["neighbors"])
    assert np.array_equal(d_hash.read().reshape(-1, 2 * n), g["sampler_table"])
    assert np.array_equal(r.get_state(), g["sampler_state"])
    r.free(); ts.free()


@pytest.mark.parametrize("tag,mode", [("wg", A.MODE_WG), ("thread", A.MODE_THREAD)])
@pytest.mark.parametrize("strict", [1, 0])
def test_golden_update_phi_pi(ctx, g, tag, mode, strict):
    p = params(g)
    K, V = int(g["K"]), len(g["nodes"])
    st = A.Store(ctx, int(g["N"]), K)
    st.write_pi(g["pi"]); st.write_phi(g["phi"])
    ts = A.DevSet(ctx, g["train_table"], int(g["train_bins"]), int(g["train_prime"]))
    d_nodes, d_nb, d_beta = ctx.from_host(g["nodes"]), ctx.from_host(g["neighbors"]), ctx.from_host(g["beta"])
    d_vec, d_sum = ctx.buf(np.float32, V * K), ctx.buf(np.float32, V)
    states = V * (32 if mode == A.MODE_WG else 1)
    for noise in (0, 1):
        r = A.Rng(ctx, states, 42, 43)
        ctx.update_phi(p, A.PhiOpts(mode, 32, 0 if noise else 1, strict), d_beta, st, ts, d_nodes, d_nb, V, 3,
                       r, d_vec, d_sum)
        got, want = d_vec.read().reshape(V, K), g["phi_vec_%s_%d" % (tag, noise)]
        if strict:
            assert np.array_equal(got, want), "strict kernel must reproduce the reference bit for bit"
        else:
            phi_old = g["pi"][g["nodes"]] * g["phi"][g["nodes"]][:, None]
            scale = np.abs(want) + phi_old + np.abs(want - phi_old) + 0.0315 * p.N / g["phi"][g["nodes"]][:, None]
            assert (np.abs(got.astype(np.float64) - want) / scale).max() < 5e-7
            assert np.median(rel_err(got, want)) < 1e-6
        if noise:
            assert np.array_equal(r.get_state(), g["phi_state_%s" % tag])
        r.free()
    ctx.update_pi(K, st, d_vec, d_sum, d_nodes, V)
    pi_after, phi_after = st.read_pi()[g["nodes"]], st.read_phi()[g["nodes"]]
    if strict:
        assert np.array_equal(pi_after, g["pi_after_%s" % tag])
        assert np.array_equal(phi_after, g["phi_after_%s" % tag])
    else:
        assert rel_err(phi_after, g["phi_after_%s" % tag]).max() < 1e-5
        assert np.median(rel_err(pi_after, g["pi_after_%s" % tag])) < 1e-6
    ts.free(); st.free()


def test_golden_beta_and_perplexity(ctx, g):
    p = params(g)
    K = int(g["K"])
    st = A.Store(ctx, int(g["N"]), K)
    st.write_pi(g["pi"]); st.write_phi(g["phi"])
    ts = A.DevSet(ctx, g["train_table"], int(g["train_bins"]), int(g["train_prime"]))
    hs = A.DevSet(ctx, g["heldout_table"], int(g["heldout_bins"]), int(g["heldout_prime"]))
    d_theta, d_beta = ctx.from_host(g["theta"]), ctx.from_host(g["beta"])
    d_edges = ctx.from_host(g["mb_edges"])
    d_ts, d_g = ctx.buf(np.float32, K), ctx.buf(np.float32, 2 * K)
    ws = ctx.buf(np.uint8, ctx.beta_workspace_bytes(K))
    # perplexity first (it reads the pre-update beta, like the golden run)
    d_h = ctx.from_host(g["heldout_edges"])
    H = len(g["heldout_edges"])
    d_ppx = ctx.buf(np.float32, H).zero()
    pws = ctx.buf(np.uint8, ctx.perplexity_workspace_bytes())
    for i, call in enumerate((1, 2, 3)):
        avg, sums = ctx.perplexity(p, st, d_beta, hs, d_h, H, d_ppx, call, pws)
        assert np.array_equal(sums[2:], g["ppx_sums"][i][2:])
        assert abs(avg - g["ppx_avg"][i]) / abs(g["ppx_avg"][i]) < 1e-5
    assert rel_err(d_ppx.read(), g["ppx_per_edge"]).max() < 1e-5
    r = A.Rng(ctx, K, 44, 45)
    ctx.update_beta(p, d_theta, d_beta, st, ts, d_edges, len(g["mb_edges"]), 17.5, 4, r, d_ts, d_g, ws)
    assert np.array_equal(r.get_state(), g["beta_state"])
    assert np.array_equal(d_ts.read(), g["theta_sum"])
    assert rel_err(d_g.read(), g["grads"]).max() < 1e-5
    th = d_theta.read()
    cond = np.abs(th.astype(np.float64) - g["theta_after"]) / (np.abs(g["theta_after"]) + g["theta"] +
                                                                 np.abs(g["theta_after"] - g["theta"]))
    assert cond.max() < 5e-7
    assert np.median(rel_err(d_beta.read(), g["beta_after"])) < 1e-6
    r.free(); ts.free(); hs.free(); st.free()


def test_golden_init_pi(ctx, g):
    st = A.Store(ctx, 200, 48)
    st.init_pi(1.0, 1.0)
    assert np.array_equal(st.read_pi(), g["init_pi"])
    assert np.array_equal(st.read_phi(), g["init_phi"])
    st.free()
