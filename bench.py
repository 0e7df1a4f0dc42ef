#!/usr/bin/env python3
"""bench.py -- the measurement contract of the repo.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one SG-MCMC iteration of the reference's hot path (Learner::Run body,
learner.cc:214-250): neighbor sampling -> update_phi -> update_pi -> update_beta/theta on one
mini-batch drawn by the reference's stratified-random-node strategy.  Workload = BASELINE.json
configs[1]: synthetic com-DBLP-shaped graph (N=317080, E=1049866), K=1024, 10% held-out,
mini-batch m=16384 edges per GPU, n=32 sampled neighbors (SURVEY.md section 8d canonical point).

Prints ONE JSON line (rank 0).  `value` = mini-batch edges/s with every input resident in HBM
(operators driven through the C ABI of include/ammsb.h); `e2e` = the same metric through the
user-facing call (mcmc::Learner::Run at N=1, the sharded driver at N>1) with HOST mini-batches:
host sampling, H2D of edges/nodes and a D2H read of beta inside the timed region.
`--impl reference` times the reference's own CPU implementation (oracle/_ref: the reference's
kernel text + host sampler compiled from /root/reference; falls back to the oracle port) on the
box's host cores for the same metric on a bounded sample of the same workload.

N > 1 (weak scaling, m edges per GPU): the pi layout is a function of the arguments
(`sharded_plan`): column shards (csrc/cols.cu) when a GPU's piece of a row is at most 512 bytes
(8 GPUs at K = 1024), else a copy per GPU while it fits, else node partitions; every run asserts one
iteration bit-identical to one GPU (`parity_vs_n1`).  AMMSB_STAGE_EVENTS=1 adds per-stage CUDA-event
times of the run itself (`stages_in_run_ms`), AMMSB_BENCH_VARIANTS="name:ENV=VAL,..;.." repeats the
timed steps under other kernel switches in the same process, --e2e-device-sampler (N = 1) also times
mcmc::Learner::Run with Config::device_sampler.

Other shapes (development; the contract run uses the defaults): --shape com-LiveJournal |
com-Friendster | com-Friendster-eighth ..., --K/--m/--n.  --graph device (default above 100 M
edges) builds the synthetic graph, the cuckoo sets, the held-out pairs and the adjacency in HBM
and draws the mini-batches on the device (csrc/graph.cu) through the sharded driver -- the only
way to run the full com-Friendster shape (1.8 G edges).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mcmc-ammsb-gpu_b200")
sys.path[:0] = [PKG]

METRIC = "sgmcmc_minibatch_edges_per_s"
UNIT = "edges/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # workload overrides (development only; the contract run uses the defaults)
    ap.add_argument("--shape", default="com-DBLP")
    ap.add_argument("--N", type=int, default=0)
    ap.add_argument("--E", type=int, default=0)
    ap.add_argument("--K", type=int, default=0)
    ap.add_argument("--m", type=int, default=16384, help="mini-batch edges per GPU")
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--store", default="auto", choices=["auto", "partitioned", "replicated", "columns"],
                    help="multi-GPU pi layout: node-partitioned (NVLink peer loads), one copy per GPU, or "
                         "column-sharded (every GPU holds K/G columns of every row; partial sums cross NVLink)")
    ap.add_argument("--collectives", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU exchange steps: own kernels over NVLink peer memory, or NCCL")
    ap.add_argument("--graph", default="auto", choices=["auto", "host", "device"],
                    help="multi-GPU runs: build the synthetic graph, the cuckoo sets and the mini-batches on the host "
                         "(the reference's path) or in HBM (csrc/graph.cu; auto: device above 100 M edges)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-device-sampler", action="store_true",
                    help="N=1: also time mcmc::Learner::Run with Config::device_sampler (reported as e2e_device_sampler)")
    return ap.parse_args()


def workload(args):
    import synth
    N, E, K, r = synth.SHAPES[args.shape]
    N, E, K = args.N or N, args.E or E, args.K or K
    return dict(shape=args.shape, N=N, E=E, K=K, heldout_ratio=r, m=args.m, n=args.n)


def workload_name(w, world):
    return ("synthetic %s-shaped graph N=%d E=%d, K=%d, heldout_ratio=%.2f, strategy=Node, "
            "mini_batch=%d edges%s, neighbors=%d" %
            (w["shape"], w["N"], w["E"], w["K"], w["heldout_ratio"], w["m"] * world,
             " (%d per GPU)" % w["m"] if world > 1 else "", w["n"]))


def sharded_plan(args, w, world):
    """layout decisions of a multi-GPU run -- a function of the arguments only, so that both
    arms describe the same run"""
    mode = args.store
    if mode == "auto":
        # column shards when a GPU's piece of a row is at most 512 bytes (the shape csrc/cols.cu's
        # update_phi keeps 12+ slots per SM in flight at: 8 GPUs at K = 1024, 4 or 8 at K = 512 --
        # measured 120 M edges/s against 97 M for the replicated layout at the DBLP shape on 8 GPUs);
        # otherwise a full copy per GPU while it is a small part of 180 GB, else node partitions
        cols_ok = world in (2, 4, 8) and w["K"] in (128, 256, 512, 1024) and w["K"] // world <= 128 and w["n"] <= 32
        mode = "columns" if cols_ok else ("replicated" if 4.0 * w["N"] * w["K"] <= 48e9 else "partitioned")
    gmode = args.graph
    if gmode == "auto":
        gmode = "device" if w["E"] > 100e6 else "host"
    return mode, args.collectives, gmode


def config_of(args, w, world):
    """the `config` object of the JSON line: identical in the b200 and the reference arm"""
    N, K, n, m = w["N"], w["K"], w["n"], w["m"] * world
    cfg = {"workload": workload_name(w, world),
           "l2": "inputs larger than L2: pi is %.2f GB and every non-link step gathers %.2f GB of rows"
                 % (4.0 * N * K / 1e9, bytes_phi(m + 1, n, K) / 1e9),
           "timing": "b200 arm: CUDA events on the launching stream, steps are whole iterations in stream order, "
                     "neighbor sampling of the next mini-batch overlaps on a second stream (the Learner::Run "
                     "schedule), max over ranks; reference arm: host wall clock, CPU only"}
    if world > 1 or args.graph == "device":
        mode, coll, gmode = sharded_plan(args, w, world)
        cfg["graph"] = "built in HBM (csrc/graph.cu)" if gmode == "device" else "built on the host"
        cfg["parallelism"] = (
            "pi column-sharded over %d GPUs (GPU g holds the K/%d columns of the reference work-items l = g mod %d "
            "of every row: all row reads are local HBM), partial sums of update_phi / update_pi / update_beta / "
            "perplexity exchanged inside the kernels through mailboxes in NVLink peer memory and completed in the "
            "WG_SUM tree order (bit-identical to one GPU); no barrier or all-reduce launches" % (world, world, world)
        ) if mode == "columns" else (
            ("pi/phi node-partitioned over %d GPUs (NVLink peer loads), " % world if mode == "partitioned" else
             "pi/phi replicated on %d GPUs (fits: %.1f GB), mini-batch slots split over GPUs, updated rows written "
             "to every copy by NVLink peer stores, " % (world, 4.0 * N * K / 1e9)) +
            ("beta gradient and perplexity sums all-reduced in rank order and phases ordered by our own kernels "
             "over NVLink peer memory" if coll == "peer" else "beta gradient and perplexity sums all-reduced (NCCL)"))
    return cfg


# --------------------------------------------------------------------- bytes --
def bytes_phi(V, n, K):      # B_phi = V[(n+2)4K + 68n + 8]
    return V * ((n + 2) * 4 * K + n * 68 + 8)


def bytes_pi(V, K):
    return V * (8 * K + 8)


def bytes_ns(V, n):
    return V * (4 + 4 * n + 32)


def bytes_beta(E_mb, K):
    return E_mb * (8 * K + 72) + 24 * K


def bytes_ppx(H, K):
    return H * (8 * K + 88) + 8 * K


# -------------------------------------------------------------------- clocks --
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.02)  # nvidia-smi takes a moment to deliver its first sample
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                    power.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(names, f[5:9]):
                    if v == "Active":
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        else:  # the sampler delivered nothing inside the window: one synchronous query, flagged as such
            try:
                txt = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                f = [x.strip() for x in txt.strip().split(",")]
                out.update(sm_mhz=float(f[1]), sm_max_mhz=float(f[2]), samples=1, power_w_max=float(f[3]),
                           reasons=[n for n, v in zip(names, f[5:9]) if v == "Active"],
                           note="sampled right after the timed region")
            except Exception:
                pass
        return out


# -------------------------------------------------------- CPU reference legs --
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(w, steps, warmup, seconds_budget, log=lambda *a: None):
    """The reference's CPU implementation of the path on the host cores: its own host sampler
    (sample.cc) + its kernel text (THREAD / EDGE_PER_THREAD variants -- the ones the reference
    selects for CPU devices, learner.cc:105-114) run as OpenMP loops over work-items, on the SAME
    mini-batch size as the B200 arm.  The run is bounded in time, never in mini-batch size:
    steps=None runs for about `seconds_budget` seconds (cpu_baseline leg); otherwise `steps`
    iterations, or as many as fit `seconds_budget` (at least 2)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    import synth
    fast = os.path.join(ROOT, "oracle", "_ref", "libref_oracle_fast.so")
    chk = os.path.join(ROOT, "oracle", "_ref", "libref_oracle.so")
    flags = ""
    try:
        flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0]
    except Exception:
        pass
    if os.path.exists(fast) and " avx2" in flags and " fma" in flags:
        orc, kind = pyoracle.Oracle(path=fast), "reference"
        build = "timing build of the reference sources: g++ -O3 -ffast-math -mavx2 -mfma -fopenmp"
    elif os.path.exists(chk):
        orc, kind = pyoracle.Oracle(path=chk), "reference"
        build = "checker build of the reference sources: g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp"
    else:
        orc, kind = pyoracle.Oracle(omp=True), "port"
        build = "oracle port: gcc -O2 -ffp-contract=off -fopenmp"
    L = orc.L
    L.orc_num_threads.restype = C.c_int
    # launchers (torchrun) export OMP_NUM_THREADS=1: the baseline gets every core of the box
    L.orc_set_num_threads.argtypes = [C.c_int]
    L.orc_set_num_threads(host_cores())
    cores = int(L.orc_num_threads())
    N, E, K, n, m = w["N"], w["E"], w["K"], w["n"], w["m"]
    t0 = time.time()
    keys = synth.make_edges(N, E, 1)
    training_len = int(np.ceil((1 - w["heldout_ratio"] / 2) * E))  # data.cc:86-88
    heldout_links, training = keys[:E - training_len], keys[E - training_len:]
    train_set = orc.set_build(training)
    heldout_set = orc.set_build(heldout_links)
    # model state: gamma-initialised pi (numpy stream: the values do not affect the timing)
    rng = np.random.default_rng(11)
    pi = rng.standard_gamma(1.0, size=(N, K), dtype=np.float32)
    phi = pi.sum(axis=1, dtype=np.float32)
    pi /= phi[:, None]
    theta = rng.standard_gamma(1.0, size=2 * K).astype(np.float32) + np.float32(1e-3)
    beta = orc.theta_to_beta(theta)
    p = orc.make_params(N, E, K, n)
    log("cpu reference (%s, %d threads, %s): setup %.1fs" % (kind, cores, build, time.time() - t0))

    if kind == "reference":
        L.ref_sampler_create.restype = C.c_void_p
        L.ref_sampler_create.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                         C.c_uint64, C.c_uint64]
        L.ref_sample.restype = C.c_float
        L.ref_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.ref_sampler_max_fan_out.restype = C.c_uint64
        L.ref_sampler_max_fan_out.argtypes = [C.c_void_p]
        h = C.c_void_p(L.ref_sampler_create(N, E, training.ctypes.data, len(training),
                                            heldout_links.ctypes.data, len(heldout_links), m))
        cap = max(2 * m, 1 + int(L.ref_sampler_max_fan_out(h)))
        eb, nb = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
        seed = C.c_uint(12345)

        def draw():
            ne, nn = C.c_uint64(0), C.c_uint64(0)
            wgt = L.ref_sample(h, 0, C.byref(seed), eb.ctypes.data, C.byref(ne), nb.ctypes.data, C.byref(nn))
            return eb[:ne.value].copy(), nb[:nn.value].copy(), float(wgt)
    else:
        import pymcmc  # oracle port: the host sampler is the repo's own (golden-checked) one
        cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=n, heldout_ratio=w["heldout_ratio"])
        cfg.set_graph(N, keys)
        seed = C.c_uint(12345)

        def draw():
            wgt, edges, nodes = cfg.sample("Node", seed)
            return edges, nodes, wgt

    state = dict(step=0, pools=None)

    def iteration():
        edges, nodes, weight = draw()
        V = len(nodes)
        if state["pools"] is None or state["pools"][0] < V:
            cap = max(2 * V, 64)
            state["pools"] = (cap, orc.rng_pool(cap * 2 * n, 56, 57), orc.rng_pool(cap, 42, 43))
        _, npool, ppool = state["pools"]
        state["step"] += 1
        nbrs, _ = orc.neighbor_sample(npool, nodes, N, n, 32)
        vec = orc.update_phi(pyoracle.MODE_THREAD, 32, p, beta, pi, phi, train_set, nodes, nbrs, state["step"], ppool)
        orc.update_pi(pyoracle.MODE_THREAD, 32, K, pi, phi, vec, nodes)
        orc.update_beta(pyoracle.MODE_THREAD, 32, p, theta, beta, pi, train_set, edges, weight, state["step"],
                        state.setdefault("bpool", orc.rng_pool(K, 44, 45)))
        return len(edges)

    # warm-up: the CPU has no clocks to ramp; at most 2 untimed iterations (pools, page faults)
    for _ in range(2 if steps is None else min(max(warmup, 1), 2)):
        iteration()
    edges_done, iters = 0, 0
    t0 = time.perf_counter()
    while True:
        edges_done += iteration()
        iters += 1
        if steps is not None and iters >= steps:
            break
        if iters >= 2 and time.perf_counter() - t0 >= seconds_budget:
            break
    dt = time.perf_counter() - t0
    sample = ("%d full iterations (host sampler + neighbor sampling + update_phi + update_pi + update_beta, "
              "THREAD-mode kernels as the reference selects for CPU devices) of the same graph/K/n at the same "
              "mini_batch=%d edges (strategy Node), %.1f s of CPU work on %d threads; %s"
              % (iters, m, dt, cores, build))
    return dict(value=edges_done / dt, unit=UNIT, cores=cores, kind=kind, sample=sample,
                iterations=iters, seconds=dt, edges=edges_done, build=build)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same workload as the B200 arm at this N: the global mini-batch is N x m edges.  The run is
    # bounded to ~150 s by timing fewer iterations, never by a smaller mini-batch.
    r = cpu_reference(dict(w, m=w["m"] * args.gpus), args.steps, args.warmup, 150.0,
                      log=lambda *a: print(*a, file=sys.stderr))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(r["iterations"], 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, args.gpus),
        "iterations_per_s": r["iterations"] / r["seconds"], "iterations_timed": r["iterations"],
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm --
class View:
    """a device pointer + size seen as a pyammsb buffer (sub-range of a DevBuf / torch tensor)"""

    def __init__(self, ptr, nbytes=0):
        self.ptr, self.nbytes = C.c_void_p(ptr), nbytes


def run_b200(args, w):
    import torch
    import pyammsb as A
    import pymcmc
    import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch N>1 with torch.distributed.run" % (args.gpus, world))
    if A.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device and no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)
    if world > 1 or args.graph == "device":  # the sharded driver (also on one GPU for device-built graphs)
        import dist as D
        return D.bench_sharded(args, w, rank, world, local_rank, log, METRIC, UNIT, config_of(args, w, world),
                               sharded_plan(args, w, world), ClockSampler)

    N, E, K, n, m = w["N"], w["E"], w["K"], w["n"], w["m"]
    t0 = time.time()
    keys = synth.make_edges(N, E, 1)
    cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=n, heldout_ratio=w["heldout_ratio"], strategy="Node")
    cfg.set_graph(N, keys)
    log("graph + split + cuckoo sets + adjacency: %.1fs (training %d, held-out pairs %d, max fan-out %d)" %
        (time.time() - t0, len(cfg.edges()[0]), len(cfg.edges()[1]), cfg.max_fan_out()))

    # ---------------- value leg: operators through the C ABI, inputs resident in HBM ----------
    stream = torch.cuda.current_stream()
    ctx = A.Ctx(local_rank)
    ctx.set_stream(stream.cuda_stream)
    p = cfg.params()
    store = A.Store(ctx, N, K)
    store.init_pi()
    t_tab, t_bins, t_prime = cfg.set_table(0)
    h_tab, h_bins, h_prime = cfg.set_table(1)
    dts, dhs = A.DevSet(ctx, t_tab, t_bins, t_prime), A.DevSet(ctx, h_tab, h_bins, h_prime)
    theta = pymcmc.init_theta_host(K)
    beta = (theta.reshape(K, 2) / theta.reshape(K, 2).sum(axis=1, keepdims=True)).astype(np.float32).ravel()
    d_theta, d_beta = ctx.from_host(theta), ctx.from_host(beta)
    Vmax, Emax = cfg.max_nodes(), cfg.max_edges()
    npools = [A.Rng(ctx, Vmax * 2 * n, 56, 57) for _ in range(2)]
    ppool, bpool = A.Rng(ctx, Vmax * 32, 42, 43), A.Rng(ctx, K, 44, 45)
    d_nb = ctx.buf(np.uint32, Vmax * n)
    d_vec, d_sum = ctx.buf(np.float32, Vmax * K), ctx.buf(np.float32, Vmax)
    d_ts, d_g = ctx.buf(np.float32, K), ctx.buf(np.float32, 2 * K)
    ws = ctx.buf(np.uint8, ctx.beta_workspace_bytes(K))
    opts = A.PhiOpts(A.MODE_WG, 32, 0, 0)

    total = args.warmup + args.steps
    seed = C.c_uint(12345)
    t0 = time.time()
    batches = [cfg.sample("Node", seed) for _ in range(total)]  # (weight, edges, nodes)
    log("host sampler: %d mini-batches in %.2fs" % (total, time.time() - t0))
    e_off = np.cumsum([0] + [len(b[1]) for b in batches])
    v_off = np.cumsum([0] + [len(b[2]) for b in batches])
    d_edges_all = ctx.from_host(np.concatenate([b[1] for b in batches]))
    d_nodes_all = ctx.from_host(np.concatenate([b[2] for b in batches]))

    # Neighbor sampling of mini-batch i+1 runs on a second stream while mini-batch i is processed
    # -- the schedule of Learner::Run (reference learner.cc:216-232: the next Sample is prepared
    # concurrently on its own queue); double-buffered neighbor lists, ordered by events.
    side = torch.cuda.Stream()
    ctx_side = A.Ctx(local_rank)
    ctx_side.set_stream(side.cuda_stream)
    d_nbs = [d_nb, ctx.buf(np.uint32, Vmax * n)]
    ns_done = [torch.cuda.Event(), torch.cuda.Event()]
    phi_done = [torch.cuda.Event(), torch.cuda.Event()]
    sampled = [-1]

    def sample_neighbors(i):
        if i >= total or sampled[0] >= i:
            return
        nodes = batches[i][2]
        side.wait_event(phi_done[i & 1])  # update_phi of mini-batch i-2 has released this buffer
        ctx_side.neighbor_sample(npools[i & 1], View(d_nodes_all.ptr.value + 4 * int(v_off[i])), len(nodes), N, n, 32,
                                 d_nbs[i & 1])
        ns_done[i & 1].record(side)
        sampled[0] = i

    def device_step(i, step_no, ev=None):
        wgt, edges, nodes = batches[i]
        V, Emb = len(nodes), len(edges)
        dn = View(d_nodes_all.ptr.value + 4 * int(v_off[i]))
        de = View(d_edges_all.ptr.value + 8 * int(e_off[i]))
        sample_neighbors(i)
        stream.wait_event(ns_done[i & 1])
        if ev is not None:
            ev[0].record(stream)
        ctx.update_phi(p, opts, d_beta, store, dts, dn, d_nbs[i & 1], V, step_no, ppool, d_vec, d_sum)
        if ev is not None:
            ev[1].record(stream)
        phi_done[i & 1].record(stream)
        sample_neighbors(i + 1)
        ctx.update_pi(K, store, d_vec, d_sum, dn, V)
        ctx.update_beta(p, d_theta, d_beta, store, dts, de, Emb, wgt, step_no, bpool, d_ts, d_g, ws)

    # clocks are sampled from before the warm-up until after the per-stage timings: the GPU is busy
    # throughout, and the timed region lies inside that window
    clocks = ClockSampler(local_rank)
    clocks.start()
    for i in range(args.warmup):
        device_step(i, i + 1)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = A.launch_count()
    torch.cuda.synchronize()
    e_start.record(stream)
    for k in range(args.steps):
        device_step(args.warmup + k, args.warmup + k + 1, evs[k])
    e_stop.record(stream)
    torch.cuda.synchronize()
    launches = A.launch_count() - launches0
    dev_ms = e_start.elapsed_time(e_stop)
    timed = batches[args.warmup:]
    edges_timed = int(sum(len(b[1]) for b in timed))
    # update_phi launches of the timed region by kernel: non-link mini-batches (V = m + 1 slots)
    # run k_update_phi_fast / _team (bandwidth-bound), link ones (V = 1 + deg(u) <= #SMs slots)
    # k_update_phi_split (a latency chain)
    sms = ctx.sm_count()
    per_launch = [(len(b[2]), a.elapsed_time(e)) for b, (a, e) in zip(timed, evs)]
    big_l = [(V, t) for V, t in per_launch if V > sms]
    small_l = [(V, t) for V, t in per_launch if V <= sms]
    phi_ms = float(sum(t for _, t in per_launch))
    phi_big_ms = float(sum(t for _, t in big_l))
    phi_big_bytes = float(sum(bytes_phi(V, n, K) for V, _ in big_l))
    value = edges_timed / (dev_ms * 1e-3)

    # per-stage table on the canonical non-link mini-batch (V = m + 1), each stage timed alone
    stages = {}
    big = max(range(total), key=lambda i: len(batches[i][2]))
    Vb, Eb = len(batches[big][2]), len(batches[big][1])
    dn = View(d_nodes_all.ptr.value + 4 * int(v_off[big]))
    de = View(d_edges_all.ptr.value + 8 * int(e_off[big]))
    H = len(cfg.edges()[1])
    d_hedges = ctx.from_host(cfg.edges()[1])
    d_ppx = ctx.buf(np.float32, H).zero()
    pws = ctx.buf(np.uint8, ctx.perplexity_workspace_bytes())
    calls = [0]

    def t_stage(name, fn, nbytes, reps=20):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        t = float(np.median(ts))
        stages[name] = {"ms": round(t, 4), "algorithmic_GB": round(nbytes / 1e9, 4),
                        "GBps": round(nbytes / t / 1e6, 1)}

    def f_ppx():
        calls[0] += 1
        ctx.perplexity(p, store, d_beta, dhs, d_hedges, H, d_ppx, calls[0], pws)

    t_stage("neighbor_sample", lambda: ctx.neighbor_sample(npools[0], dn, Vb, N, n, 32, d_nb), bytes_ns(Vb, n))
    t_stage("update_phi", lambda: ctx.update_phi(p, opts, d_beta, store, dts, dn, d_nb, Vb, 7, ppool, d_vec, d_sum),
            bytes_phi(Vb, n, K))
    t_stage("update_pi", lambda: ctx.update_pi(K, store, d_vec, d_sum, dn, Vb), bytes_pi(Vb, K))
    t_stage("update_beta", lambda: ctx.update_beta(p, d_theta, d_beta, store, dts, de, Eb, 2.0 * E / m, 7, bpool,
                                                   d_ts, d_g, ws), bytes_beta(Eb, K))
    t_stage("perplexity", f_ppx, bytes_ppx(H, K))
    clk = clocks.stop()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    # DRAM traffic is an ncu measurement (profiles/): quoted only for the launch shape it was
    # captured on, scaled by nothing
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "update_phi_traffic.json")))["canonical_launch"]
        if (cap["V"], cap["n"], cap["K"]) == (Vb, n, K):
            traffic = int(cap["dram_bytes_read"] + cap["dram_bytes_write"])
    except Exception:
        pass
    kernel = "k_update_phi_fast" if K <= 1024 else "k_update_phi_team"
    achieved = phi_big_bytes / (phi_big_ms * 1e-3) / 1e9 if big_l else 0.0
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "launches": len(big_l),
                "algorithmic_bytes_per_launch": round(phi_big_bytes / max(len(big_l), 1)),
                "avg_launch_ms": round(phi_big_ms / max(len(big_l), 1), 4),
                "share_of_step": round(phi_big_ms / dev_ms, 4),
                "what": "non-link mini-batches only (V = m + 1 slots), CUDA events around every launch of the "
                        "timed region; traffic = ncu dram read + write of one launch of this shape "
                        "(profiles/update_phi_traffic.json), null for any other shape",
                "link_launches": {"kernel": "k_update_phi_split", "launches": len(small_l),
                                  "avg_launch_ms": round(sum(t for _, t in small_l) / max(len(small_l), 1), 4),
                                  "avg_slots": round(sum(V for V, _ in small_l) / max(len(small_l), 1), 1),
                                  "bound": "latency (a handful of slots)",
                                  "share_of_step": round(sum(t for _, t in small_l) / dev_ms, 4)},
                "canonical_launch": dict(stages["update_phi"], V=Vb, frac=round(stages["update_phi"]["GBps"] / peak, 4))}
    for b in (d_edges_all, d_nodes_all, d_hedges, d_ppx, pws, d_nb, d_nbs[1], d_vec, d_sum, d_ts, d_g, ws, d_theta, d_beta):
        b.free()
    for r in npools + [ppool, bpool]:
        r.free()
    dts.free(); dhs.free(); store.free()
    ctx.sync()

    # ---------------- e2e leg: mcmc::Learner::Run with host mini-batches ----------------------
    def learner_run(c, what):
        lrn = pymcmc.Learner(c, local_rank)
        mirror = torch.empty(2 * K, dtype=torch.float32).pin_memory()
        lrn.mirror_beta(mirror.data_ptr())  # every iteration ends with a D2H copy of beta[2K]
        lrn.run(args.warmup)
        torch.cuda.synchronize()
        b0, e0 = lrn.h2d_bytes(), lrn.edges_processed()
        t0 = time.perf_counter()
        lrn.run(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert float(mirror.sum()) > 0
        if rank == 0:
            lrn.print_stats()  # stage breakdown of the host path, to stderr
        e_edges = lrn.edges_processed() - e0
        t1 = time.perf_counter()
        ppx = lrn.heldout_perplexity()
        ppx_s = time.perf_counter() - t1
        out = {"value": e_edges / dt, "unit": UNIT, "h2d_bytes_per_step": (lrn.h2d_bytes() - b0) / args.steps,
               "d2h_bytes_per_step": 8 * K, "iterations_per_s": args.steps / dt, "ms_per_step": 1e3 * dt / args.steps,
               "api": what, "heldout_perplexity": ppx, "perplexity_eval_s": ppx_s}
        lrn.close()
        return out

    e2e = e2e_dev = None
    if not args.no_e2e:
        e2e = learner_run(cfg, "mcmc::Learner::Run(steps): per iteration the host mini-batch sampler (sample.cc "
                               "strategies; two sampler streams, each a 3-stage thread pipeline over a ring of 6 "
                               "mini-batches), H2D of edges/nodes, 5 kernels (neighbor sampling, update_phi, update_pi, "
                               "beta partials, beta reduction + theta step), D2H of beta[2K] into pinned host memory; "
                               "2 iterations in flight")
        if args.e2e_device_sampler:
            cfg.set(device_sampler=1)
            e2e_dev = learner_run(cfg, "mcmc::Learner::Run(steps) with Config::device_sampler: the host draws the coin "
                                       "and the vertex, the device the mini-batch in the reference's order (the same "
                                       "mini-batches element for element), D2H of edges/nodes for the checkpoint "
                                       "state, neighbor sampling, 4 kernels, D2H of beta[2K]")
            cfg.set(device_sampler=0)

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference(w, None, 0, args.cpu_seconds, log=log)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(args, w, 1),
        "iterations_per_s": args.steps / (dev_ms * 1e-3),
        "perplexity_eval_s": stages["perplexity"]["ms"] * 1e-3,
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        "e2e_device_sampler": e2e_dev,
        "stages": stages,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    w = workload(args)
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner)
    # are sent to stderr; print() keeps writing to the real stdout through sys.stdout
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)
    real_stdout.flush()


if __name__ == "__main__":
    main()
