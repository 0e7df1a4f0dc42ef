"""dist.py -- multi-GPU driver of the SG-MCMC iteration (one process per GPU); pi/phi either
node-partitioned over the GPUs (store_mode "partitioned", the layout for graphs that need the
memory of G GPUs) or held as one full copy per GPU ("replicated", when N*K*4 bytes fit: every
read is local HBM and only the updated rows cross NVLink).

The reference is single-device; its only scale mechanism is RowPartitionedMatrix
(partitioned-alloc.h:14-141).  Here pi/phi are node-partitioned over the G GPUs of one box
(shard s owns rows [s*ceil(N/G), (s+1)*ceil(N/G)) -- the reference's row -> (block, offset)
rule with one block per GPU) and every rank maps every peer shard through CUDA IPC, so the
kernels gather neighbor rows with NVLink peer loads.  torch.distributed (NCCL) is plumbing:
handle exchange, the [2K] beta-gradient all-reduce, the 4 perplexity sums, and the ordering
points between phases.

Work split (all static functions of (V, E_mb, H, G), so every rank computes the same plan):
  update_phi / update_pi  slot i belongs to rank (unit(i) % G), unit(i) = i % min(V, 65535) -- the
                          reference work-group that owns the slot and its Langevin RNG state
                          (phi.cc:740-747).  RNG-state ownership is therefore fixed per rank and
                          the result does not depend on G.
  beta gradient           mini-batch edges split in G contiguous chunks; partial [2K] gradients
                          are summed by an all-reduce, then every rank runs the same
                          update_theta on its replica (same RNG pool) -> replicas stay identical.
  perplexity              held-out pairs split in G contiguous chunks; 4 sums all-reduced.
  mini-batch / neighbor sampling  replicated: every rank draws the same mini-batch (same seeds)
                          and runs the (cheap, integer) neighbor sampler for all slots.
"""
import ctypes as C
import json
import os
import queue
import threading
import time

import numpy as np

MAX_GROUPS = 65535  # types.cc:537


# ------------------------------------------------------------------ the plan --
def phi_units(V, mode_wg=True, wg=32):
    """work-groups (WG modes) / work-items (THREAD) of the reference launch, phi.cc:740-747"""
    if mode_wg:
        return min(V, MAX_GROUPS)
    g = min((V + wg - 1) // wg, MAX_GROUPS)
    return g * wg


def slot_ranks(V, world, mode_wg=True, wg=32):
    """rank that processes each mini-batch slot"""
    units = phi_units(V, mode_wg, wg)
    return (np.arange(V, dtype=np.int64) % units) % world


def chunk(total, rank, world):
    """contiguous, balanced [lo, hi) split of `total` work units"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange(obj, world, group=None):
    """all-gather of a small python object (IPC handles)"""
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


_fd_round = [0]


def exchange_fds(rank, world, fds):
    """every rank hands its file descriptors to every other rank (SCM_RIGHTS over unix-domain
    sockets); returns {peer rank: [fds as valid in this process]}.  Collective."""
    import socket
    import struct
    import torch.distributed as dist
    _fd_round[0] += 1
    tag = "%s_%d" % (os.environ.get("MASTER_PORT", "0"), _fd_round[0])

    def path(r):
        return "/tmp/ammsb_fd_%s_%d.sock" % (tag, r)
    try:
        os.unlink(path(rank))
    except FileNotFoundError:
        pass
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path(rank))
    srv.listen(world)

    def serve():
        for _ in range(world - 1):
            conn, _ = srv.accept()
            conn.recv(4)
            socket.send_fds(conn, [b"x"], list(fds))
            conn.recv(1)  # the peer has the descriptors
            conn.close()
    th = threading.Thread(target=serve)
    th.start()
    dist.barrier()  # every server is listening
    out = {}
    for p in range(world):
        if p == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(path(p))
        c.send(struct.pack("i", rank))
        _, got, _, _ = socket.recv_fds(c, 16, len(fds))
        c.send(b"k")
        c.close()
        out[p] = list(got)
    th.join()
    srv.close()
    os.unlink(path(rank))
    return out


class _Buf:
    """a device pointer seen as a pyammsb buffer (torch tensor storage or a sub-range)"""

    def __init__(self, ptr, nbytes=0, keep=None):
        self.ptr, self.nbytes, self.keep = C.c_void_p(ptr), nbytes, keep


def tbuf(t):
    return _Buf(t.data_ptr(), t.numel() * t.element_size(), t)


# ---------------------------------------------------------------- the driver --
class ShardedLearner:
    """Learner::Run / HeldoutPerplexity over G node-partitioned GPUs.  `cfg` is a pymcmc.Config
    whose mini_batch_size is the GLOBAL mini-batch."""

    STREAMS = 4  # independent host sampler streams (the reference has 2: its two Samples)

    def __init__(self, cfg, rank, world, local_rank, seed=12345, prefetch=True, store_mode="partitioned",
                 collectives="peer", graph=None, shape=None):
        """cfg: a pymcmc.Config (host graph, host sets, host mini-batch strategy), or None with
        graph = a devgraph.DeviceGraph built on this rank's GPU and shape = (K, n, m): sets,
        held-out pairs, adjacency and the mini-batch strategy then live on the device"""
        import torch
        import torch.distributed as dist
        import pyammsb as A
        import pymcmc
        self.torch, self.dist, self.A = torch, dist, A
        self.cfg, self.rank, self.world, self.graph = cfg, rank, world, graph
        if graph is not None:
            K_, n_, self.m = shape
            self.p = A.make_params(graph.N, graph.E, K_, n_)
        else:
            self.p = cfg.params()
        self.N, self.K, self.n = int(self.p.N), int(self.p.K), int(self.p.num_neighbors)
        self.stream = torch.cuda.current_stream()
        self.ctx = A.Ctx(local_rank)
        self.ctx.set_stream(self.stream.cuda_stream)
        dev = torch.device("cuda", local_rank)
        # ---- pi/phi store; peers mapped through shared file descriptors ----
        #   partitioned: shard `rank` of a node-partitioned matrix, neighbor rows of other
        #                shards are read over NVLink (the layout for graphs that need G GPUs)
        #   replicated:  a full copy per GPU (when N*K*4 fits), reads are local HBM and every
        #                updated row is written to all copies over NVLink
        #   columns:     GPU g holds the columns of the reference work-items l = g (mod G) of EVERY row
        #                (csrc/cols.cu): all row reads are local HBM, partial sums cross NVLink inside
        #                the kernels, no barrier / all-reduce launches at all
        assert store_mode in ("partitioned", "replicated", "columns")
        self.store_mode = store_mode if world > 1 else "partitioned"
        self.cols = None
        # Shards are allocated with the CUDA virtual-memory API and shared as file descriptors:
        # a cudaIpc import maps peer memory with small pages and NVLink row gathers from a
        # multi-GB shard drop to 195 GB/s (735 GB/s with full-size pages, tools/peer_probe.py).
        shareable = world > 1
        if self.store_mode == "columns":
            self.store = None
        elif self.store_mode == "replicated":
            self.store = A.Store(self.ctx, self.N, self.K, 1, 0, shareable=shareable)
        else:
            self.store = A.Store(self.ctx, self.N, self.K, world, rank, shareable=shareable)
        if world > 1 and self.store is not None:
            mine = self.store.export_fds()
            for peer, (fd_pi, fd_phi) in sorted(exchange_fds(rank, world, mine).items()):
                if self.store_mode == "replicated":
                    self.store.add_mirror_fds(fd_pi, fd_phi)
                else:
                    self.store.attach_fds(peer, fd_pi, fd_phi)
                os.close(fd_pi)
                os.close(fd_phi)
            for fd in mine:
                os.close(fd)
        # ---- the exchange steps: kernels over NVLink peer memory (csrc/peer.cu), or NCCL ----
        #   "peer": cross-GPU barrier and rank-ordered all-reduce as our own kernels on a mailbox
        #           shared like the shards (deterministic sum order, no library launch latency)
        #   "nccl": torch.distributed all-reduces (baseline / fallback)
        assert collectives in ("peer", "nccl")
        self.peer = None
        if world > 1 and collectives == "peer" and self.store_mode != "columns":
            self.peer = A.Peer(self.ctx, world, rank, max(8 * self.K, 64))
            fd = self.peer.export_fd()
            for peer_rank, (pfd,) in sorted(exchange_fds(rank, world, [fd]).items()):
                self.peer.attach_fd(peer_rank, pfd)
                os.close(pfd)
            os.close(fd)
        if self.store is not None:
            self.store.init_pi(float(self.p.eta0), float(self.p.eta1))
        # ---- replicated: edge sets, theta/beta, RNG pools ----
        if graph is not None:
            self.train, self.heldout = graph.train, graph.heldout
        else:
            t_tab, t_bins, t_prime = cfg.set_table(0)
            h_tab, h_bins, h_prime = cfg.set_table(1)
            self.train = A.DevSet(self.ctx, t_tab, t_bins, t_prime)
            self.heldout = A.DevSet(self.ctx, h_tab, h_bins, h_prime)
        theta = pymcmc.init_theta_host(self.K, float(self.p.eta0), float(self.p.eta1))
        th2 = theta.reshape(self.K, 2)
        beta = (th2 / (th2[:, :1] + th2[:, 1:])).astype(np.float32).ravel()
        self.theta = torch.from_numpy(theta).to(dev)
        self.beta = torch.from_numpy(beta).to(dev)
        if graph is not None:
            self.Vmax, self.Emax = graph.max_nodes(self.m), graph.max_edges(self.m)
        else:
            self.Vmax, self.Emax = cfg.max_nodes(), cfg.max_edges()
        n = self.n
        if self.store_mode == "columns":
            H_all = graph.H if graph is not None else len(cfg.edges()[1])
            self.cols = A.Cols(self.ctx, self.N, self.K, world, rank, n, self.Vmax, self.Emax, H_all)
            fd = self.cols.export_fd()
            for peer_rank, (pfd,) in sorted(exchange_fds(rank, world, [fd]).items()):
                self.cols.attach_fd(peer_rank, pfd)
                os.close(pfd)
            os.close(fd)
            self.cols.init_pi(float(self.p.eta0), float(self.p.eta1))
            self.cols.write_theta(theta, beta)
            dist.barrier()  # every mailbox is mapped and armed before the first kernel writes into one
        self.npools = [A.Rng(self.ctx, self.Vmax * 2 * n, 56, 57) for _ in range(self.STREAMS)]
        self.ppool = A.Rng(self.ctx, self.Vmax * 32, 42, 43)
        self.bpool = A.Rng(self.ctx, self.K, 44, 45)
        # ---- mini-batch buffers ----
        self.d_nodes = torch.empty(self.Vmax, dtype=torch.int32, device=dev)
        self.d_edges = torch.empty(self.Emax, dtype=torch.int64, device=dev)
        self.d_nbs = [torch.empty(self.Vmax * n, dtype=torch.int32, device=dev) for _ in range(2)]
        self.ns_stream = torch.cuda.Stream()
        self.ctx_ns = A.Ctx(local_rank)
        self.ctx_ns.set_stream(self.ns_stream.cuda_stream)
        self.ev_ns = [torch.cuda.Event() for _ in range(2)]
        self.ev_phi = [torch.cuda.Event() for _ in range(2)]
        self.ns_seq = 0
        self.d_vec = torch.empty(self.Vmax * self.K, dtype=torch.float32, device=dev)
        self.d_sum = torch.empty(self.Vmax, dtype=torch.float32, device=dev)
        self.d_tsum = torch.empty(self.K, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(2 * self.K, dtype=torch.float32, device=dev)
        self.ws = torch.empty(self.ctx.beta_workspace_bytes(self.K), dtype=torch.uint8, device=dev)
        self.flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self.h_nodes = torch.empty(self.Vmax, dtype=torch.int32).pin_memory()
        self.h_edges = torch.empty(self.Emax, dtype=torch.int64).pin_memory()
        self.h_beta = torch.empty(2 * self.K, dtype=torch.float32).pin_memory()
        self.opts = A.PhiOpts(A.MODE_WG, 32, 0, 0, rank, world)
        # ---- held-out pairs: this rank's chunk (columns: every rank walks all pairs) ----
        if graph is not None:
            self.H = graph.H
            lo, hi = chunk(self.H, rank, world) if self.cols is None else (0, self.H)
            self.hedges = _Buf(graph.d_heldout_pairs.ptr.value + 8 * lo)
        else:
            he = cfg.edges()[1]
            self.H = len(he)
            lo, hi = chunk(self.H, rank, world) if self.cols is None else (0, self.H)
            self.d_hedges = torch.from_numpy(he[lo:hi].astype(np.int64)).to(dev) if hi > lo else \
                torch.zeros(1, dtype=torch.int64, device=dev)
            self.hedges = tbuf(self.d_hedges)
        self.H_local = hi - lo
        self.d_ppx = torch.zeros(max(self.H_local, 1), dtype=torch.float32, device=dev)
        self.pws = torch.empty(self.ctx.perplexity_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.sums = torch.zeros(4, dtype=torch.float64, device=dev)
        self.ppx_calls = 0
        self.step_count = 0
        self.edges_processed = 0
        self.h2d_bytes = 0
        # ---- host sampler ----
        # prefetch=False: every rank draws every mini-batch itself (stream t % STREAMS).
        # prefetch=True (the host path): the drawing is spread over the ranks.  Mini-batch t is
        # drawn by rank t % world on one of its LOCAL sampler streams (each with its own seed,
        # each on its own thread), copied to that rank's GPU and broadcast to the others over
        # NVLink one step ahead of its use.
        self.seeds = [C.c_uint(seed + 7919 * i) for i in range(self.STREAMS)]
        self.drawn = 0
        self.q = None
        self.t = 0
        if graph is not None:
            # device mini-batches: every rank draws the same mini-batch (same seed, same stream) on
            # its own GPU, on a sampler stream of its own, one step ahead of its use
            prefetch = False
            self.smp_stream = torch.cuda.Stream()
            self.ctx_smp = A.Ctx(local_rank)
            self.ctx_smp.set_stream(self.smp_stream.cuda_stream)
            self.dsampler = graph.sampler(self.m, ctx=self.ctx_smp)
            self.dseed = C.c_uint(seed)
            self.mb_edges = [torch.empty(self.Emax, dtype=torch.int64, device=dev) for _ in range(2)]
            self.mb_nodes = [torch.empty(self.Vmax, dtype=torch.int32, device=dev) for _ in range(2)]
            self.mb_meta = [None, None]
            self.ev_mb_free = [torch.cuda.Event() for _ in range(2)]
            self.ev_mb_ready = [torch.cuda.Event() for _ in range(2)]
            self.mb_drawn = 0
        if prefetch:
            # sampler threads per rank: the host cores divided over the ranks of the box, one left
            # for the rank's launch thread (2 .. 6)
            try:
                cores = len(os.sched_getaffinity(0))
            except Exception:
                cores = os.cpu_count() or 2
            self.LOCAL = int(os.environ.get("AMMSB_SAMPLER_THREADS", max(2, min(6, cores // max(world, 1) - 1))))
            self.local_seeds = [C.c_uint(seed + 7919 * (rank * self.LOCAL + i) + 104729) for i in range(self.LOCAL)]
            self.q = [queue.Queue(maxsize=2) for _ in range(self.LOCAL)]
            self.threads = [threading.Thread(target=self._producer, args=(i,), daemon=True)
                            for i in range(self.LOCAL)]
            for th in self.threads:
                th.start()
            self.side = torch.cuda.Stream()
            self.bgroup = dist.new_group(backend="nccl") if world > 1 else None
            self.HDR = 32
            nbytes = self.HDR + 8 * self.Emax + 4 * self.Vmax
            self.payload = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.staging = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self.h_hdr = [torch.zeros(self.HDR, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self.ev_hdr = [torch.cuda.Event() for _ in range(2)]      # header of slot b is on the host
            self.ev_payload = [torch.cuda.Event() for _ in range(2)]  # payload of slot b is on this GPU
            self.ev_free = [torch.cuda.Event() for _ in range(2)]     # kernels that read slot b are done
            self.ev_staged = [torch.cuda.Event() for _ in range(2)]   # H2D out of staging[b] is done
            self.issued = 0
        torch.cuda.synchronize()

    # ---------------------------------------------------------- sampling ----
    LOCAL = 2  # sampler streams (threads) per rank on the host path

    def _producer(self, i):
        while True:
            self.q[i].put(self.cfg.sample("Node", self.local_seeds[i]))  # ctypes call releases the GIL

    def next_minibatch(self):
        """(weight, edges, nodes) of iteration t from sampler stream t % STREAMS (prefetch=False)"""
        i = self.drawn % self.STREAMS
        self.drawn += 1
        return self.cfg.sample("Node", self.seeds[i])

    def _issue(self, t):
        """enqueue, on the side stream, everything that puts mini-batch t on every GPU"""
        torch = self.torch
        b = t & 1
        src = t % self.world
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_free[b])  # kernels of mini-batch t-2 no longer read the slot
            if self.rank == src:
                weight, edges, nodes = self.q[(t // self.world) % self.LOCAL].get()
                self.ev_staged[b].synchronize()  # the previous copy out of this staging buffer is done
                st = self.staging[b].numpy()
                V, E_mb = len(nodes), len(edges)
                st[:self.HDR].view(np.int64)[:2] = (V, E_mb)
                st[:self.HDR].view(np.float64)[2] = weight
                st[self.HDR:self.HDR + 8 * E_mb] = edges.view(np.uint8)
                off = self.HDR + 8 * self.Emax
                st[off:off + 4 * V] = nodes.view(np.uint8)
                self.payload[b].copy_(self.staging[b], non_blocking=True)
                self.ev_staged[b].record(self.side)
                self.h2d_bytes += self.HDR + 8 * E_mb + 4 * V
            if self.bgroup is not None:
                self.dist.broadcast(self.payload[b], src, group=self.bgroup)
            self.h_hdr[b].copy_(self.payload[b][:self.HDR], non_blocking=True)
            self.ev_hdr[b].record(self.side)
            self.ev_payload[b].record(self.side)
        self.issued = t + 1

    # --------------------------------------------------------- iteration ----
    def barrier(self):
        if self.peer is not None:
            self.peer.barrier()
        elif self.world > 1:
            self.dist.all_reduce(self.flag)

    def enqueue_neighbors(self, d_nodes, V, pool_index, seq, after=None):
        """neighbor sampling of mini-batch number `seq`, into neighbor buffer seq & 1, ahead of its
        use -- as the reference prepares the next Sample on its own queue (learner.cc:216-232).
        Row layouts: on the sampler stream, concurrent with the kernels of the previous mini-batch.
        Column layout: the column kernels are persistent cooperative grids that wait for their
        peers, and a concurrent kernel either starves on the SMs they leave free or holds up their
        launch (measured on 8 GPUs: 0.5 ms on the side stream for 40 us of work); the sampler of the
        NEXT mini-batch is therefore held back until update_phi of the current one has finished and
        then runs beside update_pi -- an ordinary, short kernel -- before update_beta starts."""
        if self.ns_seq >= seq:
            return
        if self.cols is not None and seq > self.step_count + 1:
            self.pending_ns = (d_nodes, V, pool_index, seq, after)  # issued by device_step, after update_phi
            self.ns_seq = seq
            return
        self._issue_neighbors(d_nodes, V, pool_index, seq, after, self.stream if self.cols is not None else self.ns_stream)
        self.ns_seq = seq

    def _issue_neighbors(self, d_nodes, V, pool_index, seq, after, ns_stream):
        b = seq & 1
        ns_stream.wait_event(self.ev_phi[b])  # update_phi of mini-batch seq-2 is done with the buffer
        if after is not None:
            ns_stream.wait_event(after)  # the nodes are on this GPU
        ns_ev = getattr(self, "ns_timing", None)
        if ns_ev is not None:
            e0 = self.torch.cuda.Event(enable_timing=True)
            e0.record(ns_stream)
        if self.cols is not None:
            # each rank draws the lists of the sampler states it owns and delivers them to every
            # rank's mailbox (third seq % 3); update_phi of step seq reads them there
            ctx = self.ctx if ns_stream is self.stream else self.ctx_ns
            self.A.cols_neighbor_sample(ctx, [self.cols], d_nodes, V, 32, seq, [self.npools[pool_index]])
        else:
            self.ctx_ns.neighbor_sample(self.npools[pool_index], d_nodes, V, self.N, self.n, 32, tbuf(self.d_nbs[b]))
        if ns_ev is not None:
            e1 = self.torch.cuda.Event(enable_timing=True)
            e1.record(ns_stream)
            ns_ev.append((V, e0, e1))
        self.ev_ns[b].record(ns_stream)

    def device_step(self, d_nodes, d_edges, V, E_mb, weight, pool_index, phi_events=None, seq=None):
        """one iteration on device-resident mini-batch buffers (pyammsb-style buffers)"""
        ctx, p, K = self.ctx, self.p, self.K
        self.step_count += 1
        seq = self.step_count if seq is None else seq
        self.enqueue_neighbors(d_nodes, V, pool_index, seq)  # no-op when it was enqueued ahead
        d_nb = self.d_nbs[seq & 1]
        self.stream.wait_event(self.ev_ns[seq & 1])
        if self.cols is not None:
            # column shards: the exchange of partial sums happens inside the kernels (mailboxes in
            # peer memory), so the iteration is four launches and no barrier or all-reduce
            A = self.A
            if phi_events is not None:
                phi_events[0].record(self.stream)
            assert seq == self.step_count  # the sampler delivered the lists under this step number
            A.cols_update_phi(ctx, [self.cols], p, self.opts, self.train, d_nodes, None, V, self.step_count,
                              [self.ppool])
            if phi_events is not None:
                phi_events[1].record(self.stream)
            self.ev_phi[seq & 1].record(self.stream)
            pend, self.pending_ns = getattr(self, "pending_ns", None), None
            if pend is not None:  # the next mini-batch's lists, beside update_pi
                self.ns_stream.wait_event(self.ev_phi[seq & 1])
                self._issue_neighbors(*pend, self.ns_stream)
            A.cols_update_pi(ctx, [self.cols], d_nodes, V, self.step_count)
            if pend is not None:
                self.stream.wait_event(self.ev_ns[pend[3] & 1])  # no ordinary kernel beside a cooperative grid
            if phi_events is not None and len(phi_events) > 2:
                phi_events[2].record(self.stream)
            A.cols_update_beta(ctx, [self.cols], p, self.train, d_edges, E_mb, weight, self.step_count, [self.bpool])
            if phi_events is not None and len(phi_events) > 2:
                phi_events[3].record(self.stream)
            self.edges_processed += E_mb
            return
        if phi_events is not None:
            phi_events[0].record(self.stream)
        ctx.update_phi(p, self.opts, tbuf(self.beta), self.store, self.train, d_nodes, tbuf(d_nb), V,
                       self.step_count, self.ppool, tbuf(self.d_vec), tbuf(self.d_sum))
        if phi_events is not None:
            phi_events[1].record(self.stream)
        self.ev_phi[seq & 1].record(self.stream)
        self.barrier()  # every read of the old pi is done before any rank writes
        ctx.update_pi_part(K, self.store, tbuf(self.d_vec), tbuf(self.d_sum), d_nodes, V, self.opts)
        self.barrier()  # every write is visible before beta reads pi
        if phi_events is not None and len(phi_events) > 2:
            phi_events[2].record(self.stream)
        lo, hi = chunk(E_mb, self.rank, self.world)
        ctx.beta_grads(p, tbuf(self.theta), tbuf(self.beta), self.store, self.train,
                       _Buf(d_edges.ptr.value + 8 * lo), hi - lo, tbuf(self.d_tsum), tbuf(self.grads), tbuf(self.ws))
        if self.peer is not None:
            self.peer.allreduce_f32(tbuf(self.grads), 2 * K)  # summed in rank order on every rank
        elif self.world > 1:
            self.dist.all_reduce(self.grads)
        ctx.update_theta(p, tbuf(self.theta), tbuf(self.beta), tbuf(self.grads), weight, self.step_count, self.bpool)
        if phi_events is not None and len(phi_events) > 2:
            phi_events[3].record(self.stream)
        self.edges_processed += E_mb

    def host_step(self):
        """one iteration from a HOST mini-batch: sampled on rank t % world, H2D there, broadcast,
        sharded kernels, D2H of beta; mini-batch t+1 travels while t is processed"""
        t = self.t
        b = t & 1
        if self.issued <= t:
            self._issue(t)
        self.ev_hdr[b].synchronize()
        hdr = self.h_hdr[b].numpy()
        V, E_mb = (int(x) for x in hdr.view(np.int64)[:2])
        weight = float(hdr.view(np.float64)[2])
        self.stream.wait_event(self.ev_payload[b])
        base = self.payload[b].data_ptr()
        d_nodes = _Buf(base + self.HDR + 8 * self.Emax)
        self.enqueue_neighbors(d_nodes, V, t % self.STREAMS, t + 1, after=self.ev_payload[b])
        self.device_step(d_nodes, _Buf(base + self.HDR), V, E_mb, weight, t % self.STREAMS, seq=t + 1)
        self.ev_free[b].record(self.stream)
        self._issue(t + 1)  # travels while the kernels of t run
        self._copy_beta_to_host()
        self.t = t + 1
        return E_mb

    def draw_device_minibatch(self, d_edges=None, d_nodes=None):
        """the next mini-batch of the device sampler stream into the given buffers (default: the
        double-buffered slots of device_graph_step); returns (weight, E_mb, V)"""
        t = self.mb_drawn
        b = t & 1
        if d_edges is None:
            self.smp_stream.wait_event(self.ev_mb_free[b])  # the kernels of mini-batch t-2 are done with the slot
            d_edges, d_nodes = tbuf(self.mb_edges[b]), tbuf(self.mb_nodes[b])
        meta = self.dsampler.sample(self.dseed, d_edges, d_nodes, ctx=self.ctx_smp)
        self.ev_mb_ready[b].record(self.smp_stream)
        self.mb_drawn = t + 1
        return meta

    def device_graph_step(self):
        """one iteration with the mini-batch drawn on the device: mini-batch t+1 is drawn (sampler
        stream) while the kernels of t run"""
        t = self.t
        b = t & 1
        if self.mb_drawn <= t:
            self.mb_meta[b] = self.draw_device_minibatch()
        weight, E_mb, V = self.mb_meta[b]
        self.h2d_bytes += 16  # the coin, the vertex and the rand_r state travel as kernel arguments
        self.stream.wait_event(self.ev_mb_ready[b])
        d_nodes, d_edges = tbuf(self.mb_nodes[b]), tbuf(self.mb_edges[b])
        self.enqueue_neighbors(d_nodes, V, t % self.STREAMS, t + 1, after=self.ev_mb_ready[b])
        self.device_step(d_nodes, d_edges, V, E_mb, weight, t % self.STREAMS, seq=t + 1)
        self.ev_mb_free[b].record(self.stream)
        self.mb_meta[1 - b] = self.draw_device_minibatch()  # blocks the host, not the compute stream
        self.enqueue_neighbors(tbuf(self.mb_nodes[1 - b]), self.mb_meta[1 - b][2], (t + 1) % self.STREAMS, t + 2,
                               after=self.ev_mb_ready[1 - b])
        self._copy_beta_to_host()
        self.t = t + 1
        return E_mb

    def _copy_beta_to_host(self):
        """the D2H read of the step's result (beta[2K]) into pinned host memory"""
        if self.cols is not None:
            _, d_beta = self.cols.beta_ptrs()
            self.A._ck(self.A.lib().ammsb_d2h_async(self.ctx.h, C.c_void_p(self.h_beta.data_ptr()), C.c_void_p(d_beta),
                                                    C.c_size_t(8 * self.K)))
        else:
            self.h_beta.copy_(self.beta, non_blocking=True)

    def run(self, iters):
        for _ in range(iters):
            if self.graph is not None:
                self.device_graph_step()
            else:
                self.host_step()

    def heldout_perplexity(self):
        self.ppx_calls += 1
        if self.cols is not None:
            avg, _ = self.A.cols_perplexity(self.ctx, [self.cols], self.p, self.heldout, self.hedges, self.H,
                                            self.ppx_calls)
            self.cols.check()
            return float(np.exp(np.float32(avg[0])))
        if self.H_local > 0:
            self.ctx.perplexity_partial(self.p, self.store, tbuf(self.beta), self.heldout, self.hedges,
                                        self.H_local, tbuf(self.d_ppx), self.ppx_calls, tbuf(self.sums), tbuf(self.pws))
        else:
            self.sums.zero_()
        if self.peer is not None:
            self.peer.allreduce_f64(tbuf(self.sums), 4)
        elif self.world > 1:
            self.dist.all_reduce(self.sums)
        s = self.sums.cpu().numpy()
        avg = (s[0] + s[1]) / (s[2] + s[3]) if (s[2] + s[3]) != 0 else 0.0  # perplexity.cc:264-273
        return float(np.exp(np.float32(-avg)))

    # ------------------------------------------------------------- state ----
    def read_local_pi(self):
        """rows this rank holds: its shard (partitioned) or the whole matrix (replicated); columns:
        all rows, with only the columns this rank owns filled in (the others are NaN)"""
        if self.cols is not None:
            out = np.full((self.N, self.K), np.nan, np.float32)
            return self.cols.read_pi(out)
        return self.store.read_pi()

    def local_rows(self):
        if self.cols is not None:
            return 0, self.N
        return self.store.first_row, self.store.first_row + self.store.local_rows

    def read_beta(self):
        if self.cols is not None:
            self.torch.cuda.synchronize()
            return self.cols.read_theta()[1]
        return self.beta.cpu().numpy()


# ------------------------------------------------------------------ bench ----
def bench_sharded(args, w, rank, world, local_rank, log, METRIC, UNIT, config, plan, ClockSampler):
    """bench.py --gpus N (N > 1): weak scaling, mini-batch = N x m edges on the same graph."""
    import torch
    import torch.distributed as dist
    import pyammsb as A
    import pymcmc
    import synth

    for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29541"), ("RANK", "0"), ("WORLD_SIZE", "1")):
        os.environ.setdefault(k, v)  # a plain `python bench.py --graph device` run on one GPU
    if plan[0] == "columns":
        # the column kernels are persistent and wait for the other ranks; an NCCL kernel (the
        # mini-batch broadcast of the host path) that waits for a peer as well must always find an
        # SM: the kernels leave 4 SMs free (csrc/cols.cu) and NCCL is held to 2 channels
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N, E, K, n = w["N"], w["E"], w["K"], w["n"]
    m = w["m"] * world
    t0 = time.time()
    mode, coll, gmode = plan  # bench.sharded_plan: a function of the arguments only
    graph = cfg = None
    if gmode == "device":
        # every rank builds the same graph, sets, held-out pairs and adjacency in its own HBM
        import devgraph
        torch.cuda.set_device(local_rank)
        gctx = A.Ctx(local_rank)
        graph = devgraph.DeviceGraph(gctx, N, E, w["heldout_ratio"], seed=1, log=log)

        def make_learner(prefetch):
            return ShardedLearner(None, rank, world, local_rank, store_mode=mode, collectives=coll, graph=graph,
                                  shape=(K, n, m))
    else:
        # rank 0 generates the synthetic edge list once; the other ranks load it
        path = os.path.join("/tmp", "ammsb_edges_%d_%d_%d.npy" % (N, E, int(os.environ.get("MASTER_PORT", "0"))))
        if rank == 0:
            np.save(path, synth.make_edges(N, E, 1))
        dist.barrier()
        keys = np.load(path)
        dist.barrier()
        if rank == 0:
            os.unlink(path)
        cfg = pymcmc.Config(K=K, mini_batch_size=m, num_node_sample=n, heldout_ratio=w["heldout_ratio"],
                            strategy="Node")
        cfg.set_graph(N, keys)

        def make_learner(prefetch):
            return ShardedLearner(cfg, rank, world, local_rank, prefetch=prefetch, store_mode=mode, collectives=coll)
    log("graph + split + sets (%s): %.1fs" % (gmode, time.time() - t0))

    # ---- parity of the sharded run against ONE GPU, in the run itself: one iteration on the same
    #      (non-link) mini-batch from the same initial state, pi compared bit for bit on the rows /
    #      columns this rank holds (results must not depend on the GPU count) ----
    parity = "skipped: a full copy of pi (%.0f GB) beside the shard does not fit one GPU" % (4.0 * N * K / 1e9)
    if world > 1 and 4.0 * N * K <= 60e9 and not os.environ.get("AMMSB_BENCH_NO_PARITY"):
        os.environ["AMMSB_PHI_NOSPLIT"] = "1"  # one association of the gradient sum on both sides
        try:
            pl = make_learner(False)
            if graph is not None:
                single = ShardedLearner(None, 0, 1, local_rank, graph=graph, shape=(K, n, m))
            else:
                single = ShardedLearner(cfg, 0, 1, local_rank, prefetch=False)
            mb = None
            for _ in range(8):  # the first non-link mini-batch of the stream
                if graph is not None:
                    d_e = pl.ctx.buf(np.uint64, pl.Emax)
                    d_n = pl.ctx.buf(np.uint32, pl.Vmax)
                    wgt, E_mb, V = pl.draw_device_minibatch(d_e, d_n)
                    torch.cuda.synchronize()
                else:
                    wgt, edges, nodes = pl.next_minibatch()
                    E_mb, V = len(edges), len(nodes)
                    d_e, d_n = pl.ctx.from_host(edges), pl.ctx.from_host(nodes)
                if V > 1024:
                    mb = (wgt, E_mb, V, d_e, d_n)
                    break
                d_e.free(); d_n.free()
            if mb is None:
                parity = "skipped: no non-link mini-batch among the first 8"
            else:
                wgt, E_mb, V, d_e, d_n = mb
                for L in (pl, single):
                    L.device_step(d_n, d_e, V, E_mb, wgt, 0)
                torch.cuda.synchronize()
                dist.barrier()
                nodes_h = d_n.read()[:V].astype(np.int64)
                a_pi, b_pi = pl.read_local_pi(), single.read_local_pi()
                lo, hi = pl.local_rows()
                b_pi = b_pi[lo:hi]
                own = ~np.isnan(a_pi)
                same = bool(np.array_equal(a_pi[own], b_pi[own]))
                changed = int((b_pi[nodes_h[(nodes_h >= lo) & (nodes_h < hi)] - lo] != 0).any(axis=1).sum())
                ok = torch.tensor([1.0 if same and changed > 0 else 0.0], device="cuda")
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                parity = ("ok: pi after one iteration (mini-batch of %d edges, %d nodes) bit-identical to the same "
                          "iteration on one GPU, on every rank" % (E_mb, V)) if float(ok[0]) == 1.0 else "FAILED"
                d_e.free(); d_n.free()
            del pl, single
            torch.cuda.empty_cache()
        finally:
            del os.environ["AMMSB_PHI_NOSPLIT"]
        log("parity vs one GPU: %s" % parity)
        if parity == "FAILED":
            raise SystemExit("bench.py: the sharded run differs from the one-GPU run")
    lrn = make_learner(False)
    stream = lrn.stream
    ctx = lrn.ctx

    # ---- value leg: pre-sampled mini-batches resident in HBM ----
    # AMMSB_BENCH_VARIANTS="name:ENV=VAL,ENV2=VAL;name2:..." (development): after the contract run, the
    # same number of steps again under each variant's environment, in the same process (the kernel
    # switches are read at every launch) -- one set-up for several A/B measurements
    variants = []
    for item in filter(None, os.environ.get("AMMSB_BENCH_VARIANTS", "").split(";")):
        name, _, kv = item.partition(":")
        variants.append((name, dict(x.split("=", 1) for x in kv.split(",") if x)))
    total = args.warmup + args.steps * (1 + len(variants))
    t0 = time.time()
    if graph is not None:
        # (weight, E_mb, V) per mini-batch, drawn by the device sampler into one resident buffer
        d_edges_all = ctx.buf(np.uint64, total * lrn.Emax)
        d_nodes_all = ctx.buf(np.uint32, total * lrn.Vmax)
        e_off = np.arange(total + 1, dtype=np.int64) * lrn.Emax
        v_off = np.arange(total + 1, dtype=np.int64) * lrn.Vmax
        batches = []
        for i in range(total):
            wgt, E_mb, V = lrn.draw_device_minibatch(_Buf(d_edges_all.ptr.value + 8 * int(e_off[i])),
                                                     _Buf(d_nodes_all.ptr.value + 4 * int(v_off[i])))
            batches.append((wgt, E_mb, V))
        torch.cuda.synchronize()
        log("device sampler: %d mini-batches of up to %d edges in %.2fs" % (total, m, time.time() - t0))
    else:
        drawn = [lrn.next_minibatch() for _ in range(total)]
        log("host sampler: %d mini-batches of up to %d edges in %.1fs" % (total, m, time.time() - t0))
        e_off = np.cumsum([0] + [len(b[1]) for b in drawn])
        v_off = np.cumsum([0] + [len(b[2]) for b in drawn])
        d_edges_all = ctx.from_host(np.concatenate([b[1] for b in drawn]))
        d_nodes_all = ctx.from_host(np.concatenate([b[2] for b in drawn]))
        batches = [(b[0], len(b[1]), len(b[2])) for b in drawn]
        del drawn

    def nodes_of(i):
        return _Buf(d_nodes_all.ptr.value + 4 * int(v_off[i]))

    def step(i, ev=None):
        wgt, E_mb, V = batches[i]
        lrn.enqueue_neighbors(nodes_of(i), V, i % lrn.STREAMS, i + 1)
        if i + 1 < total:  # the next mini-batch's neighbors are drawn while this one is processed
            lrn.enqueue_neighbors(nodes_of(i + 1), batches[i + 1][2], (i + 1) % lrn.STREAMS, i + 2)
        lrn.device_step(nodes_of(i), _Buf(d_edges_all.ptr.value + 8 * int(e_off[i])), V, E_mb, wgt,
                        i % lrn.STREAMS, ev, seq=i + 1)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()  # sampled from before the warm-up to the end of the timed region
    for i in range(args.warmup):
        step(i)
    # AMMSB_STAGE_EVENTS=1: events after update_pi and update_beta as well (per-stage times of the run itself)
    nev = 4 if os.environ.get("AMMSB_STAGE_EVENTS") else 2
    evs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(nev)) for _ in range(args.steps)]
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = A.launch_count()
    if nev == 4:
        lrn.ns_timing = []
    e_start.record(stream)
    for k in range(args.steps):
        step(args.warmup + k, evs[k])
    e_stop.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    launches = A.launch_count() - launches0
    main_ns_timing = list(getattr(lrn, "ns_timing", None) or [])
    t = torch.tensor([e_start.elapsed_time(e_stop), sum(e[0].elapsed_time(e[1]) for e in evs)],
                     dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
    dev_ms, phi_ms = float(t[0]), float(t[1])
    clk = clocks.stop() if rank == 0 else None
    timed = batches[args.warmup:args.warmup + args.steps]
    edges_timed = int(sum(b[1] for b in timed))
    value = edges_timed / (dev_ms * 1e-3)
    Vs = [b[2] for b in timed]
    nvlink_peak = 770.0  # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    local_bytes = float(sum((V / world) * ((n + 2) * 4 * K + 68 * n + 8) for V in Vs))  # per GPU
    if mode == "columns":
        # every GPU walks ALL slots on its K/G columns: local HBM bytes per GPU are those of the
        # one-GPU kernel at the global mini-batch divided by G; NVLink carries 4 bytes per
        # (slot, neighbor, peer)
        col_bytes = float(sum(V * ((n + 2) * 4 * K / world + 68 * n + 8) for V in Vs))
        ach = col_bytes / (phi_ms * 1e-3) / 1e9
        out_bytes = float(sum(V * (n + 1) * 4 * (world - 1) for V in Vs))
        roofline = {"bound": "hbm", "kernel": "k_cols_phi2 (per GPU, column-sharded pi)", "unit": "GB/s",
                    "achieved": round(ach, 1), "peak": hbm_peak, "frac": round(ach / hbm_peak, 4),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)", "traffic": None,
                    "share_of_step": round(phi_ms / dev_ms, 4),
                    "nvlink_outbound_GB_per_gpu_per_step": round(out_bytes / args.steps / 1e9, 4),
                    "nvlink_outbound_GBps": round(out_bytes / (phi_ms * 1e-3) / 1e9, 1)}
    elif mode == "partitioned":
        # update_phi is bound by the neighbor rows that cross the switch into each GPU
        remote_bytes = float(sum((V / world) * n * 4 * K * (world - 1) / world for V in Vs))
        ach = remote_bytes / (phi_ms * 1e-3) / 1e9
        roofline = {"bound": "nvlink", "kernel": "k_update_phi_fast (NVLink peer loads of neighbor rows)",
                    "unit": "GB/s", "achieved": round(ach, 1), "peak": nvlink_peak, "frac": round(ach / nvlink_peak, 4),
                    "peak_source": "770 GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)",
                    "traffic": None, "inbound_remote_GB_per_gpu_per_step": round(remote_bytes / args.steps / 1e9, 4),
                    "hbm_side_GBps": round(local_bytes / (phi_ms * 1e-3) / 1e9, 1),
                    "share_of_step": round(phi_ms / dev_ms, 4)}
    else:
        # every read is local HBM; NVLink only carries the updated rows (update_pi peer stores)
        ach = local_bytes / (phi_ms * 1e-3) / 1e9
        out_bytes = float(sum((V / world) * (world - 1) * (4 * K + 4) for V in Vs))
        roofline = {"bound": "hbm", "kernel": "k_update_phi_fast (per GPU, replicated pi)", "unit": "GB/s",
                    "achieved": round(ach, 1), "peak": hbm_peak, "frac": round(ach / hbm_peak, 4),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)", "traffic": None,
                    "share_of_step": round(phi_ms / dev_ms, 4),
                    "nvlink_outbound_GB_per_gpu_per_step": round(out_bytes / args.steps / 1e9, 4)}
    roofline["store"] = mode
    timed = timed[:args.steps]
    Vs = Vs[:args.steps]

    def stage_means(evs, Vs, ns_timing):
        # mean ms per stage over the timed steps of this rank, non-link and link mini-batches apart;
        # "gap" = from the end of update_beta to the start of the next update_phi on the stream
        out = {}
        for kind, sel in (("non_link", lambda V: V > 1024), ("link", lambda V: V <= 1024)):
            ks = [k for k in range(len(evs)) if sel(Vs[k])]
            if not ks:
                continue
            d = {"steps": len(ks),
                 "update_phi": float(np.mean([evs[k][0].elapsed_time(evs[k][1]) for k in ks])),
                 "update_pi": float(np.mean([evs[k][1].elapsed_time(evs[k][2]) for k in ks])),
                 "update_beta": float(np.mean([evs[k][2].elapsed_time(evs[k][3]) for k in ks]))}
            ns = [a.elapsed_time(b) for V, a, b in ns_timing if sel(V)]
            d["neighbor_sample"] = float(np.mean(ns)) if ns else 0.0
            gaps = [evs[k][3].elapsed_time(evs[k + 1][0]) for k in ks if k + 1 < len(evs)]
            d["gap_to_next"] = float(np.mean(gaps)) if gaps else 0.0
            out[kind] = {a: round(b, 4) if isinstance(b, float) else b for a, b in d.items()}
        return out

    variant_results = {}
    for vi, (vname, venv) in enumerate(variants):
        saved = {k: os.environ.get(k) for k in venv}
        os.environ.update(venv)
        first = args.warmup + args.steps * (1 + vi)
        v_evs = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(nev)) for _ in range(args.steps)]
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        if nev == 4:
            lrn.ns_timing = []
        v0.record(stream)
        for k in range(args.steps):
            step(first + k, v_evs[k])
        v1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        tv = torch.tensor([v0.elapsed_time(v1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        vb = batches[first:first + args.steps]
        res = {"env": venv, "ms_per_step": float(tv[0]) / args.steps,
               "value": int(sum(b[1] for b in vb)) / (float(tv[0]) * 1e-3)}
        if nev == 4:
            res["stages_in_run_ms"] = stage_means(v_evs, [b[2] for b in vb], lrn.ns_timing)
        variant_results[vname] = res
        for k, v in saved.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    stages_in_run = None
    if nev == 4:
        stages_in_run = stage_means(evs, Vs, main_ns_timing)

    # ---- perplexity (sharded) ----
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ppx = lrn.heldout_perplexity()
    torch.cuda.synchronize()
    ppx_s = time.perf_counter() - t1

    # ---- e2e leg: host mini-batches through the sharded driver ----
    e2e = None
    if not args.no_e2e:
        if graph is None and os.environ.get("AMMSB_E2E_DEVICE_SAMPLER"):
            # the host-built graph is put on the device once; the Node mini-batches are then drawn
            # there in the reference's order (the host draws the coin and the vertex)
            import devgraph
            up = devgraph.UploadedGraph(A.Ctx(local_rank), cfg, log=log)
            l2 = ShardedLearner(None, rank, world, local_rank, store_mode=mode, collectives=coll, graph=up,
                                shape=(K, n, m))
            e2e_api = ("dist.ShardedLearner.device_graph_step() over the HOST-built graph (the reference's split, "
                       "cuckoo tables and Graph, uploaded once): coin and vertex drawn on the host (rand_r), the Node "
                       "mini-batch on every GPU by the device sampler in the reference's order (bit-identical to the "
                       "host strategy), one step ahead (D2H of its 16-byte header), sharded kernels, D2H of beta")
        else:
            l2 = make_learner(True)
            e2e_api = None
        l2.run(args.warmup)
        dist.barrier()
        torch.cuda.synchronize()
        b0, e0 = l2.h2d_bytes, l2.edges_processed
        t0 = time.perf_counter()
        l2.run(args.steps)
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt[0])
        hb = torch.tensor([l2.h2d_bytes - b0], dtype=torch.float64, device="cuda")
        dist.all_reduce(hb)  # each rank counts the mini-batches it drew and copied
        e2e = {"value": (l2.edges_processed - e0) / dt, "unit": UNIT,
               "h2d_bytes_per_step": float(hb[0]) / args.steps, "d2h_bytes_per_step": (8 * K + 32) * world,
               "iterations_per_s": args.steps / dt, "ms_per_step": 1e3 * dt / args.steps,
               "api": e2e_api if e2e_api is not None else
                      ("dist.ShardedLearner.host_step(): mini-batch t drawn on rank t % N (2 sampler threads per "
                       "rank), H2D from pinned memory there, NCCL broadcast one step ahead, sharded kernels + "
                       "all-reduces, D2H of beta").replace("2 sampler threads", "%d sampler threads" % l2.LOCAL) if graph is None else
                      ("dist.ShardedLearner.device_graph_step(): coin and vertex drawn on the host (rand_r), the "
                       "mini-batch on every GPU by the device sampler one step ahead (D2H of its 16-byte header), "
                       "sharded kernels + all-reduces, D2H of beta; the graph, the cuckoo sets and the adjacency "
                       "were built in HBM, nothing but kernel arguments goes host to device")}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "iterations_per_s": args.steps / (dev_ms * 1e-3), "perplexity_eval_s": ppx_s, "heldout_perplexity": ppx,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": None,
            "parity_vs_n1": parity,
        }
        if stages_in_run is not None:
            line["stages_in_run_ms"] = stages_in_run
        if variant_results:
            line["variants"] = variant_results
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
