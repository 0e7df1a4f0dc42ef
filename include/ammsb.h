/*
 * ammsb.h -- C ABI of the B200-native SG-MCMC a-MMSB hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, plain pointers and
 * sizes.  Each entry point replaces one operator of the reference
 * (ielhelw/mcmc-ammsb-gpu); the reference interface it stands in for is cited
 * as file:line relative to the reference tree.  The reference has no FFI layer
 * of its own -- its boundary is the C++ API of libmcmc (mcmc::Learner et al.),
 * which mcmc-ammsb-gpu_b200/host/ re-implements on top of this header.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message of
 *     the last failure on the calling thread is ammsb_last_error().
 *   - no exceptions cross the boundary; no ownership of host pointers is taken.
 *   - "d_" parameters are device pointers valid on the context's device
 *     (ammsb_malloc or any CUDA allocation, e.g. a torch tensor's data_ptr).
 *   - operators are enqueued on the context's stream and return without
 *     waiting; ammsb_ctx_sync() is the reference's queue.Finish().
 *   - there is NO CPU fallback: without a CUDA device every call fails.
 *   - handles may be used from two host threads at once only on disjoint
 *     handles/contexts (the reference's sampler thread, learner.cc:216-229).
 */
#ifndef AMMSB_H_
#define AMMSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMMSB_MAX_SHARDS 8
#define AMMSB_IPC_HANDLE_BYTES 64

typedef struct ammsb_ctx ammsb_ctx;     /* clcuda::Context + clcuda::Queue        */
typedef struct ammsb_rng ammsb_rng;     /* random::OpenClRandom   (random.h:20-42) */
typedef struct ammsb_set ammsb_set;     /* cuckoo::OpenClSet      (cuckoo.h:71-86) */
typedef struct ammsb_store ammsb_store; /* RowPartitionedMatrix pi[N,K] + phi[N]
                                           (partitioned-alloc.h:73-141, learner.h:52-53) */

/* Hyper-parameters that the reference bakes into its kernels as -D flags
 * (MakeCompileFlags, config.cc:66-83).  The float members must already have
 * gone through ammsb_round_param(), which reproduces the "%e" text rounding of
 * config.cc:57-64. */
typedef struct {
  uint64_t N;             /* -DN                */
  uint64_t E;             /* -DE                */
  uint32_t K;             /* -DK                */
  uint32_t num_neighbors; /* -DNUM_NEIGHBORS    */
  float alpha;            /* -DALPHA            */
  float a, b, c;          /* -DEPS_A/B/C        */
  float epsilon;          /* -DEPSILON          */
  float eta0, eta1;       /* -DETA0/1           */
} ammsb_params;

/* Work-item mapping of the reference launch whose RNG stream / summation
 * association is reproduced (Config::phi_mode, phi_wg_size, phi_vector_width,
 * phi_disable_noise; config.h:55-66). */
enum { AMMSB_MODE_THREAD = 0, AMMSB_MODE_WG = 1 };
typedef struct {
  uint32_t mode;          /* AMMSB_MODE_*; WG covers WG-NAIVE/SHARED/CODE_GEN */
  uint32_t wg;            /* reference work-group size (state pool stride)    */
  uint32_t disable_noise; /* PHI_RANDN -> literal 1 (phi.cc:673-677)           */
  uint32_t strict;        /* 1: IEEE, reference-association kernel (slow);
                             0: production kernel (HBM-roofline path)          */
  /* multi-GPU: this call processes only the mini-batch slots whose RNG unit u (the
   * reference work-group / work-item that owns the slot, phi.cc:740-747) satisfies
   * u % part_count == part_index.  part_count 0 or 1 = every slot.  The ownership of
   * RNG state by rank is therefore static and results do not depend on the GPU count. */
  uint32_t part_index;
  uint32_t part_count;
} ammsb_phi_opts;

const char* ammsb_last_error(void);
const char* ammsb_version(void);
float ammsb_round_param(float f); /* config.cc:57-64 float_to_string */
float ammsb_eps_t(const ammsb_params* p, uint32_t step_count); /* learner.cc:41-43 */

/* ---- context: clcuda::Platform/Device/Context/Queue (main.cc:17-20,99-101) ---- */
int ammsb_device_count(int* count);
int ammsb_ctx_create(int device, ammsb_ctx** out);
int ammsb_ctx_destroy(ammsb_ctx* ctx);
int ammsb_ctx_sync(ammsb_ctx* ctx);                    /* Queue::Finish()            */
int ammsb_ctx_set_stream(ammsb_ctx* ctx, void* cuda_stream); /* borrow a cudaStream_t */
int ammsb_ctx_device(const ammsb_ctx* ctx, int* device);
int ammsb_ctx_device_name(const ammsb_ctx* ctx, char* buf, size_t len); /* Device::Name() */
int ammsb_ctx_sm_count(const ammsb_ctx* ctx, int* sms);

/* ---- buffers: clcuda::Buffer<T> ctor / Read / Write / CopyTo ---- */
int ammsb_malloc(ammsb_ctx* ctx, size_t bytes, void** d_ptr);
int ammsb_free(ammsb_ctx* ctx, void* d_ptr);
int ammsb_memset(ammsb_ctx* ctx, void* d_ptr, int value, size_t bytes);
int ammsb_h2d(ammsb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes); /* synchronous */
int ammsb_d2h(ammsb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes); /* synchronous */
int ammsb_h2d_async(ammsb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int ammsb_d2h_async(ammsb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int ammsb_d2d(ammsb_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
int ammsb_host_alloc(size_t bytes, void** h_ptr); /* pinned */
int ammsb_host_free(void* h_ptr);

/* ---- timing: clcuda::Event::GetElapsedTime() (phi.cc:755-762 ...) ---- */
int ammsb_timer_start(ammsb_ctx* ctx);
int ammsb_timer_stop_ms(ammsb_ctx* ctx, float* ms); /* waits for the stop event */

/* ---- completion markers: clcuda::Event as a synchronisation point (Kernel::Launch's last
 *      argument).  record = "everything enqueued on ctx so far"; sync waits for it on the host
 *      without draining later work, which lets a caller keep the stream fed. ---- */
typedef struct ammsb_event ammsb_event;
int ammsb_event_create(ammsb_ctx* ctx, ammsb_event** out);
int ammsb_event_record(ammsb_ctx* ctx, ammsb_event* ev);
int ammsb_event_sync(ammsb_event* ev);
int ammsb_event_destroy(ammsb_event* ev);

/* ---- RNG pool: OpenClRandomFactory::CreateRandom(size, seed) (random.h:44-58),
 *      RandomInit kernel (random.cc:31-44): state[i] = (seed_x + i, seed_y + i) ---- */
int ammsb_rng_create(ammsb_ctx* ctx, uint64_t num_states, uint64_t seed_x,
                     uint64_t seed_y, ammsb_rng** out);
int ammsb_rng_destroy(ammsb_rng* rng);
int ammsb_rng_size(const ammsb_rng* rng, uint64_t* num_states);
int ammsb_rng_get_state(ammsb_rng* rng, uint64_t* h_xy /* [2*num_states] */); /* Serialize */
int ammsb_rng_set_state(ammsb_rng* rng, const uint64_t* h_xy);                /* Parse     */
/* test/diagnostic streams (random-test.cc:47-99): every state draws `draws` values */
int ammsb_rng_draw_u64(ammsb_rng* rng, uint32_t draws, uint64_t* h_out /* [num_states*draws] */);
int ammsb_rng_draw_randn(ammsb_rng* rng, uint32_t draws, float* h_out);
int ammsb_rng_draw_gamma(ammsb_rng* rng, uint32_t draws, float a, float b, float* h_out);

/* ---- cuckoo edge set: OpenClSetFactory::CreateSet(const Set&) (cuckoo.h:88-103),
 *      device Set_HasEdge (cuckoo.cc:39-65).  `table` is Set::Serialize()
 *      (cuckoo.cc:211-220): [2 buckets][num_bins][4 slots] u64, empty = ~0. ---- */
int ammsb_set_create(ammsb_ctx* ctx, const uint64_t* h_table, uint64_t num_bins,
                     uint32_t prime_idx, ammsb_set** out);
int ammsb_set_destroy(ammsb_set* set);
int ammsb_set_has(ammsb_set* set, const uint64_t* h_keys, uint64_t n, uint8_t* h_out);
int ammsb_set_has_device(ammsb_set* set, const uint64_t* d_keys, uint64_t n, uint8_t* d_out);

/* On-device build of the same set (cuckoo::Set::SetContents, cuckoo.cc:117-161): table geometry
 * (bins = 1 + ceil(1.15 n / 8), cuckoo.cc:98-104) and hash functions are the host's, so the
 * lookup is unchanged; keys are placed in parallel, so the slot a key ends up in differs from the
 * host's random walk while membership is the same.  Falls back through the four hash-constant
 * pairs like the host build and fails with "Failed to insert into the cuckoo set" after that. */
int ammsb_set_build(ammsb_ctx* ctx, const uint64_t* h_keys, uint64_t n, ammsb_set** out);
int ammsb_set_build_device(ammsb_ctx* ctx, const uint64_t* d_keys, uint64_t n, ammsb_set** out);
int ammsb_set_info(const ammsb_set* set, uint64_t* num_bins, uint32_t* prime_idx);
int ammsb_set_read_table(ammsb_set* set, uint64_t* h_table /* [8 * num_bins] */); /* Set::Serialize() */

/* ---- graph inputs built in HBM (the reference builds them on the host: data.cc:12-128).
 *      ammsb_graph_generate: the synthetic graphs of the named shapes -- E distinct undirected
 *      pairs u < v with uniform endpoints, in a pseudo-random order, a function of (N, E, seed).
 *      ammsb_graph_nonlinks: `count` distinct pairs that are in neither set (the held-out
 *      non-links of GenerateSetsFromEdges, data.cc:110-126; b may be NULL).
 *      ammsb_graph_csr: mcmc::Graph (data.cc:12-25) -- d_offsets [N+1], d_adj [2E] (neighbors of v
 *      ascending), d_degree [N] (may be NULL). ---- */
int ammsb_graph_generate(ammsb_ctx* ctx, uint64_t N, uint64_t E, uint64_t seed, uint64_t* d_edges);
int ammsb_graph_nonlinks(ammsb_ctx* ctx, uint64_t N, uint64_t count, uint64_t seed, ammsb_set* a,
                         ammsb_set* b, uint64_t* d_out);
int ammsb_graph_csr(ammsb_ctx* ctx, uint64_t N, const uint64_t* d_edges, uint64_t E, uint64_t* d_offsets,
                    uint32_t* d_adj, uint32_t* d_degree);

/* ---- the Node mini-batch strategy on the device (sampleNodeLink / sampleNodeNonLink,
 *      sample.cc:253-293, + ExtractNodesFromMiniBatch, learner.cc:162-173) for graphs whose host
 *      copy is impractical.  The caller draws the coin and the vertex u with rand_r as sampleNode
 *      does.  ammsb_minibatch_nonlink examines the same rand_r candidate stream as the host
 *      strategy (*seed = state after u was drawn), refuses, drops and stops like it, and leaves
 *      *seed where the host leaves it: the mini-batch has the same edges and nodes; they are
 *      emitted in draw order (nodes: u first), not in std::unordered_set order.  It waits for the
 *      stream (the draw count decides the next seed).  d_edges [m], d_nodes [m + 1].
 *      ammsb_minibatch_link emits every training edge of u (degree > 0): d_edges [degree],
 *      d_nodes [degree + 1]; asynchronous. ---- */
typedef struct ammsb_sampler ammsb_sampler;
int ammsb_sampler_create(ammsb_ctx* ctx, uint64_t N, uint32_t mini_batch_size, ammsb_sampler** out);
int ammsb_sampler_destroy(ammsb_sampler* sampler);
int ammsb_minibatch_nonlink(ammsb_sampler* sampler, ammsb_ctx* ctx, uint32_t u, unsigned int* seed,
                            ammsb_set* train, ammsb_set* heldout, uint64_t* d_edges, uint32_t* d_nodes,
                            uint32_t* num_edges, uint32_t* num_nodes);
int ammsb_minibatch_link(ammsb_ctx* ctx, uint32_t u, uint32_t degree, const uint64_t* d_offsets,
                         const uint32_t* d_adj, uint64_t* d_edges, uint32_t* d_nodes);

/* ---- the reference's emission ORDER on the device.  The reference emits a mini-batch in the
 *      iteration order of libstdc++'s std::unordered_set (edges: sample.cc:267,290; nodes:
 *      ExtractNodesFromMiniBatch, learner.cc:162-173).  ammsb_orderset_apply turns keys in insertion
 *      order (no duplicates; std::hash of an integer is the identity) into that iteration order;
 *      ammsb_minibatch_finish is the tail of a device strategy: d_edges (E edges in insertion
 *      order -- what ammsb_minibatch_nonlink / _link produce) are reordered in place, d_nodes
 *      receives the endpoints' unordered_set<Vertex> order, *num_nodes their count (waits for the
 *      stream).  With the adjacency in the host Graph's order (data.cc:12-25) the Node strategy on
 *      the device is then bit-identical to sampleNode + ExtractNodesFromMiniBatch.
 *      max_keys >= 2 * the largest mini-batch (the endpoint sequence has 2 E entries). ---- */
typedef struct ammsb_orderset ammsb_orderset;
int ammsb_orderset_create(ammsb_ctx* ctx, uint32_t max_keys, ammsb_orderset** out);
int ammsb_orderset_destroy(ammsb_orderset* os);
int ammsb_orderset_apply(ammsb_orderset* os, ammsb_ctx* ctx, const uint64_t* d_keys, uint32_t n, uint64_t* d_out);
int ammsb_minibatch_finish(ammsb_orderset* os, ammsb_ctx* ctx, uint64_t* d_edges, uint32_t E, uint32_t* d_nodes,
                           uint32_t* num_nodes);

/* ---- pi/phi store: RowPartitionedMatrixFactory<Float>::CreateMatrix(rows, cols)
 *      (partitioned-alloc.h:152-157) + the phi[N] buffer (learner.cc:83).
 *      Node-partitioned over `num_shards` GPUs: shard s owns rows
 *      [s*rows_per_shard, (s+1)*rows_per_shard), rows_per_shard = ceil(N/num_shards)
 *      (the reference's row -> (block, offset) rule, partitioned-alloc.h:24-28).
 *      A process owns shard `shard_id`; peers are attached by CUDA IPC handle. ---- */
int ammsb_store_create(ammsb_ctx* ctx, uint64_t N, uint32_t K, uint32_t num_shards,
                       uint32_t shard_id, ammsb_store** out);
int ammsb_store_destroy(ammsb_store* store);
int ammsb_store_export(ammsb_store* store, uint8_t* pi_handle /* [64] */, uint8_t* phi_handle /* [64] */);
int ammsb_store_attach(ammsb_store* store, uint32_t shard, const uint8_t* pi_handle, const uint8_t* phi_handle);
/* single-process multi-device: attach another store's shard by direct peer access */
int ammsb_store_attach_local(ammsb_store* store, uint32_t shard, ammsb_store* peer);
/* Shareable stores.  ammsb_store_create_shareable allocates pi/phi with the CUDA virtual-memory
 * API so that another process maps them with full-size pages: a cudaIpc import maps peer memory
 * with small pages, and NVLink row gathers from a multi-GB shard then run at 195 GB/s instead of
 * 735 GB/s (measured, tools/peer_probe.py).  export_fd yields two POSIX file descriptors (pi,
 * phi) to be passed to the peer process (SCM_RIGHTS); attach_fd / add_mirror_fd are the
 * counterparts of ammsb_store_attach / ammsb_store_add_mirror.  The caller closes the fds. */
int ammsb_store_create_shareable(ammsb_ctx* ctx, uint64_t N, uint32_t K, uint32_t num_shards,
                                 uint32_t shard_id, ammsb_store** out);
int ammsb_store_export_fd(ammsb_store* store, int* pi_fd, int* phi_fd);
int ammsb_store_attach_fd(ammsb_store* store, uint32_t shard, int pi_fd, int phi_fd);
int ammsb_store_add_mirror_fd(ammsb_store* store, int pi_fd, int phi_fd);

/* Replicated mode (pi fits on every GPU): each GPU holds a full copy (num_shards = 1) and
 * registers the other GPUs' copies as mirrors; reads stay in local HBM and ammsb_update_pi*
 * writes every updated row to the local copy and to all mirrors (NVLink peer stores). */
int ammsb_store_add_mirror(ammsb_store* store, const uint8_t* pi_handle, const uint8_t* phi_handle);
int ammsb_store_add_mirror_local(ammsb_store* store, ammsb_store* peer);
int ammsb_store_rows(const ammsb_store* store, uint64_t* first_row, uint64_t* num_rows);
int ammsb_store_local_ptrs(ammsb_store* store, float** d_pi, float** d_phi);
/* make the local shard use caller-owned memory for phi (>= rows_per_shard floats): the
 * reference keeps phi in a Buffer of its own (learner.h:53) that callers read directly */
int ammsb_store_bind_phi(ammsb_store* store, float* d_phi);
/* host access to locally-owned rows (global row numbering) */
int ammsb_store_write_pi(ammsb_store* store, uint64_t row0, uint64_t nrows, const float* h_src);
int ammsb_store_read_pi(ammsb_store* store, uint64_t row0, uint64_t nrows, float* h_dst);
int ammsb_store_write_phi(ammsb_store* store, uint64_t row0, uint64_t nrows, const float* h_src);
int ammsb_store_read_phi(ammsb_store* store, uint64_t row0, uint64_t nrows, float* h_dst);
/* RandomGammaAndNormalize (random.cc:159-167, learner.cc:154-155): pi ~ Gamma(eta0,eta1)
 * row-normalised, phi = row sums; stream = pool N*32 seeded {11,113}, group per row. */
int ammsb_store_init_pi(ammsb_store* store, float eta0, float eta1);

/* ---- NeighborSampler::operator()(num_samples, nodes) (sample.h:16-28, sample.cc:48-121).
 *      Bit-exact with the reference stream for the given work-group size `wg`
 *      (Config::neighbor_sampler_wg_size).  d_hash_out (may be NULL) receives the
 *      per-slot open-addressing tables [V, 2n] (NeighborSampler::GetHash()). ---- */
int ammsb_neighbor_sample(ammsb_ctx* ctx, ammsb_rng* pool, const uint32_t* d_nodes,
                          uint32_t V, uint32_t N, uint32_t n, uint32_t wg,
                          uint32_t* d_neighbors /* [V, n] */, uint32_t* d_hash_out);

/* ---- PhiUpdater::operator()(nodes, neighbors, V) (phi.h:20-23, phi.cc:728-763),
 *      split at the reference's own kernel boundary: update_phi writes phi_vec
 *      (and the row sums), update_pi overwrites pi/phi for the mini-batch nodes. ---- */
int ammsb_update_phi(ammsb_ctx* ctx, const ammsb_params* p, const ammsb_phi_opts* opts,
                     const float* d_beta /* [2K] */, ammsb_store* store, ammsb_set* train,
                     const uint32_t* d_nodes /* [V] */, const uint32_t* d_neighbors /* [V,n] */,
                     uint32_t V, uint32_t step_count, ammsb_rng* pool,
                     float* d_phi_vec /* [V,K] */, float* d_phi_sum /* [V] */);
int ammsb_update_pi(ammsb_ctx* ctx, uint32_t K, ammsb_store* store,
                    const float* d_phi_vec, const float* d_phi_sum,
                    const uint32_t* d_nodes, uint32_t V);
/* the slots of one rank only (same partition rule as ammsb_update_phi; rows owned by
 * another GPU are written with NVLink peer stores) */
int ammsb_update_pi_part(ammsb_ctx* ctx, uint32_t K, ammsb_store* store,
                         const float* d_phi_vec, const float* d_phi_sum,
                         const uint32_t* d_nodes, uint32_t V, const ammsb_phi_opts* opts);

/* ---- BetaUpdater::operator()(edges, E_mb, scale) (beta.h:25, beta.cc:334-384).
 *      ammsb_beta_grads = sum_theta + calculate_grads_partial + sum_grads;
 *      ammsb_update_theta = update_theta + theta->beta copy + row normalise.
 *      In a multi-GPU run the [2K] gradient is all-reduced between the two. ---- */
int ammsb_beta_workspace_bytes(ammsb_ctx* ctx, uint32_t K, size_t* bytes);
int ammsb_beta_grads(ammsb_ctx* ctx, const ammsb_params* p, const float* d_theta,
                     const float* d_beta, ammsb_store* store, ammsb_set* train,
                     const uint64_t* d_edges, uint32_t E_mb, float* d_theta_sum /* [K] */,
                     float* d_grads /* [2K] */, void* d_workspace, size_t workspace_bytes);
int ammsb_update_theta(ammsb_ctx* ctx, const ammsb_params* p, float* d_theta, float* d_beta,
                       const float* d_grads, float scale, uint32_t step_count, ammsb_rng* pool);
int ammsb_update_beta(ammsb_ctx* ctx, const ammsb_params* p, float* d_theta, float* d_beta,
                      ammsb_store* store, ammsb_set* train, const uint64_t* d_edges,
                      uint32_t E_mb, float scale, uint32_t step_count, ammsb_rng* pool,
                      float* d_theta_sum, float* d_grads, void* d_workspace, size_t workspace_bytes);

/* ---- PerplexityCalculator::operator()() (perplexity.h:34, perplexity.cc:251-274).
 *      ammsb_perplexity_partial leaves {link_lik, non_link_lik, link_count,
 *      non_link_count} as 4 doubles in d_sums (all-reduced across GPUs by the
 *      caller); ammsb_perplexity also copies them back and returns
 *      -(sum lik)/(sum count) (Learner::HeldoutPerplexity exponentiates). ---- */
int ammsb_perplexity_workspace_bytes(ammsb_ctx* ctx, size_t* bytes);
int ammsb_perplexity_partial(ammsb_ctx* ctx, const ammsb_params* p, ammsb_store* store,
                             const float* d_beta, ammsb_set* heldout, const uint64_t* d_edges,
                             uint32_t H, float* d_ppx_per_edge, uint32_t call_count,
                             double* d_sums /* [4] */, void* d_workspace, size_t workspace_bytes);
int ammsb_perplexity(ammsb_ctx* ctx, const ammsb_params* p, ammsb_store* store,
                     const float* d_beta, ammsb_set* heldout, const uint64_t* d_edges,
                     uint32_t H, float* d_ppx_per_edge, uint32_t call_count,
                     double* h_sums /* [4], may be NULL */, double* h_avg,
                     void* d_workspace, size_t workspace_bytes);

/* ---- multi-GPU exchange steps as kernels over NVLink peer memory (one process per GPU; the
 *      reference has no multi-device code).  A mailbox per rank, shared with the peers as a
 *      file descriptor like the pi shards.  ammsb_peer_barrier: every rank's earlier work on its
 *      stream is complete and its peer stores are visible.  ammsb_peer_allreduce_*: sum over
 *      ranks in RANK ORDER (bit-identical on every rank; count * sizeof(T) <= slot_bytes).
 *      All ranks must issue the same sequence of these calls. ---- */
typedef struct ammsb_peer ammsb_peer;
int ammsb_peer_create(ammsb_ctx* ctx, uint32_t world, uint32_t rank, size_t slot_bytes, ammsb_peer** out);
int ammsb_peer_destroy(ammsb_peer* peer);
int ammsb_peer_export_fd(ammsb_peer* peer, int* fd);
int ammsb_peer_attach_fd(ammsb_peer* peer, uint32_t peer_rank, int fd);
int ammsb_peer_barrier(ammsb_ctx* ctx, ammsb_peer* peer);
int ammsb_peer_allreduce_f32(ammsb_ctx* ctx, ammsb_peer* peer, float* d_inout, uint32_t count);
int ammsb_peer_allreduce_f64(ammsb_ctx* ctx, ammsb_peer* peer, double* d_inout, uint32_t count);
/* a wait inside one of the kernels above gives up after a few seconds (a peer died or never
 * launched) instead of hanging the GPU; *timed_out_epoch is the call number of the first such
 * wait, 0 if there was none.  Synchronous. */
int ammsb_peer_check(ammsb_peer* peer, uint32_t* timed_out_epoch);

/* ---- column-sharded pi over the GPUs of one box (csrc/cols.cu): the multi-GPU successor of
 *      RowPartitionedMatrix (partitioned-alloc.h:14-141) in which partial SUMS cross NVLink, not pi
 *      rows.  GPU g of `world` (2, 4 or 8) holds, for EVERY row, the columns of the reference
 *      work-items l = g (mod world) of the default update_phi launch (work-item l owns
 *      k = l, l+32, ...; phi.cc:214-302) as a local matrix [N][K/world]; phi[N] (row sums) and the
 *      held-out running means are held identically by every rank; theta/beta [2K] live in a
 *      mailbox that the peers map, every rank stepping the columns it owns and publishing them.
 *      The K-wide sums of update_phi / update_pi / update_beta / perplexity are formed as
 *      per-GPU partials and completed in the reference's WG_SUM tree order (sum.cc:20-42) from
 *      self-validating 4-byte mailbox words written by peer stores inside the kernels: results
 *      are bit-identical on every rank, independent of `world`, and update_phi/update_pi are
 *      bit-identical to ammsb_update_phi/ammsb_update_pi on one GPU.
 *      One process per GPU exports its mailbox as a file descriptor (export_fd / attach_fd);
 *      several ranks of one process attach with attach_local.  The step functions take the
 *      `nv` ranks that live on ctx's device -- 1 in production; world ranks on one GPU is the
 *      emulation the single-GPU tests use (one cooperative launch over all ranks' data).
 *      K in {128, 256, 512, 1024}; reproduces the WG launch with phi_wg_size 32. ---- */
typedef struct ammsb_cols ammsb_cols;
int ammsb_cols_create(ammsb_ctx* ctx, uint64_t N, uint32_t K, uint32_t world, uint32_t rank,
                      uint32_t num_neighbors, uint32_t max_nodes, uint32_t max_edges, uint64_t max_pairs,
                      ammsb_cols** out);
int ammsb_cols_destroy(ammsb_cols* cols);
int ammsb_cols_mailbox_bytes(const ammsb_cols* cols, size_t* bytes);
int ammsb_cols_export_fd(ammsb_cols* cols, int* fd);
int ammsb_cols_attach_fd(ammsb_cols* cols, uint32_t peer_rank, int fd);
int ammsb_cols_attach_local(ammsb_cols* cols, uint32_t peer_rank, ammsb_cols* peer);
/* timing diagnostics only (with AMMSB_COLS_LOOPBACK=1 in the environment the kernels neither
 * send to nor wait for a peer): unattached peer mailboxes alias the own one, so that one rank's
 * share of a step can be timed on one GPU; the results of such a run are meaningless */
int ammsb_cols_alias_self(ammsb_cols* cols);
int ammsb_cols_init_pi(ammsb_cols* cols, float eta0, float eta1); /* random.cc:131-167 */
/* host access in the reference's layout: full rows [nrows][K]; write scatters the columns this
 * rank owns, read fills them in and leaves the other columns of h_rows untouched */
int ammsb_cols_write_pi(ammsb_cols* cols, uint64_t row0, uint64_t nrows, const float* h_rows);
int ammsb_cols_read_pi(ammsb_cols* cols, uint64_t row0, uint64_t nrows, float* h_rows);
int ammsb_cols_write_phi(ammsb_cols* cols, uint64_t row0, uint64_t nrows, const float* h_src);
int ammsb_cols_read_phi(ammsb_cols* cols, uint64_t row0, uint64_t nrows, float* h_dst);
int ammsb_cols_write_theta(ammsb_cols* cols, const float* h_theta /* [2K] */, const float* h_beta /* [2K] */);
int ammsb_cols_read_theta(ammsb_cols* cols, float* h_theta, float* h_beta); /* either may be NULL */
int ammsb_cols_beta_ptr(ammsb_cols* cols, float** d_theta, float** d_beta);
int ammsb_cols_read_phi_vec(ammsb_cols* cols, uint32_t V, float* h_rows /* [V][K] */);
int ammsb_cols_check(ammsb_cols* cols, uint32_t* timed_out); /* 1: an exchange wait gave up */
/* PhiUpdater (phi.cc:728-763), BetaUpdater (beta.cc:334-384), PerplexityCalculator
 * (perplexity.cc:251-274) on the column shards; pools[i] is rank i's RNG pool of the operator
 * (same sizes and seeds as on one GPU: a rank only touches the states of its own lanes/columns).
 * step_count / call_count also select the mailbox half and must agree on every rank. */
/* NeighborSampler::operator() (sample.h:16-28) partitioned over the ranks: the reference work-item
 * gid (state gid of the sampler pool) runs on rank gid % world and delivers its slots' lists to
 * every rank's mailbox; ammsb_cols_update_phi with d_neighbors == NULL reads them there.  Same
 * lists as ammsb_neighbor_sample. */
int ammsb_cols_neighbor_sample(ammsb_ctx* ctx, ammsb_cols* const* ranks, uint32_t nv, const uint32_t* d_nodes,
                               uint32_t V, uint32_t wg, uint32_t step_count, ammsb_rng* const* pools);
int ammsb_cols_update_phi(ammsb_ctx* ctx, ammsb_cols* const* ranks, uint32_t nv, const ammsb_params* p,
                          const ammsb_phi_opts* opts, ammsb_set* train, const uint32_t* d_nodes,
                          const uint32_t* d_neighbors, uint32_t V, uint32_t step_count, ammsb_rng* const* pools);
int ammsb_cols_update_pi(ammsb_ctx* ctx, ammsb_cols* const* ranks, uint32_t nv, const uint32_t* d_nodes,
                         uint32_t V, uint32_t step_count);
int ammsb_cols_update_beta(ammsb_ctx* ctx, ammsb_cols* const* ranks, uint32_t nv, const ammsb_params* p,
                           ammsb_set* train, const uint64_t* d_edges, uint32_t E_mb, float scale,
                           uint32_t step_count, ammsb_rng* const* pools);
int ammsb_cols_perplexity(ammsb_ctx* ctx, ammsb_cols* const* ranks, uint32_t nv, const ammsb_params* p,
                          ammsb_set* heldout, const uint64_t* d_edges, uint32_t H, uint32_t call_count,
                          double* h_sums /* [nv][4] or NULL */, double* h_avg /* [nv] or NULL */);
/* results of a rank's last ammsb_cols_perplexity launch (waits for the stream): for callers that
 * drive several devices from one thread and launch every device's kernel (NULL outputs) before
 * they wait for any of them -- the kernels of the ranks wait for each other */
int ammsb_cols_perplexity_result(ammsb_ctx* ctx, ammsb_cols* cols, double* h_sums /* [4] or NULL */,
                                 double* h_avg /* or NULL */);

/* ---- work-group helpers the reference tests directly (wg-sum-test.cc,
 *      wg-normalize-test.cc): rows of `len` floats, one warp per row, reference
 *      association for wg = 32 (sum.cc:31-42, normalize.cc:13-32). ---- */
int ammsb_row_sum(ammsb_ctx* ctx, const float* d_in, uint32_t rows, uint32_t len, float* d_out);
int ammsb_row_normalize(ammsb_ctx* ctx, float* d_inout, uint32_t rows, uint32_t len, float* d_sum_out);

/* number of kernels this library has launched on the calling process so far */
int ammsb_launch_count(uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* AMMSB_H_ */
