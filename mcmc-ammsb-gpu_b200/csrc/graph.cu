// graph.cu -- the inputs of the hot path built on the device: cuckoo edge set, synthetic edge
// list, adjacency, and the Node mini-batch strategy.
//
// The reference builds all of this on the host (cuckoo.cc:117-197, data.cc:12-128,
// sample.cc:253-302) and uploads it.  At the com-Friendster shape (1.8 G edges) a host build
// takes minutes and tens of GB per process; here the same objects are produced in HBM:
//   ammsb_set_build*       cuckoo::Set::SetContents -- same table geometry and hash functions
//                          (so Set_HasEdge, cuckoo.cc:39-65, is unchanged and the table can be
//                          handed to the host Set's consumers); placement is parallel, so the
//                          LAYOUT differs from the host's random walk, the MEMBERSHIP does not.
//   ammsb_graph_generate   the synthetic graphs of BASELINE.json's shapes: E distinct pairs u < v
//   ammsb_graph_nonlinks   the held-out non-links of GenerateSetsFromEdges (data.cc:110-126)
//   ammsb_graph_csr        mcmc::Graph (data.cc:12-25): neighbors of every vertex
//   ammsb_minibatch_*      sampleNodeLink / sampleNodeNonLink (sample.cc:253-293) + the node
//                          extraction of learner.cc:162-173.  The candidate stream is glibc's
//                          rand_r stream of the host strategy (LCG jump-ahead), every candidate
//                          gets the reference's fate (refused if in either set, dropped if seen
//                          before, the batch ends with the m-th pick), so a mini-batch holds the
//                          SAME EDGES AND NODES as the host strategy's for the same seed and
//                          leaves the seed where the host leaves it; they are emitted in draw
//                          order, not in libstdc++'s std::unordered_set order.
#include <cub/cub.cuh>

#include <cmath>
#include <string>

#include "common.cuh"

#define SET_EMPTY 0xffffffffffffffffull

static const uint64_t kPrimes[4][2] = {  // cuckoo.cc:30-35
    {15485807ull, 920429591ull}, {379906717ull, 740320571ull}, {256204747ull, 379927517ull}, {13ull, 17ull}};

// splitmix64 finaliser: a bijection of the 64-bit integers (and its inverse)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ uint64_t unxorshift(uint64_t y, int s) {
  uint64_t x = y;
  x = y ^ (x >> s);
  x = y ^ (x >> s);
  x = y ^ (x >> s);
  return x;
}
__host__ __device__ __forceinline__ uint64_t unmix64(uint64_t x) {
  x = unxorshift(x, 31);
  x *= 0x319642b2d24d8ec3ull;
  x = unxorshift(x, 27);
  x *= 0x96de1b173f119089ull;
  x = unxorshift(x, 30);
  return x;
}

// ------------------------------------------------------------ cuckoo build ----

// One thread per key.  A key goes to a free slot of one of its two bins (claimed with a
// compare-and-swap); when both bins are full it swaps itself with a pseudo-randomly chosen
// occupant and carries on with the displaced key (atomic exchange: no key is ever lost).
__global__ void k_set_build(unsigned long long* table, uint64_t num_bins, uint64_t p1, uint64_t p2,
                            const uint64_t* __restrict__ keys, uint64_t n, uint32_t max_moves, uint32_t* failed) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned long long k = keys[i];
    if (k == SET_EMPTY) continue;  // KEY_INVALID cannot be stored
    uint64_t salt = i;
    for (uint32_t moves = 0;; ++moves) {
      bool placed = false;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (placed) break;
        const uint64_t bin = b == 0 ? (p1 * k) % num_bins : (k ^ p2) % num_bins;
        unsigned long long* cell = table + ((b ? num_bins : 0) + bin) * 4;
        for (int s = 0; s < 4; ++s) {
          const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&cell[s]);
          if (cur == k) {
            placed = true;
            break;
          }
          if (cur == SET_EMPTY) {
            const unsigned long long old = atomicCAS(&cell[s], SET_EMPTY, k);
            if (old == SET_EMPTY || old == k) {
              placed = true;
              break;
            }
          }
        }
      }
      if (placed) break;
      if (moves >= max_moves) {
        atomicExch(failed, 1u);
        break;
      }
      salt += 0x9E3779B97F4A7C15ull;
      const uint64_t r = mix64(k ^ salt);
      const int b = (int)(r & 1), s = (int)((r >> 1) & 3);
      const uint64_t bin = b == 0 ? (p1 * k) % num_bins : (k ^ p2) % num_bins;
      k = atomicExch(&table[((b ? num_bins : 0) + bin) * 4 + s], k);
      if (k == SET_EMPTY) break;
    }
  }
}

static uint64_t set_bins_for(uint64_t n) {  // cuckoo.cc:98-104
  return static_cast<uint64_t>(1 + std::ceil((1.15 * n) / 8));
}

extern "C" int ammsb_set_build_device(ammsb_ctx* c, const uint64_t* d_keys, uint64_t n, ammsb_set** out) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_set* s = new ammsb_set();
  s->ctx = c;
  s->num_bins = set_bins_for(n);
  const size_t bytes = sizeof(uint64_t) * 8 * s->num_bins;
  uint32_t* d_failed = nullptr;
  if (cudaMalloc((void**)&s->d_table, bytes) != cudaSuccess || cudaMalloc((void**)&d_failed, 4) != cudaSuccess) {
    cudaFree(s->d_table);
    delete s;
    AMMSB_REQUIRE(false, "out of device memory for the cuckoo table");
  }
  int rc = 1;
  // like the host build, fall back to the next pair of hash constants if placement fails
  for (uint32_t idx = 0; idx < 4 && rc; ++idx) {
    s->prime_idx = idx;
    cudaMemsetAsync(s->d_table, 0xff, bytes, c->stream);
    cudaMemsetAsync(d_failed, 0, 4, c->stream);
    if (n > 0) {
      uint64_t blocks = (n + 255) / 256;
      if (blocks > (uint64_t)c->sm_count * 16) blocks = (uint64_t)c->sm_count * 16;
      k_set_build<<<(unsigned)blocks, 256, 0, c->stream>>>((unsigned long long*)s->d_table, s->num_bins,
                                                           kPrimes[idx][0], kPrimes[idx][1], d_keys, n, 4096, d_failed);
      g_launch_count.fetch_add(1);
    }
    uint32_t failed = 1;
    if (cudaMemcpyAsync(&failed, d_failed, 4, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      break;
    if (!failed) rc = 0;
  }
  cudaFree(d_failed);
  if (rc) {
    cudaFree(s->d_table);
    delete s;
    AMMSB_REQUIRE(false, "Failed to insert into the cuckoo set");
  }
  *out = s;
  return 0;
}

extern "C" int ammsb_set_build(ammsb_ctx* c, const uint64_t* h_keys, uint64_t n, ammsb_set** out) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint64_t* d_keys = nullptr;
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_keys, 8 * (n ? n : 1)));
  int rc = ammsb_h2d(c, d_keys, h_keys, 8 * n);
  if (!rc) rc = ammsb_set_build_device(c, d_keys, n, out);
  cudaFree(d_keys);
  return rc;
}

extern "C" int ammsb_set_info(const ammsb_set* s, uint64_t* num_bins, uint32_t* prime_idx) {
  if (num_bins) *num_bins = s->num_bins;
  if (prime_idx) *prime_idx = s->prime_idx;
  return 0;
}

extern "C" int ammsb_set_read_table(ammsb_set* s, uint64_t* h_table) {
  return ammsb_d2h(s->ctx, h_table, s->d_table, sizeof(uint64_t) * 8 * s->num_bins);
}

// ----------------------------------------------------- synthetic edge lists ----

__device__ __forceinline__ uint64_t random_pair(uint64_t N, uint64_t seed, uint64_t i) {
  const uint64_t r1 = mix64(seed + 2 * i), r2 = mix64(seed + 2 * i + 1);
  const uint32_t u = (uint32_t)(r1 % N);
  uint32_t v = (uint32_t)(r2 % (N - 1));  // uniform over the vertices other than u
  if (v >= u) ++v;
  return make_edge(min(u, v), max(u, v));
}

// candidate i in scrambled form: mix64 is a bijection, so equal scrambled values are equal
// edges, and sorting by the scrambled value both groups duplicates and shuffles the list
__global__ void k_gen_candidates(uint64_t N, uint64_t seed, uint64_t salt, uint64_t count, uint64_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) out[i] = mix64(random_pair(N, seed, i) ^ salt);
}
__global__ void k_gen_nonlink_candidates(uint64_t N, uint64_t seed, uint64_t salt, uint64_t count, SetView a,
                                         SetView b, int has_b, uint64_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    const uint64_t e = random_pair(N, seed, i);
    const bool taken = set_has(a, e) || (has_b && set_has(b, e));
    out[i] = taken ? ~0ull : mix64(e ^ salt);  // refused candidates sort to the very end
  }
}
__global__ void k_unscramble(uint64_t* keys, uint64_t count, uint64_t salt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) keys[i] = unmix64(keys[i]) ^ salt;
}

static unsigned grid_for(const ammsb_ctx* c, uint64_t n) {
  uint64_t blocks = (n + 255) / 256;
  const uint64_t cap = (uint64_t)c->sm_count * 16;
  return (unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

// `want` distinct pairs in pseudo-random order into d_out; nonlink: skip members of a (and b)
static int generate_unique(ammsb_ctx* c, uint64_t N, uint64_t want, uint64_t seed, const ammsb_set* a,
                           const ammsb_set* b, uint64_t* d_out) {
  AMMSB_REQUIRE(N >= 2 && N < 0xffffffffull, "N out of range");
  AMMSB_REQUIRE(want < (1ull << 31) - (1ull << 26), "edge count too large for one generation pass");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  if (want == 0) return 0;
  const uint64_t salt = mix64(seed ^ 0x5851f42d4c957f2dull);
  uint64_t count = want + want / 32 + 4096;
  for (int attempt = 0; attempt < 6; ++attempt, count += count / 2) {
    AMMSB_REQUIRE(count < (1ull << 31), "too many candidates for one generation pass (graph too dense?)");
    uint64_t *d_a = nullptr, *d_b = nullptr, *d_num = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_sort = 0, tmp_uniq = 0;
    cub::DoubleBuffer<uint64_t> buf(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, buf, (int)count, 0, 64, c->stream);
    cub::DeviceSelect::Unique(nullptr, tmp_uniq, d_a, d_b, d_num, (int)count, c->stream);
    const size_t tmp_bytes = tmp_sort > tmp_uniq ? tmp_sort : tmp_uniq;
    cudaError_t e1 = cudaMalloc((void**)&d_a, 8 * count), e2 = cudaMalloc((void**)&d_b, 8 * count),
                e3 = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 8), e4 = cudaMalloc((void**)&d_num, 8);
    auto release = [&]() {
      cudaFree(d_a);
      cudaFree(d_b);
      cudaFree(d_tmp);
      cudaFree(d_num);
    };
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
      release();
      cudaGetLastError();
      AMMSB_REQUIRE(false, "out of device memory while generating edges");
    }
    // every step is checked on its own: at the Friendster size a failure must say where
#define GEN_STEP(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      release();                                                                              \
      ammsb_set_error(std::string("generate_unique: ") + #expr + ": " + cudaGetErrorString(_e) + \
                      " (count " + std::to_string(count) + ")");                              \
      cudaGetLastError();                                                                     \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)
    if (a != nullptr)
      k_gen_nonlink_candidates<<<grid_for(c, count), 256, 0, c->stream>>>(N, seed, salt, count, a->view(),
                                                                          b ? b->view() : a->view(), b != nullptr, d_a);
    else
      k_gen_candidates<<<grid_for(c, count), 256, 0, c->stream>>>(N, seed, salt, count, d_a);
    g_launch_count.fetch_add(1);
    GEN_STEP(cudaGetLastError());
    buf = cub::DoubleBuffer<uint64_t>(d_a, d_b);
    size_t tb = tmp_bytes;
    // count < 2^31 (checked above): the 32-bit item count is the well-trodden path of both primitives
    GEN_STEP(cub::DeviceRadixSort::SortKeys(d_tmp, tb, buf, (int)count, 0, 64, c->stream));
    uint64_t* sorted = buf.Current();
    uint64_t* other = buf.Alternate();
    tb = tmp_bytes;
    GEN_STEP(cub::DeviceSelect::Unique(d_tmp, tb, sorted, other, d_num, (int)count, c->stream));
    uint64_t num = 0;
    GEN_STEP(cudaMemcpyAsync(&num, d_num, 8, cudaMemcpyDeviceToHost, c->stream));
    GEN_STEP(cudaStreamSynchronize(c->stream));
    // refused candidates were all mapped to ~0: at most one survivor, and it is the last
    if (a != nullptr && num > 0) {
      uint64_t last = 0;
      GEN_STEP(cudaMemcpy(&last, other + num - 1, 8, cudaMemcpyDeviceToHost));
      if (last == ~0ull) --num;
    }
    if (num >= want) {
      k_unscramble<<<grid_for(c, want), 256, 0, c->stream>>>(other, want, salt);
      g_launch_count.fetch_add(1);
      GEN_STEP(cudaGetLastError());
      GEN_STEP(cudaMemcpyAsync(d_out, other, 8 * want, cudaMemcpyDeviceToDevice, c->stream));
      GEN_STEP(cudaStreamSynchronize(c->stream));
      release();
      return 0;
    }
    release();
#undef GEN_STEP
  }
  AMMSB_REQUIRE(false, "could not draw enough distinct pairs (graph too dense?)");
}

extern "C" int ammsb_graph_generate(ammsb_ctx* c, uint64_t N, uint64_t E, uint64_t seed, uint64_t* d_edges) {
  return generate_unique(c, N, E, seed, nullptr, nullptr, d_edges);
}

extern "C" int ammsb_graph_nonlinks(ammsb_ctx* c, uint64_t N, uint64_t count, uint64_t seed, ammsb_set* a,
                                    ammsb_set* b, uint64_t* d_out) {
  AMMSB_REQUIRE(a != nullptr, "a set to avoid is required");
  return generate_unique(c, N, count, seed, a, b, d_out);
}

// ---------------------------------------------------------------- adjacency ----

__global__ void k_degree(const uint64_t* __restrict__ edges, uint64_t E, unsigned long long* deg) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) {
    const uint64_t e = edges[i];
    atomicAdd(&deg[(uint32_t)(e >> 32)], 1ull);
    atomicAdd(&deg[(uint32_t)e], 1ull);
  }
}
__global__ void k_fill_adj(const uint64_t* __restrict__ edges, uint64_t E, const uint64_t* __restrict__ offsets,
                           unsigned long long* cursor, uint32_t* adj) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < E; i += stride) {
    const uint64_t e = edges[i];
    const uint32_t u = (uint32_t)(e >> 32), v = (uint32_t)e;
    adj[offsets[u] + atomicAdd(&cursor[u], 1ull)] = v;
    adj[offsets[v] + atomicAdd(&cursor[v], 1ull)] = u;
  }
}
// the fill order depends on the scheduling of the atomics; sorting every list makes the result a
// function of the edge set alone (lists are short: insertion sort, one thread per vertex)
__global__ void k_sort_adj(const uint64_t* __restrict__ offsets, uint64_t N, uint32_t* adj, uint32_t* degree32) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; v < N; v += stride) {
    uint32_t* a = adj + offsets[v];
    const uint64_t d = offsets[v + 1] - offsets[v];
    uint32_t buf[128];  // short lists are sorted in local memory (L1) and written back once
    uint32_t* w = a;
    if (d <= 128) {
      for (uint64_t i = 0; i < d; ++i) buf[i] = a[i];
      w = buf;
    }
    for (uint64_t i = 1; i < d; ++i) {
      const uint32_t x = w[i];
      uint64_t j = i;
      for (; j > 0 && w[j - 1] > x; --j) w[j] = w[j - 1];
      w[j] = x;
    }
    if (d <= 128)
      for (uint64_t i = 0; i < d; ++i) a[i] = buf[i];
    if (degree32) degree32[v] = (uint32_t)d;
  }
}

// d_offsets [N+1] u64, d_adj [2E] u32 (neighbors of v: d_adj[d_offsets[v] .. d_offsets[v+1]), ascending),
// d_degree [N] u32 (may be NULL)
extern "C" int ammsb_graph_csr(ammsb_ctx* c, uint64_t N, const uint64_t* d_edges, uint64_t E,
                               uint64_t* d_offsets, uint32_t* d_adj, uint32_t* d_degree) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  unsigned long long* d_deg = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  AMMSB_REQUIRE(N < 0x7fffffffull, "N out of range");
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)(N + 1), c->stream);
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_deg, 8 * (N + 1)));
  if (cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 8) != cudaSuccess) {
    cudaFree(d_deg);
    AMMSB_REQUIRE(false, "out of device memory while building the adjacency");
  }
  cudaMemsetAsync(d_deg, 0, 8 * (N + 1), c->stream);
  if (E > 0) k_degree<<<grid_for(c, E), 256, 0, c->stream>>>(d_edges, E, d_deg);
  cudaError_t es = cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, (uint64_t*)d_deg, d_offsets, (int)(N + 1), c->stream);
  cudaMemsetAsync(d_deg, 0, 8 * (N + 1), c->stream);
  if (E > 0) k_fill_adj<<<grid_for(c, E), 256, 0, c->stream>>>(d_edges, E, d_offsets, d_deg, d_adj);
  k_sort_adj<<<grid_for(c, N), 256, 0, c->stream>>>(d_offsets, N, d_adj, d_degree);
  g_launch_count.fetch_add(3);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(d_deg);
  cudaFree(d_tmp);
  AMMSB_CHECK_CUDA(es);
  AMMSB_CHECK_CUDA(e);
  AMMSB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------- Node mini-batches ----

// glibc rand_r: three steps of x -> 1103515245 x + 12345 (mod 2^32) per draw
__host__ __device__ __forceinline__ uint32_t lcg_jump(uint32_t x, uint64_t steps) {
  uint32_t am = 1103515245u, cm = 12345u;  // x -> am x + cm, to the power 2^bit
  for (; steps; steps >>= 1) {
    if (steps & 1) x = am * x + cm;
    cm = (am + 1u) * cm;
    am = am * am;
  }
  return x;
}
__device__ __forceinline__ uint32_t rand_r_dev(uint32_t* seed) {
  uint32_t next = *seed;
  next = next * 1103515245u + 12345u;
  uint32_t result = (next >> 16) & 2047u;
  next = next * 1103515245u + 12345u;
  result = (result << 10) ^ ((next >> 16) & 1023u);
  next = next * 1103515245u + 12345u;
  result = (result << 10) ^ ((next >> 16) & 1023u);
  *seed = next;
  return result;
}

struct NonLinkArgs {
  SetView train, heldout;
  int has_heldout;
  uint32_t N, u, seed;   // seed: rand_r state after u was drawn
  uint32_t count;        // candidates examined by this pass
  uint32_t m;
  uint32_t cap_mask;     // de-duplication table: cap_mask + 1 slots
  uint32_t* tab_v;       // [cap] vertex, 0xffffffff = free
  uint32_t* tab_i;       // [cap] first candidate index with that vertex
  uint32_t* cand_v;      // [count]
  uint32_t* cand_slot;   // [count] table slot of the candidate, 0xffffffff = refused
  uint64_t* flags;       // [count] picked | (picked and v != u) << 32, then their exclusive sums
  uint64_t* edges;       // out [m]
  uint32_t* nodes;       // out [m + 1]
  uint32_t* header;      // out: {draws consumed, picked, nodes, 0}
};

#define NL_PER_THREAD 8
// candidate i = Canonical(u, rand_r_i % N) (sample.cc:284-286), its fate in the two sets, and its
// entry in the de-duplication table (the smallest index wins: "first seen")
__global__ void k_nonlink_draw(const NonLinkArgs a) {
  const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) * NL_PER_THREAD;
  if (first >= a.count) return;
  uint32_t s = lcg_jump(a.seed, 3ull * first);
  for (uint32_t i = first; i < first + NL_PER_THREAD && i < a.count; ++i) {
    const uint32_t v = rand_r_dev(&s) % a.N;
    const uint64_t e = make_edge(min(a.u, v), max(a.u, v));
    a.cand_v[i] = v;
    const bool refused = (a.has_heldout && set_has(a.heldout, e)) || set_has(a.train, e);
    uint32_t slot = 0xffffffffu;
    if (!refused) {
      uint32_t h = (v * 2654435761u) & a.cap_mask;
      for (;;) {
        const uint32_t old = atomicCAS(&a.tab_v[h], 0xffffffffu, v);
        if (old == 0xffffffffu || old == v) break;
        h = (h + 1) & a.cap_mask;
      }
      atomicMin(&a.tab_i[h], i);
      slot = h;
    }
    a.cand_slot[i] = slot;
  }
}
__global__ void k_nonlink_flags(const NonLinkArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const uint32_t slot = a.cand_slot[i];
  const bool picked = slot != 0xffffffffu && a.tab_i[slot] == i;
  a.flags[i] = picked ? (1ull | ((uint64_t)(a.cand_v[i] != a.u) << 32)) : 0ull;
}
// flags now hold exclusive sums {picks before i, picks with v != u before i}
__global__ void k_nonlink_emit(const NonLinkArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const uint32_t slot = a.cand_slot[i];
  const bool picked = slot != 0xffffffffu && a.tab_i[slot] == i;
  const uint32_t pos = (uint32_t)a.flags[i], npos = (uint32_t)(a.flags[i] >> 32);
  if (i == 0) a.nodes[0] = a.u;
  if (picked && pos < a.m) {
    const uint32_t v = a.cand_v[i];
    a.edges[pos] = make_edge(min(a.u, v), max(a.u, v));
    if (v != a.u) a.nodes[1 + npos] = v;
    if (pos == a.m - 1) {  // the pick that completes the mini-batch: the strategy stops drawing here
      a.header[0] = i + 1;
      a.header[1] = a.m;
      a.header[2] = 1 + npos + (v != a.u ? 1 : 0);
    }
  }
  if (i == a.count - 1 && pos + (picked ? 1 : 0) < a.m) {  // not enough picks in this pass
    a.header[0] = 0;
    a.header[1] = pos + (picked ? 1 : 0);
    a.header[2] = 0;
  }
}

struct ammsb_sampler {
  ammsb_ctx* ctx = nullptr;
  uint64_t N = 0;
  uint32_t m = 0, max_count = 0, cap = 0;
  uint32_t *tab_v = nullptr, *tab_i = nullptr, *cand_v = nullptr, *cand_slot = nullptr, *header = nullptr;
  uint64_t* flags = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  uint32_t* h_header = nullptr;  // pinned
};

extern "C" int ammsb_sampler_destroy(ammsb_sampler* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  cudaFree(s->tab_v);
  cudaFree(s->tab_i);
  cudaFree(s->cand_v);
  cudaFree(s->cand_slot);
  cudaFree(s->header);
  cudaFree(s->flags);
  cudaFree(s->d_tmp);
  cudaFreeHost(s->h_header);
  delete s;
  return 0;
}

extern "C" int ammsb_sampler_create(ammsb_ctx* c, uint64_t N, uint32_t mini_batch_size, ammsb_sampler** out) {
  AMMSB_REQUIRE(N >= 1 && N < 0xffffffffull, "N out of range");
  AMMSB_REQUIRE(mini_batch_size >= 1 && mini_batch_size < (1u << 27), "mini-batch size out of range");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_sampler* s = new ammsb_sampler();
  s->ctx = c;
  s->N = N;
  s->m = mini_batch_size;
  // candidates per pass: the picks plus room for repeats and refusals; a pass that falls short is
  // redone with all of max_count
  s->max_count = 4 * mini_batch_size + 4096;
  s->cap = 1;
  while (s->cap < 2 * s->max_count) s->cap <<= 1;
  cub::DeviceScan::ExclusiveSum(nullptr, s->tmp_bytes, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)s->max_count,
                                c->stream);
  bool ok = cudaMalloc((void**)&s->tab_v, 4 * (size_t)s->cap) == cudaSuccess &&
            cudaMalloc((void**)&s->tab_i, 4 * (size_t)s->cap) == cudaSuccess &&
            cudaMalloc((void**)&s->cand_v, 4 * (size_t)s->max_count) == cudaSuccess &&
            cudaMalloc((void**)&s->cand_slot, 4 * (size_t)s->max_count) == cudaSuccess &&
            cudaMalloc((void**)&s->flags, 8 * (size_t)s->max_count) == cudaSuccess &&
            cudaMalloc((void**)&s->header, 16) == cudaSuccess &&
            cudaMalloc(&s->d_tmp, s->tmp_bytes ? s->tmp_bytes : 8) == cudaSuccess &&
            cudaMallocHost((void**)&s->h_header, 16) == cudaSuccess;
  if (!ok) {
    ammsb_sampler_destroy(s);
    cudaGetLastError();
    AMMSB_REQUIRE(false, "out of memory for the mini-batch sampler");
  }
  *out = s;
  return 0;
}

// sampleNodeNonLink (sample.cc:275-293) for vertex u; *seed is the rand_r state after u was drawn
// and is advanced past the draw that completed the mini-batch.  d_edges [m], d_nodes [m + 1].
// Synchronises the context's stream (the number of draws decides the next seed).
extern "C" int ammsb_minibatch_nonlink(ammsb_sampler* s, ammsb_ctx* c, uint32_t u, unsigned int* seed, ammsb_set* train,
                                       ammsb_set* heldout, uint64_t* d_edges, uint32_t* d_nodes, uint32_t* num_edges,
                                       uint32_t* num_nodes) {
  AMMSB_REQUIRE(u < s->N, "vertex out of range");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  NonLinkArgs a;
  a.train = train->view();
  a.heldout = heldout ? heldout->view() : train->view();
  a.has_heldout = heldout != nullptr;
  a.N = (uint32_t)s->N;
  a.u = u;
  a.seed = *seed;
  a.m = s->m;
  a.cap_mask = s->cap - 1;
  a.tab_v = s->tab_v;
  a.tab_i = s->tab_i;
  a.cand_v = s->cand_v;
  a.cand_slot = s->cand_slot;
  a.flags = s->flags;
  a.edges = d_edges;
  a.nodes = d_nodes;
  a.header = s->header;
  uint32_t count = s->m + s->m / 8 + 1024;
  if (count > s->max_count) count = s->max_count;
  for (int pass = 0; pass < 2; ++pass, count = s->max_count) {
    a.count = count;
    AMMSB_CHECK_CUDA(cudaMemsetAsync(s->tab_v, 0xff, 4 * (size_t)s->cap, c->stream));
    AMMSB_CHECK_CUDA(cudaMemsetAsync(s->tab_i, 0xff, 4 * (size_t)s->cap, c->stream));
    const uint32_t draw_threads = (count + NL_PER_THREAD - 1) / NL_PER_THREAD;
    k_nonlink_draw<<<(draw_threads + 127) / 128, 128, 0, c->stream>>>(a);
    k_nonlink_flags<<<(count + 255) / 256, 256, 0, c->stream>>>(a);
    size_t tb = s->tmp_bytes;
    cub::DeviceScan::ExclusiveSum(s->d_tmp, tb, a.flags, a.flags, (int)count, c->stream);
    k_nonlink_emit<<<(count + 255) / 256, 256, 0, c->stream>>>(a);
    g_launch_count.fetch_add(4);
    AMMSB_CHECK_CUDA(cudaMemcpyAsync(s->h_header, s->header, 16, cudaMemcpyDeviceToHost, c->stream));
    AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    if (s->h_header[0] != 0) {
      *seed = lcg_jump(*seed, 3ull * s->h_header[0]);
      *num_edges = s->h_header[1];
      *num_nodes = s->h_header[2];
      return 0;
    }
    AMMSB_REQUIRE(count < s->max_count, "mini-batch not filled: too few vertices left to pair with");
  }
  return 1;
}

__global__ void k_link_emit(uint32_t u, const uint64_t* __restrict__ offsets, const uint32_t* __restrict__ adj,
                            uint64_t* edges, uint32_t* nodes) {
  const uint64_t lo = offsets[u], d = offsets[u + 1] - lo;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) nodes[0] = u;
  if (i < d) {
    const uint32_t v = adj[lo + i];
    edges[i] = make_edge(min(u, v), max(u, v));
    nodes[1 + i] = v;
  }
}

// sampleNodeLink (sample.cc:253-269) for a vertex u with `degree` > 0 training neighbors: all of
// its training edges.  d_edges [degree], d_nodes [degree + 1].  Asynchronous.
extern "C" int ammsb_minibatch_link(ammsb_ctx* c, uint32_t u, uint32_t degree, const uint64_t* d_offsets,
                                    const uint32_t* d_adj, uint64_t* d_edges, uint32_t* d_nodes) {
  AMMSB_REQUIRE(degree > 0, "vertex without training neighbors");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  k_link_emit<<<(degree + 127) / 128, 128, 0, c->stream>>>(u, d_offsets, d_adj, d_edges, d_nodes);
  AMMSB_LAUNCH_CHECK();
  return 0;
}
