// small_kernels.cu -- RNG pools, cuckoo membership, neighbor sampler, pi init and
// the work-group sum/normalise helpers.
#include <stdlib.h>

#include "common.cuh"

// ------------------------------------------------------------------- RNG ----

// RandomInit, random.cc:31-44 -- the reference runs it as ONE work-item looping
// over all states; here one thread per state.
__global__ void k_rng_init(ulonglong2* st, uint64_t n, uint64_t sx, uint64_t sy) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) st[i] = make_ulonglong2(sx + i, sy + i);
}

extern "C" int ammsb_rng_create(ammsb_ctx* c, uint64_t n, uint64_t sx, uint64_t sy,
                                ammsb_rng** out) {
  AMMSB_REQUIRE(n > 0, "empty RNG pool");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ulonglong2* d_state = nullptr;  // before the handle exists: a failed allocation leaks nothing
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_state, sizeof(ulonglong2) * n));
  ammsb_rng* r = new ammsb_rng();
  r->ctx = c;
  r->n = n;
  r->d_state = d_state;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)c->sm_count * 16) blocks = (uint64_t)c->sm_count * 16;
  k_rng_init<<<(unsigned)blocks, 256, 0, c->stream>>>(r->d_state, n, sx, sy);
  AMMSB_LAUNCH_CHECK();
  *out = r;
  return 0;
}

extern "C" int ammsb_rng_destroy(ammsb_rng* r) {
  if (!r) return 0;
  cudaSetDevice(r->ctx->device);
  cudaFree(r->d_state);
  delete r;
  return 0;
}
extern "C" int ammsb_rng_size(const ammsb_rng* r, uint64_t* n) {
  *n = r->n;
  return 0;
}
extern "C" int ammsb_rng_get_state(ammsb_rng* r, uint64_t* h_xy) {
  return ammsb_d2h(r->ctx, h_xy, r->d_state, sizeof(ulonglong2) * r->n);
}
extern "C" int ammsb_rng_set_state(ammsb_rng* r, const uint64_t* h_xy) {
  return ammsb_h2d(r->ctx, r->d_state, h_xy, sizeof(ulonglong2) * r->n);
}

template <int KIND>
__global__ void k_rng_draw(ulonglong2* st, uint64_t n, uint32_t draws, float a, float b,
                           void* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng s = rng_load(st, i);
  for (uint32_t d = 0; d < draws; ++d) {
    if (KIND == 0) ((uint64_t*)out)[i * draws + d] = rng_next(s);
    if (KIND == 1) ((float*)out)[i * draws + d] = rng_randn(s);
    if (KIND == 2) ((float*)out)[i * draws + d] = rng_gamma(s, a, b);
  }
  rng_store(st, i, s);
}

template <int KIND>
static int rng_draw(ammsb_rng* r, uint32_t draws, float a, float b, void* h_out, size_t elt) {
  ammsb_ctx* c = r->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  void* d_out = nullptr;
  size_t bytes = elt * r->n * draws;
  AMMSB_CHECK_CUDA(cudaMalloc(&d_out, bytes));
  k_rng_draw<KIND><<<(unsigned)((r->n + 127) / 128), 128, 0, c->stream>>>(r->d_state, r->n, draws,
                                                                         a, b, d_out);
  AMMSB_LAUNCH_CHECK();
  int rc = ammsb_d2h(c, h_out, d_out, bytes);
  cudaFree(d_out);
  return rc;
}
extern "C" int ammsb_rng_draw_u64(ammsb_rng* r, uint32_t draws, uint64_t* h) {
  return rng_draw<0>(r, draws, 0, 0, h, 8);
}
extern "C" int ammsb_rng_draw_randn(ammsb_rng* r, uint32_t draws, float* h) {
  return rng_draw<1>(r, draws, 0, 0, h, 4);
}
extern "C" int ammsb_rng_draw_gamma(ammsb_rng* r, uint32_t draws, float a, float b, float* h) {
  return rng_draw<2>(r, draws, a, b, h, 4);
}

// ---------------------------------------------------------------- cuckoo ----

__global__ void k_set_has(SetView s, const uint64_t* keys, uint64_t n, uint8_t* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = set_has(s, keys[i]) ? 1 : 0;
}

extern "C" int ammsb_set_has_device(ammsb_set* s, const uint64_t* d_keys, uint64_t n,
                                    uint8_t* d_out) {
  if (n == 0) return 0;
  ammsb_ctx* c = s->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint64_t blocks = (n + 255) / 256;
  if (blocks > (uint64_t)c->sm_count * 8) blocks = (uint64_t)c->sm_count * 8;
  k_set_has<<<(unsigned)blocks, 256, 0, c->stream>>>(s->view(), d_keys, n, d_out);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_set_has(ammsb_set* s, const uint64_t* h_keys, uint64_t n, uint8_t* h_out) {
  if (n == 0) return 0;
  ammsb_ctx* c = s->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint64_t* d_keys = nullptr;
  uint8_t* d_out = nullptr;
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_keys, 8 * n));
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_out, n));
  int rc = ammsb_h2d(c, d_keys, h_keys, 8 * n);
  if (!rc) rc = ammsb_set_has_device(s, d_keys, n, d_out);
  if (!rc) rc = ammsb_d2h(c, h_out, d_out, n);
  cudaFree(d_keys);
  cudaFree(d_out);
  return rc;
}

// ------------------------------------------------------- neighbor sampler ----

// generate_random_int_kernel, sample.cc:48-77.  One thread per reference work-item
// (state index = reference global id; item g serves slots g, g+gsize, ...), so the
// u64 stream and therefore every sampled id is bit-identical.  The per-slot
// open-addressing table lives in shared memory, interleaved by thread so that
// probing is bank-conflict free; it is only spilled to global memory when the
// caller asks for it (NeighborSampler::GetHash()).
__global__ void k_neighbor_sample(ulonglong2* pool, const uint32_t* __restrict__ nodes, uint32_t V,
                                  uint32_t N, uint32_t n, uint32_t gsize,
                                  uint32_t* __restrict__ packed_all, uint32_t* hash_out) {
  extern __shared__ uint32_t s_tab[];
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t capacity = 2 * n;
  if (gid >= gsize || gid >= V) return;
  Rng seed = rng_load(pool, gid);
  uint32_t* tab;
  uint32_t stride;
  for (uint32_t i = gid; i < V; i += gsize) {
    if (hash_out) {
      tab = hash_out + (size_t)i * capacity;
      stride = 1;
    } else {
      tab = s_tab + threadIdx.x;
      stride = blockDim.x;
    }
    const uint32_t node = nodes[i];
    for (uint32_t j = 0; j < capacity; ++j) tab[j * stride] = N;
    for (uint32_t j = 0; j < n; ++j) {
      // generate_random_int, sample.cc:23-46; max_id = N - 1
      uint32_t r, val;
      do {
        do {
          // randint(seed, 0, max_id) = rand % (max_id + 1)   (random.cl.inc:37-39)
          r = (uint32_t)(rng_next(seed) % (uint64_t)N);
        } while (r == node);
        const uint32_t l1 = (r ^ 553105253u) % capacity;
        const uint32_t l2 = 1u + (capacity << 1);
        for (uint32_t q = 0;; ++q) {
          const uint32_t off = (l1 + q * l2) % capacity;
          val = tab[off * stride];
          if (val == r) break;
          if (val == N) {
            tab[off * stride] = r;
            break;
          }
        }
      } while (val == r);
    }
    uint32_t* packed = packed_all + (size_t)i * n;
    uint32_t count = 0;
    for (uint32_t j = 0; j < capacity && count < n; ++j) {
      const uint32_t v = tab[j * stride];
      if (v != N) packed[count++] = v;
    }
  }
  rng_store(pool, gid, seed);
}

// The production launch: a warp per 32 work-items.  Each lane draws its slot's ids into its own
// column of a shared-memory table (stride 33 words: conflict-free for the per-lane probes AND for
// the row-wise read-out), then the warp packs the 32 tables one after the other -- table order,
// first n entries, as sample.cc:64-74 -- with ballots, so that a slot's list leaves the SM as one
// coalesced 4n-byte store instead of n scattered words.  Same draws, same lists, same pool state as
// k_neighbor_sample; no divide in the loop (NsGeom).
#define NS_WARP_STRIDE 33
__global__ void __launch_bounds__(32)
    k_neighbor_sample_warp(ulonglong2* pool, const uint32_t* __restrict__ nodes, uint32_t V, uint32_t gsize,
                           const NsGeom g, uint32_t* __restrict__ packed_all) {
  extern __shared__ uint32_t s_tab[];  // [capacity][33]
  const uint32_t lane = threadIdx.x, first = blockIdx.x * 32, gid = first + lane;
  const bool owner = gid < gsize && gid < V;
  Rng seed;
  seed.x = seed.y = 0;
  if (owner) seed = rng_load(pool, gid);
  const uint32_t lanes_lt = (1u << lane) - 1;
  for (uint32_t base = first; base < V; base += gsize) {  // warp-uniform: the passes of work-item `first`
    const uint32_t i = base + lane;
    if (owner && i < V) ns_draw_slot(seed, __ldg(&nodes[i]), g, s_tab + lane, NS_WARP_STRIDE);
    __syncwarp();
    for (uint32_t t = 0; t < 32; ++t) {
      if (first + t >= gsize || base + t >= V) break;
      uint32_t* packed = packed_all + (size_t)(base + t) * g.n;
      uint32_t count = 0;
      for (uint32_t c = 0; c < g.capacity && count < g.n; c += 32) {
        const uint32_t j = c + lane;
        const uint32_t v = j < g.capacity ? s_tab[j * NS_WARP_STRIDE + t] : g.N;
        const uint32_t m = __ballot_sync(FULL_MASK, v != g.N);
        const uint32_t pos = count + __popc(m & lanes_lt);
        if (v != g.N && pos < g.n) packed[pos] = v;
        count += __popc(m);
      }
    }
    __syncwarp();
  }
  if (owner) rng_store(pool, gid, seed);
}

extern "C" int ammsb_neighbor_sample(ammsb_ctx* c, ammsb_rng* pool, const uint32_t* d_nodes,
                                     uint32_t V, uint32_t N, uint32_t n, uint32_t wg,
                                     uint32_t* d_neighbors, uint32_t* d_hash_out) {
  AMMSB_REQUIRE(V > 0, "mini-batch nodes size = 0!");  // learner.cc:179
  AMMSB_REQUIRE(wg > 0 && n > 0, "bad sampler geometry");
  AMMSB_REQUIRE((uint64_t)n < (uint64_t)N, "num_node_sample must be < N");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  // sample.cc:116-119: global = min(ceil(V/wg), 65535/wg) * wg
  uint32_t groups = V / wg + (V % wg ? 1 : 0);
  if (groups > 65535u / wg) groups = 65535u / wg;
  const uint32_t gsize = groups * wg;
  AMMSB_REQUIRE(pool->n >= (uint64_t)(gsize < V ? gsize : V), "Num seeds smaller than global threads");
  if (d_hash_out == nullptr && sizeof(uint32_t) * 2 * n * NS_WARP_STRIDE <= 48 * 1024 && !getenv("AMMSB_NS_THREAD")) {
    const uint32_t active = gsize < V ? gsize : V;
    k_neighbor_sample_warp<<<(active + 31) / 32, 32, sizeof(uint32_t) * 2 * n * NS_WARP_STRIDE, c->stream>>>(
        pool->d_state, d_nodes, V, gsize, ns_geom(N, n), d_neighbors);
    AMMSB_LAUNCH_CHECK();
    return 0;
  }
  const uint32_t block = 64;
  size_t smem = d_hash_out ? 0 : sizeof(uint32_t) * 2 * n * block;
  if (smem > c->smem_optin) {
    ammsb_set_error("num_node_sample too large for the shared-memory table; pass d_hash_out");
    return 1;
  }
  if (smem > 48 * 1024)
    AMMSB_CHECK_CUDA(cudaFuncSetAttribute(k_neighbor_sample,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint32_t active = gsize < V ? gsize : V;
  k_neighbor_sample<<<(active + block - 1) / block, block, smem, c->stream>>>(
      pool->d_state, d_nodes, V, N, n, gsize, d_neighbors, d_hash_out);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- pi init ----

// generate_gamma (random.cc:108-127) + WG_NORMALIZE_PARTITIONED_KERNEL
// (normalize.cc:34-52), fused.  Reference geometry: G = min(N,65535) groups of 32;
// group g draws rows g, g+G, ...; lane l draws columns l, l+32, ...; state index =
// g*32 + l of a pool seeded {11,113} (random.cc:159-166).  One warp per reference
// group.  Every shard walks the whole row sequence of its groups (the stream is
// sequential per group) but only stores the rows it owns.
__global__ void k_init_pi(StoreView sv, uint32_t shard_id, uint32_t G, float eta0, float eta1) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= G) return;
  const uint64_t id = (uint64_t)warp * 32 + lane;
  Rng s;
  s.x = 11 + id;
  s.y = 113 + id;
  const uint32_t K = sv.K;
  const uint32_t row_lo = shard_id * sv.rows_per_shard;
  const uint32_t row_hi = min(sv.N, row_lo + sv.rows_per_shard);
  for (uint32_t row = warp; row < row_hi; row += G) {
    const bool mine = row >= row_lo;
    float* r = mine ? store_row(sv, row) : nullptr;
    float lsum = 0.f;
    for (uint32_t j = lane; j < K; j += 32) {
      const float g = rng_gamma(s, eta0, eta1);
      if (mine) {
        r[j] = g;
        lsum = __fadd_rn(lsum, g);
      }
    }
    if (mine) {
      const float sum = warp_sum(lsum);
      for (uint32_t j = lane; j < K; j += 32) r[j] = __fdiv_rn(r[j], sum);
      if (lane == 0) *store_phi(sv, row) = sum;
    }
  }
}

extern "C" int ammsb_store_init_pi(ammsb_store* s, float eta0, float eta1) {
  ammsb_ctx* c = s->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  const uint32_t G = s->N < 65535 ? (uint32_t)s->N : 65535u;
  const uint32_t block = 128;
  const uint32_t blocks = (G * 32 + block - 1) / block;
  k_init_pi<<<blocks, block, 0, c->stream>>>(s->view(), s->shard_id, G, eta0, eta1);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------ row sum / row normalise ----

// WG_SUM_KERNEL / WG_NORMALIZE_KERNEL with WG_SIZE 32 (sum.cc:44-53, normalize.cc:25-32):
// one warp per row, lane-strided partials, shuffle tree.
__global__ void k_row_sum(const float* in, uint32_t rows, uint32_t len, float* out, float* norm) {
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  for (; warp < rows; warp += nwarps) {
    const float* r = in + (size_t)warp * len;
    float lsum = 0.f;
    for (uint32_t i = lane; i < len; i += 32) lsum = __fadd_rn(lsum, r[i]);
    const float sum = warp_sum(lsum);
    if (norm) {
      float* w = norm + (size_t)warp * len;
      for (uint32_t i = lane; i < len; i += 32) w[i] = __fdiv_rn(r[i], sum);
    }
    if (out && lane == 0) out[warp] = sum;
  }
}

extern "C" int ammsb_row_sum(ammsb_ctx* c, const float* d_in, uint32_t rows, uint32_t len,
                             float* d_out) {
  if (rows == 0) return 0;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint32_t blocks = (rows + 3) / 4;
  if (blocks > (uint32_t)c->sm_count * 16) blocks = c->sm_count * 16;
  k_row_sum<<<blocks, 128, 0, c->stream>>>(d_in, rows, len, d_out, nullptr);
  AMMSB_LAUNCH_CHECK();
  return 0;
}
extern "C" int ammsb_row_normalize(ammsb_ctx* c, float* d_inout, uint32_t rows, uint32_t len,
                                   float* d_sum_out) {
  if (rows == 0) return 0;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint32_t blocks = (rows + 3) / 4;
  if (blocks > (uint32_t)c->sm_count * 16) blocks = c->sm_count * 16;
  k_row_sum<<<blocks, 128, 0, c->stream>>>(d_inout, rows, len, d_sum_out, d_inout);
  AMMSB_LAUNCH_CHECK();
  return 0;
}
