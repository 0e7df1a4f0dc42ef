// orderset.cu -- the ITERATION ORDER of libstdc++'s std::unordered_set, computed on the device.
//
// The reference emits every mini-batch in std::unordered_set iteration order (edges:
// sample.cc:267,290; nodes: learner.cc:162-173), so a device sampler is a drop-in only if it
// reproduces that order.  host/mcmc/std_order_set.h derives it from two facts about libstdc++'s
// hashtable (std::hash of an integer is the identity):
//   (1) feeding the sequence S to an empty table of B buckets leaves the node list as
//         R(S, B) = reverse( S stably grouped by bucket, groups in order of first appearance ),
//       and a rehash of list L to B' buckets followed by the inserts T gives R(L ++ T, B');
//   (2) when a rehash happens, and to how many buckets, depends only on the element count
//       (std::__detail::_Prime_rehash_policy -- the very object is replayed here, on the host).
// R(S, B) is a sort: descending by (first index at which the key's bucket appears, index).  So the
// order is ~log2(n) growth phases, each one "bucket of every key, minimum index per bucket, sort":
// short phases run in one CTA (rank by counting), long ones through a device radix sort.
//
//   ammsb_orderset_apply    keys in insertion order (no duplicates) -> keys in iteration order
//   ammsb_minibatch_finish  the tail of a device mini-batch strategy: the drawn edges (insertion
//                           order) -> edges in the reference's order, then
//                           ExtractNodesFromMiniBatch (learner.cc:162-173): the endpoints u, v of
//                           every edge in that order, first occurrences kept, in the iteration
//                           order of std::unordered_set<Vertex>.
#include <cub/cub.cuh>

#include <unordered_set>
#include <utility>
#include <vector>

#include "common.cuh"

struct ammsb_orderset {
  ammsb_ctx* ctx = nullptr;
  uint32_t max_n = 0, max_buckets = 0;
  uint64_t *cur = nullptr, *nxt = nullptr, *sk = nullptr, *sk2 = nullptr;
  uint32_t* first = nullptr;
  // node extraction
  uint64_t *pairs = nullptr, *flags = nullptr;
  uint32_t *tab_v = nullptr, *tab_i = nullptr, *count = nullptr;
  uint32_t cap = 0;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  uint32_t* h_count = nullptr;  // pinned
};

// the growth phases of an unordered_set that receives n distinct keys: (elements in the table when
// the phase ends, bucket count during the phase) -- std_order_set.h ComputeOrder()
static std::vector<std::pair<uint32_t, uint32_t>> orderset_phases(uint32_t n) {
  std::vector<std::pair<uint32_t, uint32_t>> ph;
  std::__detail::_Prime_rehash_policy policy;
  size_t buckets = 1, phase_buckets = 0, have = 0, i = 0;
  auto close = [&](size_t upto) {
    if (phase_buckets == 0 || upto == have) return;
    ph.emplace_back((uint32_t)upto, (uint32_t)phase_buckets);
    have = upto;
  };
  while (i < n) {
    if (i + 1 > policy._M_next_resize) {
      const std::pair<bool, size_t> grow = policy._M_need_rehash(buckets, i, 1);
      if (grow.first) {
        close(i);
        buckets = phase_buckets = grow.second;
      }
    }
    i = std::max<size_t>(i + 1, policy._M_next_resize);
  }
  close(n);
  return ph;
}

struct ModMagic {  // exact a % d (Lemire et al.), d fixed per phase
  uint64_t m_hi, m_lo, d;
};
static ModMagic mod_magic(uint64_t d) {
  const unsigned __int128 m = ~static_cast<unsigned __int128>(0) / d + 1;
  return ModMagic{(uint64_t)(m >> 64), (uint64_t)m, d};
}
__device__ __forceinline__ uint32_t mod_apply(const ModMagic& g, uint64_t a) {
  const uint64_t low_lo = g.m_lo * a;
  const uint64_t low_hi = __umul64hi(g.m_lo, a) + g.m_hi * a;
  const uint64_t bottom_hi = __umul64hi(low_lo, g.d);
  const uint64_t top_lo = low_hi * g.d, top_hi = __umul64hi(low_hi, g.d);
  return (uint32_t)(top_hi + ((top_lo + bottom_hi) < top_lo ? 1 : 0));
}

// ---- short phase: one CTA.  list = cur[0, have) ++ seq[have, upto); out = R(list, buckets) ----
#define OS_SMALL 2048
__global__ void __launch_bounds__(1024) k_orderset_small(const uint64_t* __restrict__ cur, const uint64_t* __restrict__ seq,
                                                         uint32_t have, uint32_t upto, ModMagic g, uint32_t* __restrict__ first,
                                                         uint64_t* __restrict__ out) {
  __shared__ uint64_t s_key[OS_SMALL];
  __shared__ uint64_t s_sk[OS_SMALL];
  for (uint32_t b = threadIdx.x; b < (uint32_t)g.d; b += blockDim.x) first[b] = 0xffffffffu;
  for (uint32_t i = threadIdx.x; i < upto; i += blockDim.x) s_key[i] = i < have ? cur[i] : seq[i];
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < upto; i += blockDim.x) atomicMin(&first[mod_apply(g, s_key[i])], i);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < upto; i += blockDim.x)
    s_sk[i] = ((uint64_t)first[mod_apply(g, s_key[i])] << 32) | i;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < upto; i += blockDim.x) {  // rank in descending order of the sort key
    const uint64_t mine = s_sk[i];
    uint32_t r = 0;
    for (uint32_t j = 0; j < upto; ++j) r += s_sk[j] > mine;
    out[r] = s_key[i];
  }
}

// ---- long phase ----
__global__ void k_orderset_list(const uint64_t* __restrict__ cur, const uint64_t* __restrict__ seq, uint32_t have,
                                uint32_t upto, ModMagic g, uint32_t* __restrict__ first, uint64_t* __restrict__ list) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= upto) return;
  const uint64_t k = i < have ? cur[i] : seq[i];
  list[i] = k;
  atomicMin(&first[mod_apply(g, k)], i);
}
__global__ void k_orderset_sortkey(const uint64_t* __restrict__ list, uint32_t upto, ModMagic g,
                                   const uint32_t* __restrict__ first, uint64_t* __restrict__ sk) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= upto) return;
  sk[i] = ((uint64_t)first[mod_apply(g, list[i])] << 32) | i;
}

static int orderset_apply(ammsb_orderset* s, ammsb_ctx* c, const uint64_t* d_keys, uint32_t n, uint64_t* d_out) {
  AMMSB_REQUIRE(n <= s->max_n, "ordered set larger than the workspace");
  if (n == 0) return 0;
  const auto phases = orderset_phases(n);
  uint64_t* cur = s->cur;
  uint64_t* nxt = s->nxt;
  uint32_t have = 0;
  for (size_t k = 0; k < phases.size(); ++k) {
    const uint32_t upto = phases[k].first, buckets = phases[k].second;
    AMMSB_REQUIRE(buckets <= s->max_buckets, "ordered set: bucket count exceeds the workspace");
    const ModMagic g = mod_magic(buckets);
    uint64_t* dst = (k + 1 == phases.size()) ? d_out : nxt;
    if (upto <= OS_SMALL) {
      k_orderset_small<<<1, 1024, 0, c->stream>>>(cur, d_keys, have, upto, g, s->first, dst);
      g_launch_count.fetch_add(1);
    } else {
      AMMSB_CHECK_CUDA(cudaMemsetAsync(s->first, 0xff, 4 * (size_t)buckets, c->stream));
      k_orderset_list<<<(upto + 255) / 256, 256, 0, c->stream>>>(cur, d_keys, have, upto, g, s->first, s->sk2);
      k_orderset_sortkey<<<(upto + 255) / 256, 256, 0, c->stream>>>(s->sk2, upto, g, s->first, s->sk);
      // sort keys are (first index of the bucket, index): both below upto
      int bits = 1;
      while ((1u << bits) < upto) ++bits;
      size_t tb = s->tmp_bytes;
      // values: the list itself (sk2); sorted keys land in the tail of the value scratch
      AMMSB_CHECK_CUDA(cub::DeviceRadixSort::SortPairsDescending(s->d_tmp, tb, s->sk, s->sk + s->max_n, s->sk2, dst, (int)upto,
                                                                 0, 32 + bits, c->stream));
      g_launch_count.fetch_add(3);
    }
    AMMSB_CHECK_CUDA(cudaGetLastError());
    if (dst != d_out) std::swap(cur, nxt);
    have = upto;
  }
  return 0;
}

extern "C" int ammsb_orderset_destroy(ammsb_orderset* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  cudaFree(s->cur); cudaFree(s->nxt); cudaFree(s->sk); cudaFree(s->sk2); cudaFree(s->first);
  cudaFree(s->pairs); cudaFree(s->flags); cudaFree(s->tab_v); cudaFree(s->tab_i); cudaFree(s->count);
  cudaFree(s->d_tmp);
  cudaFreeHost(s->h_count);
  delete s;
  return 0;
}

extern "C" int ammsb_orderset_create(ammsb_ctx* c, uint32_t max_keys, ammsb_orderset** out) {
  AMMSB_REQUIRE(max_keys >= 1 && max_keys < (1u << 28), "ordered set capacity out of range");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_orderset* s = new ammsb_orderset();
  s->ctx = c;
  s->max_n = max_keys;
  const auto ph = orderset_phases(max_keys);
  s->max_buckets = ph.empty() ? 16 : ph.back().second;
  for (const auto& p : ph) s->max_buckets = std::max(s->max_buckets, p.second);
  s->cap = 1;
  while (s->cap < 2 * max_keys) s->cap <<= 1;
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, t1, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                            (uint64_t*)nullptr, (int)max_keys, 0, 64, c->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, t2, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)max_keys, c->stream);
  s->tmp_bytes = std::max(t1, t2);
  const bool ok = cudaMalloc((void**)&s->cur, 8 * (size_t)max_keys) == cudaSuccess &&
                  cudaMalloc((void**)&s->nxt, 8 * (size_t)max_keys) == cudaSuccess &&
                  cudaMalloc((void**)&s->sk, 16 * (size_t)max_keys) == cudaSuccess &&
                  cudaMalloc((void**)&s->sk2, 8 * (size_t)max_keys) == cudaSuccess &&
                  cudaMalloc((void**)&s->first, 4 * (size_t)s->max_buckets) == cudaSuccess &&
                  cudaMalloc((void**)&s->pairs, 8 * (size_t)max_keys) == cudaSuccess &&
                  cudaMalloc((void**)&s->flags, 8 * (size_t)max_keys) == cudaSuccess &&
                  cudaMalloc((void**)&s->tab_v, 4 * (size_t)s->cap) == cudaSuccess &&
                  cudaMalloc((void**)&s->tab_i, 4 * (size_t)s->cap) == cudaSuccess &&
                  cudaMalloc((void**)&s->count, 16) == cudaSuccess &&
                  cudaMalloc(&s->d_tmp, s->tmp_bytes ? s->tmp_bytes : 8) == cudaSuccess &&
                  cudaMallocHost((void**)&s->h_count, 16) == cudaSuccess;
  if (!ok) {
    ammsb_orderset_destroy(s);
    cudaGetLastError();
    AMMSB_REQUIRE(false, "out of memory for the ordered-set workspace");
  }
  *out = s;
  return 0;
}

extern "C" int ammsb_orderset_apply(ammsb_orderset* s, ammsb_ctx* c, const uint64_t* d_keys, uint32_t n, uint64_t* d_out) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  return orderset_apply(s, c, d_keys, n, d_out);
}

// ---- ExtractNodesFromMiniBatch ----
// endpoint sequence u0, v0, u1, v1, ... ; the first occurrence of a vertex wins (smallest position)
__global__ void k_nodes_mark(const uint64_t* __restrict__ edges, uint32_t E, uint32_t cap_mask, uint32_t* tab_v, uint32_t* tab_i,
                             uint32_t* slot_of) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * E) return;
  const uint64_t e = edges[i >> 1];
  const uint32_t v = (i & 1) ? (uint32_t)(e & 0xffffffffu) : (uint32_t)(e >> 32);
  uint32_t h = (v * 2654435761u) & cap_mask;
  for (;;) {
    const uint32_t old = atomicCAS(&tab_v[h], 0xffffffffu, v);
    if (old == 0xffffffffu || old == v) break;
    h = (h + 1) & cap_mask;
  }
  atomicMin(&tab_i[h], i);
  slot_of[i] = h;
}
__global__ void k_nodes_flag(uint32_t E, const uint32_t* __restrict__ tab_i, const uint32_t* __restrict__ slot_of, uint64_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * E) return;
  flags[i] = tab_i[slot_of[i]] == i ? 1ull : 0ull;
}
__global__ void k_nodes_compact(const uint64_t* __restrict__ edges, uint32_t E, const uint32_t* __restrict__ tab_i,
                                const uint32_t* __restrict__ slot_of, const uint64_t* __restrict__ pos, uint64_t* seq,
                                uint32_t* count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * E) return;
  const bool first = tab_i[slot_of[i]] == i;
  if (first) {
    const uint64_t e = edges[i >> 1];
    seq[pos[i]] = (i & 1) ? (e & 0xffffffffull) : (e >> 32);
  }
  if (i == 2 * E - 1) count[0] = (uint32_t)pos[i] + (first ? 1 : 0);
}
__global__ void k_narrow(const uint64_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)in[i];
}

// d_edges: E edges in INSERTION order on entry, in the reference's emission order on return;
// d_nodes: the mini-batch nodes in the reference's order; *num_nodes their count.  Waits for the
// stream (the node count sizes the kernels that follow).
extern "C" int ammsb_minibatch_finish(ammsb_orderset* s, ammsb_ctx* c, uint64_t* d_edges, uint32_t E, uint32_t* d_nodes,
                                      uint32_t* num_nodes) {
  AMMSB_REQUIRE(E >= 1 && 2 * (uint64_t)E <= s->max_n, "mini-batch larger than the ordered-set workspace");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  // edges: iteration order of unordered_set<Edge> (sample.cc:267,290)
  if (orderset_apply(s, c, d_edges, E, s->pairs)) return 1;
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(d_edges, s->pairs, 8 * (size_t)E, cudaMemcpyDeviceToDevice, c->stream));
  // nodes: first occurrences of the endpoints, in order
  AMMSB_CHECK_CUDA(cudaMemsetAsync(s->tab_v, 0xff, 4 * (size_t)s->cap, c->stream));
  AMMSB_CHECK_CUDA(cudaMemsetAsync(s->tab_i, 0xff, 4 * (size_t)s->cap, c->stream));
  uint32_t* slot_of = reinterpret_cast<uint32_t*>(s->sk2);  // scratch: 2E u32
  const uint32_t blocks = (2 * E + 255) / 256;
  k_nodes_mark<<<blocks, 256, 0, c->stream>>>(d_edges, E, s->cap - 1, s->tab_v, s->tab_i, slot_of);
  k_nodes_flag<<<blocks, 256, 0, c->stream>>>(E, s->tab_i, slot_of, s->flags);
  size_t tb = s->tmp_bytes;
  AMMSB_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(s->d_tmp, tb, s->flags, s->flags, (int)(2 * E), c->stream));
  k_nodes_compact<<<blocks, 256, 0, c->stream>>>(d_edges, E, s->tab_i, slot_of, s->flags, s->pairs, s->count);
  g_launch_count.fetch_add(4);
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(s->h_count, s->count, 4, cudaMemcpyDeviceToHost, c->stream));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  const uint32_t V = s->h_count[0];
  // iteration order of unordered_set<Vertex> (learner.cc:162-173); keys widened to 64 bits
  uint64_t* ordered = s->flags;  // free again
  if (orderset_apply(s, c, s->pairs, V, ordered)) return 1;
  k_narrow<<<(V + 255) / 256, 256, 0, c->stream>>>(ordered, V, d_nodes);
  AMMSB_LAUNCH_CHECK();
  *num_nodes = V;
  return 0;
}
