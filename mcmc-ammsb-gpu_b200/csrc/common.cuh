// common.cuh -- shared host/device plumbing for the sm_100a kernels.
#ifndef AMMSB_COMMON_CUH_
#define AMMSB_COMMON_CUH_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/ammsb.h"
#include "vmm.h"
#include "zig_tables.h"

// ---------------------------------------------------------------- host side --

void ammsb_set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;

#define AMMSB_CHECK_CUDA(expr)                                                       \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ammsb_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                      __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

#define AMMSB_REQUIRE(cond, msg)                                            \
  do {                                                                      \
    if (!(cond)) {                                                          \
      ammsb_set_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" + \
                      std::to_string(__LINE__) + ")");                      \
      return 1;                                                             \
    }                                                                       \
  } while (0)

#define AMMSB_LAUNCH_CHECK()                 \
  do {                                       \
    g_launch_count.fetch_add(1);             \
    AMMSB_CHECK_CUDA(cudaGetLastError());    \
  } while (0)

struct ammsb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  char name[256] = {0};
};

struct ammsb_rng {
  ammsb_ctx* ctx = nullptr;
  ulonglong2* d_state = nullptr;
  uint64_t n = 0;
};

// Device-visible view of a cuckoo set (reference: struct Set, cuckoo.cc:17-23).
struct SetView {
  const uint64_t* base;  // [2][num_bins][4]
  uint64_t num_bins;
  uint64_t p1, p2;  // SET_PRIMES[2*idx], SET_PRIMES[2*idx+1]
  uint64_t m_hi, m_lo;  // ceil(2^128 / num_bins): exact remainder without a divide (set_mod)
};

struct ammsb_set {
  ammsb_ctx* ctx = nullptr;
  uint64_t* d_table = nullptr;
  uint64_t num_bins = 0;
  uint32_t prime_idx = 0;
  SetView view() const;
};

// Device-visible view of the node-partitioned pi/phi store (successor of struct
// TTRowPartitionedMatrix, partitioned-alloc.h:14-29: blocks_[32] -> one base per GPU).
struct StoreView {
  float* pi[AMMSB_MAX_SHARDS];
  float* phi[AMMSB_MAX_SHARDS];
  uint32_t rows_per_shard;
  uint32_t num_shards;
  uint32_t K;
  uint32_t N;
  // replicated mode: full copies on other GPUs that receive every row update
  float* mirror_pi[AMMSB_MAX_SHARDS];
  float* mirror_phi[AMMSB_MAX_SHARDS];
  uint32_t num_mirrors;
};

struct ammsb_store {
  ammsb_ctx* ctx = nullptr;
  uint64_t N = 0;
  uint32_t K = 0;
  uint32_t num_shards = 1, shard_id = 0;
  uint64_t rows_per_shard = 0, first_row = 0, local_rows = 0;
  float* d_pi = nullptr;   // local shard [local_rows, K]
  float* d_phi = nullptr;  // local shard [local_rows]
  float* peer_pi[AMMSB_MAX_SHARDS] = {nullptr};
  float* peer_phi[AMMSB_MAX_SHARDS] = {nullptr};
  bool peer_is_ipc[AMMSB_MAX_SHARDS] = {false};
  bool owns_phi = true;
  float* mirror_pi[AMMSB_MAX_SHARDS] = {nullptr};
  float* mirror_phi[AMMSB_MAX_SHARDS] = {nullptr};
  bool mirror_is_ipc[AMMSB_MAX_SHARDS] = {false};
  uint32_t num_mirrors = 0;
  // shareable stores (virtual-memory API): own allocations and the peers' imported ones
  bool shareable = false;
  VmmAlloc vmm_pi, vmm_phi;
  VmmAlloc vmm_peer_pi[AMMSB_MAX_SHARDS], vmm_peer_phi[AMMSB_MAX_SHARDS];
  VmmAlloc vmm_mirror_pi[AMMSB_MAX_SHARDS], vmm_mirror_phi[AMMSB_MAX_SHARDS];
  StoreView view() const;
};

// -------------------------------------------------------------- device side --
#ifdef __CUDACC__

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ float* store_row(const StoreView& s, uint32_t row) {
  // partitioned-alloc.h:22-28
  uint32_t shard = 0, off = row;
  if (s.num_shards > 1) {
    shard = row / s.rows_per_shard;
    off = row - shard * s.rows_per_shard;
  }
  return s.pi[shard] + (size_t)off * s.K;
}
__device__ __forceinline__ float* store_phi(const StoreView& s, uint32_t row) {
  uint32_t shard = 0, off = row;
  if (s.num_shards > 1) {
    shard = row / s.rows_per_shard;
    off = row - shard * s.rows_per_shard;
  }
  return s.phi[shard] + off;
}

// learner.cc:26-28
__device__ __forceinline__ uint64_t make_edge(uint32_t u, uint32_t v) {
  return ((uint64_t)u << 32) | (uint64_t)v;
}

// Set_HasEdge, cuckoo.cc:39-65.  Two independent 32-byte bin reads (2 x 128-bit
// loads each), issued before either compare so both are in flight together.
// a % num_bins, exactly, for every 64-bit a (Lemire, Kaser, Kurz: "Faster remainder by direct
// computation", 2019, with 128 fractional bits): six 64-bit multiplies instead of the ~120
// instruction software divide -- the two hashes of a lookup are a fixed cost per sampled neighbor.
__device__ __forceinline__ uint64_t set_mod(const SetView& s, uint64_t a) {
  const uint64_t low_lo = s.m_lo * a;                                  // (m * a) mod 2^128
  const uint64_t low_hi = __umul64hi(s.m_lo, a) + s.m_hi * a;
  const uint64_t bottom_hi = __umul64hi(low_lo, s.num_bins);           // floor(low * d / 2^128)
  const uint64_t top_lo = low_hi * s.num_bins, top_hi = __umul64hi(low_hi, s.num_bins);
  return top_hi + ((top_lo + bottom_hi) < top_lo ? 1 : 0);
}

__device__ __forceinline__ bool set_has(const SetView& s, uint64_t k) {
  uint64_t h1 = set_mod(s, s.p1 * k);
  uint64_t h2 = set_mod(s, k ^ s.p2);
  const ulonglong2* b1 = reinterpret_cast<const ulonglong2*>(s.base + h1 * 4);
  const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(s.base + (s.num_bins + h2) * 4);
  ulonglong2 a0 = __ldg(b1), a1 = __ldg(b1 + 1);
  ulonglong2 c0 = __ldg(b2), c1 = __ldg(b2 + 1);
  return (a0.x == k) | (a0.y == k) | (a1.x == k) | (a1.y == k) |
         (c0.x == k) | (c0.y == k) | (c1.x == k) | (c1.y == k);
}

// ---- RNG: random.cl.inc:13-49 ----
struct Rng {
  uint64_t x, y;
};
__device__ __forceinline__ Rng rng_load(const ulonglong2* pool, uint64_t i) {
  ulonglong2 v = pool[i];
  Rng r;
  r.x = v.x;
  r.y = v.y;
  return r;
}
__device__ __forceinline__ void rng_store(ulonglong2* pool, uint64_t i, const Rng& r) {
  pool[i] = make_ulonglong2(r.x, r.y);
}
__device__ __forceinline__ uint64_t rng_next(Rng& s) {  // xorshift_128plus
  uint64_t s1 = s.x;
  const uint64_t s0 = s.y;
  s.x = s0;
  s1 ^= s1 << 23;
  s.y = s1 ^ s0 ^ (s1 >> 17) ^ (s0 >> 26);
  return s.y + s0;
}
// random(): FL(1.0) * rand / ULONG_MAX.  (float)ULONG_MAX == 2^64, so this is the
// round-to-nearest u64->f32 conversion scaled by an exact power of two.
__device__ __forceinline__ float rng_uniform(Rng& s) {
  return __ull2float_rn(rng_next(s)) * 5.42101086242752217e-20f;  // 2^-64
}

// ---- NeighborSampler building blocks (sample.cc:13-77), shared by small_kernels.cu / cols.cu ----
// generate_random_int draws `rand % N` from a 64-bit xorshift value and probes an open-addressing
// table with `(l1 + q * (1 + 2 * capacity)) % capacity`.  The same values, without a divide: the
// 64-bit remainder by N is Lemire's multiply form (as set_mod), `x % capacity` a mask or the 32-bit
// multiply form, and the probe sequence is l1, l1 + 1, ... (mod capacity) because
// 1 + 2 * capacity = 1 (mod capacity).
struct NsGeom {
  uint64_t n_m_hi, n_m_lo;  // ceil(2^128 / N)
  uint64_t cap_m;           // ceil(2^64 / capacity)
  uint32_t N, n, capacity, cap_pow2;
};
static inline NsGeom ns_geom(uint32_t N, uint32_t n) {
  NsGeom g;
  const unsigned __int128 m = ~static_cast<unsigned __int128>(0) / N + 1;
  g.n_m_hi = static_cast<uint64_t>(m >> 64);
  g.n_m_lo = static_cast<uint64_t>(m);
  g.N = N;
  g.n = n;
  g.capacity = 2 * n;
  g.cap_m = ~0ull / g.capacity + 1;
  g.cap_pow2 = (g.capacity & (g.capacity - 1)) == 0;
  return g;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ns_mod_N(const NsGeom& g, uint64_t a) {
  const uint64_t low_lo = g.n_m_lo * a;
  const uint64_t low_hi = __umul64hi(g.n_m_lo, a) + g.n_m_hi * a;
  const uint64_t bottom_hi = __umul64hi(low_lo, (uint64_t)g.N);
  const uint64_t top_lo = low_hi * g.N, top_hi = __umul64hi(low_hi, (uint64_t)g.N);
  return (uint32_t)(top_hi + ((top_lo + bottom_hi) < top_lo ? 1 : 0));
}
// one slot: n distinct ids != node into the table tab[j * stride], j < capacity (empty = N)
__device__ __forceinline__ void ns_draw_slot(Rng& seed, uint32_t node, const NsGeom& g, uint32_t* tab, uint32_t stride) {
  const uint32_t cap = g.capacity, N = g.N;
  for (uint32_t j = 0; j < cap; ++j) tab[j * stride] = N;
  for (uint32_t j = 0; j < g.n; ++j) {
    for (;;) {
      uint32_t r;
      do {
        r = ns_mod_N(g, rng_next(seed));  // randint(seed, 0, N - 1), random.cl.inc:37-39
      } while (r == node);
      const uint32_t h = r ^ 553105253u;
      uint32_t off = g.cap_pow2 ? (h & (cap - 1)) : (uint32_t)__umul64hi(g.cap_m * h, (uint64_t)cap);
      bool dup = false;
      for (;;) {
        const uint32_t val = tab[off * stride];
        if (val == r) {
          dup = true;
          break;
        }
        if (val == N) {
          tab[off * stride] = r;
          break;
        }
        off = off + 1 == cap ? 0 : off + 1;
      }
      if (!dup) break;
    }
  }
}
#endif

__device__ const uint32_t d_zig_ytab[128] = AMMSB_ZIG_YTAB_BITS_INIT;
__device__ const uint32_t d_zig_ktab[128] = AMMSB_ZIG_KTAB_INIT;
__device__ const uint32_t d_zig_wtab[128] = AMMSB_ZIG_WTAB_BITS_INIT;

// Where the three 128-entry ziggurat tables are read from.  ZigGlobal: the __device__ arrays through
// the read-only path.  ZigShared: a per-CTA copy (ytab | ktab | wtab) for the gather warps that draw
// their own noise (K <= 512: measured +1..3 % together with drawing at the start of a slot); the
// noise-producer warps of the K >= 1024 kernels keep ZigGlobal, shared memory being the busy
// resource there (K = 1024: 0.3838 ms with the shared copy, 0.3798 ms without).
#define ZIG_WORDS 384
struct ZigGlobal {
  __device__ __forceinline__ uint32_t y(uint32_t i) const { return __ldg(&d_zig_ytab[i]); }
  __device__ __forceinline__ uint32_t k(uint32_t i) const { return __ldg(&d_zig_ktab[i]); }
  __device__ __forceinline__ uint32_t w(uint32_t i) const { return __ldg(&d_zig_wtab[i]); }
};
struct ZigShared {
  const uint32_t* t;
  __device__ __forceinline__ uint32_t y(uint32_t i) const { return t[i]; }
  __device__ __forceinline__ uint32_t k(uint32_t i) const { return t[128 + i]; }
  __device__ __forceinline__ uint32_t w(uint32_t i) const { return t[256 + i]; }
};
// every thread of the CTA takes part; the caller synchronises the CTA before the first draw
__device__ __forceinline__ void zig_stage(uint32_t* s_zig) {
  for (uint32_t i = threadIdx.x; i < ZIG_WORDS; i += blockDim.x)
    s_zig[i] = i < 128 ? d_zig_ytab[i] : (i < 256 ? d_zig_ktab[i - 128] : d_zig_wtab[i - 256]);
}

// gsl_ran_gaussian_ziggurat, random.cl.inc:221-274 (sigma = 1).  The wedge/tail
// tests use the precise expf/logf (never the fast-math intrinsics) so that the
// accept/reject decisions -- and with them the stream position -- agree with the
// reference evaluated on an IEEE host.
template <class Tab>
__device__ __forceinline__ float rng_randn_t(Rng& s, const Tab& tab) {
  const float R = 3.44428647676f;
  float x;
  uint32_t sign;
  for (;;) {
    const uint64_t k = rng_next(s);
    uint32_t i = (uint32_t)k & 0xFFu;
    const uint32_t j = (uint32_t)(k >> 8) & 0xFFFFFFu;
    sign = i & 0x80u;
    i &= 0x7fu;
    x = __fmul_rn((float)j, __uint_as_float(tab.w(i)));
    if (j < tab.k(i)) break;
    float y;
    if (i < 127) {
      const float y0 = __uint_as_float(tab.y(i));
      const float y1 = __uint_as_float(tab.y(i + 1));
      const float U1 = rng_uniform(s);
      y = __fadd_rn(y1, __fmul_rn(__fsub_rn(y0, y1), U1));
    } else {
      const float U1 = __fsub_rn(1.0f, rng_uniform(s));
      const float U2 = rng_uniform(s);
      x = __fsub_rn(R, __fdiv_rn(logf(U1), R));
      y = __fmul_rn(expf(__fmul_rn(-R, __fsub_rn(x, __fmul_rn(0.5f, R)))), U2);
    }
    if (y < expf(__fmul_rn(__fmul_rn(-0.5f, x), x))) break;
  }
  return sign ? x : -x;
}
__device__ __forceinline__ float rng_randn(Rng& s) { return rng_randn_t(s, ZigGlobal()); }

__device__ __forceinline__ float rng_uniform_pos(Rng& s) {  // random.cl.inc:311-318
  float x;
  do {
    x = rng_uniform(s);
  } while (x == 0.0f);
  return x;
}

// gsl_ran_gamma, random.cl.inc:353-391.  IEEE ops, no contraction: the comparisons
// steer the stream.
__device__ __forceinline__ float rng_gamma(Rng& s, float a, float b) {
  float f = 1.0f;
  while (a < 1.0f) {
    const float u = rng_uniform_pos(s);
    f = __fmul_rn(f, powf(u, __fdiv_rn(1.0f, a)));
    a = __fadd_rn(1.0f, a);
  }
  const float d = __fsub_rn(a, __fdiv_rn(1.0f, 3.0f));
  const float c = __fdiv_rn(__fdiv_rn(1.0f, 3.0f), __fsqrt_rn(d));
  float x, v, u;
  for (;;) {
    do {
      x = rng_randn(s);
      v = __fadd_rn(1.0f, __fmul_rn(c, x));
    } while (v <= 0.0f);
    v = __fmul_rn(__fmul_rn(v, v), v);
    u = rng_uniform_pos(s);
    const float x2 = __fmul_rn(x, x);
    // 1 - 0.0331f * x * x * x * x
    const float t = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.0331f, x), x), x), x);
    if (u < __fsub_rn(1.0f, t)) break;
    // log(u) < 0.5 * x * x + d * (1 - v + log(v))
    const float rhs = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, x), x),
                                __fmul_rn(d, __fadd_rn(__fsub_rn(1.0f, v), logf(v))));
    (void)x2;
    if (logf(u) < rhs) break;
  }
  return __fmul_rn(__fmul_rn(__fmul_rn(f, b), d), v);
}

// get_eps_t, learner.cc:41-43: EPS_A * pow(1 + step_count / EPS_B, -EPS_C)
__device__ __forceinline__ float eps_t_of(float a, float b, float c, uint32_t step) {
  return __fmul_rn(a, powf(__fadd_rn(1.0f, __fdiv_rn((float)step, b)), -c));
}

// The Langevin step of one phi element (phi.cc:266-274) as ONE expression with explicit FMA
// contraction, shared by every production update_phi kernel (single-GPU and column-sharded), so
// that they round identically: phi' = max(|phi + eps/2 (alpha - phi + N/n g) + sqrt(eps phi) xi|, 1e-24)
__device__ __forceinline__ float phi_langevin(float pi_k, float phi_sum, float g, float noise, float half_eps,
                                              float eps_t, float alpha, float Nn) {
  const float phi_k = pi_k * phi_sum;
  const float drift = fmaf(Nn, g, alpha - phi_k);
  const float v = fmaf(sqrtf(eps_t * phi_k), noise, fmaf(half_eps, drift, phi_k));
  return fmaxf(fabsf(v), 1e-24f);
}

// Butterfly all-reduce.  For every lane the result has the association of the
// reference's WG_SUM tree with WG_SIZE 32 (sum.cc:20-29: aux[l] += aux[l+p2],
// p2 = 16,8,4,2,1) because fp32 addition is commutative.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// 128-bit streaming loads for rows that are read exactly once.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared::cta bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

#endif  // __CUDACC__
#endif  // AMMSB_COMMON_CUH_
