/*
 * oracle/ammsb_oracle.c -- TEST INFRASTRUCTURE ONLY.  See ammsb_oracle.h.
 *
 * Plain-C restatement of the reference's device kernels, evaluated work-item by
 * work-item in the reference's own association order.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math  (see oracle/Makefile)
 * so that every fp32 expression is evaluated exactly as written.
 */
#include "ammsb_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "zig_tables.h"

#define MAX_GROUPS 65535u /* types.cc:537 */

static const uint32_t kYtabBits[128] = AMMSB_ZIG_YTAB_BITS_INIT;
static const uint64_t kKtab[128] = AMMSB_ZIG_KTAB_INIT;
static const uint32_t kWtabBits[128] = AMMSB_ZIG_WTAB_BITS_INIT;

static inline float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1: the timing legs set the count explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* config.cc:57-64: float -> "%e" text -> float literal */
float orc_round_param(float f) {
  char buf[64];
  snprintf(buf, sizeof buf, "%e", (double)f);
  return strtof(buf, NULL);
}

/* learner.cc:41-43: EPS_A * pow(1 + step_count / EPS_B, -EPS_C) */
float orc_eps_t(const orc_params* p, uint32_t step) {
  return p->a * powf(1 + step / p->b, -p->c);
}

/* ---------------------------------------------------------------- RNG ---- */

/* random.cc:31-44 */
void orc_rng_init(orc_rng* pool, uint64_t n, uint64_t sx, uint64_t sy) {
  for (uint64_t i = 0; i < n; ++i) { pool[i].x = sx + i; pool[i].y = sy + i; }
}

/* random.cl.inc:13-25 */
uint64_t orc_rand(orc_rng* s) {
  uint64_t s1 = s->x;
  uint64_t s0 = s->y;
  s->x = s0;
  s1 ^= s1 << 23;
  s->y = s1 ^ s0 ^ (s1 >> 17) ^ (s0 >> 26);
  return s->y + s0;
}

/* random.cl.inc:34-35: FL(1.0) * rand(s) / ULONG_MAX */
float orc_random(orc_rng* s) { return 1.0f * orc_rand(s) / 0xffffffffffffffffUL; }

/* random.cl.inc:37-39 */
int orc_randint(orc_rng* s, int from, int upto) {
  return (orc_rand(s) % (upto + 1 - from)) + from;
}

/* random.cl.inc:221-274 (range >= 0xFFFFFFFF branch) */
float orc_randn(orc_rng* s) {
  const float R = 3.44428647676f; /* PARAM_R, random.cl.inc:4 */
  uint64_t i, j;
  int sign;
  float x, y;
  for (;;) {
    uint64_t k = orc_rand(s);
    i = (k & 0xFF);
    j = (k >> 8) & 0xFFFFFF;
    sign = (i & 0x80) ? +1 : -1;
    i &= 0x7f;
    x = j * bits2f(kWtabBits[i]);
    if (j < kKtab[i]) break;
    if (i < 127) {
      float y0 = bits2f(kYtabBits[i]);
      float y1 = bits2f(kYtabBits[i + 1]);
      float U1 = orc_random(s);
      y = y1 + (y0 - y1) * U1;
    } else {
      float U1 = 1.0f - orc_random(s);
      float U2 = orc_random(s);
      x = R - logf(U1) / R;
      y = expf(-R * (x - 0.5f * R)) * U2;
    }
    if (y < expf(-0.5f * x * x)) break;
  }
  return sign * 1.0f * x;
}

static float uniform_pos(orc_rng* s) { /* random.cl.inc:311-318 */
  float x;
  do { x = orc_random(s); } while (x == 0);
  return x;
}

/* random.cl.inc:353-391 (non-recursive branch) */
float orc_rand_gamma(orc_rng* s, float a, float b) {
  float f = 1.0f;
  while (a < 1) {
    float u = uniform_pos(s);
    f = f * powf(u, 1.0f / a);
    a = 1.0f + a;
  }
  float x, v, u;
  float d = a - 1.0f / 3.0f;
  float c = (1.0f / 3.0f) / sqrtf(d);
  for (;;) {
    do {
      x = orc_randn(s);
      v = 1.0f + c * x;
    } while (v <= 0);
    v = v * v * v;
    u = uniform_pos(s);
    if (u < 1 - 0.0331f * x * x * x * x) break;
    if (logf(u) < 0.5f * x * x + d * (1 - v + logf(v))) break;
  }
  return f * b * d * v;
}

/* ------------------------------------------------------------- cuckoo ---- */

static const uint64_t kPrimes[4][2] = { /* cuckoo.cc:30-35, :92-96 */
    {15485807ull, 920429591ull}, {379906717ull, 740320571ull},
    {256204747ull, 379927517ull}, {13ull, 17ull}};
#define KEY_INVALID 0xffffffffffffffffull

/* cuckoo.cc:98-104 */
uint64_t orc_set_bins_for(uint64_t n) { return (uint64_t)(1 + ceil((1.15 * n) / (2 * 4))); }

static inline uint64_t set_hash(const orc_set* s, uint64_t k, int b) { /* cuckoo.cc:199-209 */
  return b == 0 ? (kPrimes[s->prime_idx][0] * k) % s->num_bins
                : (k ^ kPrimes[s->prime_idx][1]) % s->num_bins;
}
static inline uint64_t* set_slot(const orc_set* s, int b, uint64_t h) {
  return s->table + ((uint64_t)b * s->num_bins + h) * 4;
}

/* cuckoo.cc:140-161 (+ :131-138, :187-197) */
static int set_insert(orc_set* s, uint64_t k, unsigned* seed, uint64_t disp_max) {
  uint64_t displacements = 0;
  do {
    for (int b = 0; b < 2; ++b) {
      uint64_t* slot = set_slot(s, b, set_hash(s, k, b));
      int full = 1, present = 0;
      for (int i = 0; i < 4; ++i) {
        if (slot[i] == KEY_INVALID) full = 0;
        if (slot[i] == k) { present = 1; break; }
      }
      if (!present && !full) {
        for (int i = 0; i < 4; ++i)
          if (slot[i] == KEY_INVALID) { slot[i] = k; break; }
        ++s->count;
        return 1;
      }
    }
    int b = rand_r(seed) % 2;
    uint64_t* slot = set_slot(s, b, set_hash(s, k, b));
    int placed = 0;
    for (int i = 0; i < 4; ++i)
      if (slot[i] == KEY_INVALID) { slot[i] = k; k = KEY_INVALID; placed = 1; break; }
    if (!placed) {
      int alt = rand_r(seed) % 4;
      uint64_t old = slot[alt];
      slot[alt] = k;
      k = old;
    }
  } while (++displacements < disp_max);
  return 0;
}

/* cuckoo.cc:98-129: Set(n) + SetContents */
int orc_set_build(const uint64_t* keys, uint64_t n, orc_set* s) {
  s->num_bins = orc_set_bins_for(n);
  s->count = 0;
  s->table = (uint64_t*)malloc(sizeof(uint64_t) * 2 * 4 * s->num_bins);
  unsigned seed = 42; /* seed_ persists across attempts */
  uint64_t disp_max = n / 2 + 1;
  for (s->prime_idx = 0; s->prime_idx < 4; ++s->prime_idx) {
    memset(s->table, 0xff, sizeof(uint64_t) * 2 * 4 * s->num_bins);
    int ok = 1;
    for (uint64_t i = 0; ok && i < n; ++i) ok = set_insert(s, keys[i], &seed, disp_max);
    if (ok) return 1;
  }
  return 0;
}

void orc_set_free(orc_set* s) { free(s->table); s->table = NULL; }

/* cuckoo.cc:39-65 */
int orc_set_has(const orc_set* s, uint64_t k) {
  for (int b = 0; b < 2; ++b) {
    const uint64_t* slot = set_slot(s, b, set_hash(s, k, b));
    if (k == slot[0] || k == slot[1] || k == slot[2] || k == slot[3]) return 1;
  }
  return 0;
}

void orc_set_has_many(const orc_set* s, const uint64_t* keys, uint64_t n, uint8_t* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)orc_set_has(s, keys[i]);
}

/* --------------------------------------------------- neighbor sampler ---- */

/* sample.cc:15-46 */
static void gen_random_int(orc_rng* seed, uint32_t* out, uint32_t capacity,
                           uint32_t max_id, uint32_t node) {
  uint32_t r, val;
  do {
    do {
      r = (uint32_t)orc_randint(seed, 0, (int)max_id);
    } while (r == node);
    uint32_t l1 = (r ^ 553105253u) % capacity;
    uint32_t l2 = 1 + (capacity << 1);
    for (uint32_t i = 0;; ++i) {
      uint32_t offset = (l1 + i * l2) % capacity;
      val = out[offset];
      if (val == r) break;
      if (val == max_id + 1) { out[offset] = r; break; }
    }
  } while (val == r);
}

/* sample.cc:48-77 kernel; :111-121 launch: global = min(ceil(V/wg), 65535/wg)*wg */
void orc_neighbor_sample(orc_rng* pool, const uint32_t* nodes, uint32_t V, uint32_t N,
                         uint32_t n, uint32_t wg, uint32_t* hash, uint32_t* packed_all) {
  uint32_t capacity = 2 * n;
  uint32_t groups = V / wg + (V % wg ? 1 : 0);
  if (groups > MAX_GROUPS / wg) groups = MAX_GROUPS / wg;
  uint32_t gsize = groups * wg;
#pragma omp parallel for schedule(static)
  for (uint32_t gid = 0; gid < gsize; ++gid) {
    if (gid >= V) continue;
    orc_rng seed = pool[gid];
    for (uint32_t i = gid; i < V; i += gsize) {
      uint32_t* out = hash + (uint64_t)i * capacity;
      uint32_t* packed = packed_all + (uint64_t)i * n;
      uint32_t node = nodes[i];
      for (uint32_t j = 0; j < capacity; ++j) out[j] = N;
      for (uint32_t j = 0; j < n; ++j) gen_random_int(&seed, out, capacity, N - 1, node);
      uint32_t count = 0;
      for (uint32_t j = 0; j < capacity && count < n; ++j)
        if (out[j] != N) packed[count++] = out[j];
    }
    pool[gid] = seed;
  }
}

/* ------------------------------------------------ work-group reductions -- */

static uint32_t power_of_2(uint32_t v) { /* sum.cc:11-18 */
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}

/* sum.cc:20-29, evaluated for all lanes in lock-step */
static void wg_tree_f32(float* aux, uint32_t lsize) {
  for (uint32_t p2 = power_of_2(lsize) >> 1; p2 > 0; p2 >>= 1)
    for (uint32_t lid = 0; lid < p2; ++lid)
      if (lid + p2 < lsize) aux[lid] += aux[lid + p2];
}
static void wg_tree_u32(uint32_t* aux, uint32_t lsize) {
  for (uint32_t p2 = power_of_2(lsize) >> 1; p2 > 0; p2 >>= 1)
    for (uint32_t lid = 0; lid < p2; ++lid)
      if (lid + p2 < lsize) aux[lid] += aux[lid + p2];
}

/* sum.cc:31-42 */
float orc_wg_sum_f32(const float* in, uint32_t len, uint32_t wg) {
  float* aux = (float*)malloc(sizeof(float) * wg);
  for (uint32_t lid = 0; lid < wg; ++lid) {
    float lsum = 0;
    for (uint32_t i = lid; i < len; i += wg) lsum += in[i];
    aux[lid] = lsum;
  }
  wg_tree_f32(aux, wg);
  float r = aux[0];
  free(aux);
  return r;
}
uint32_t orc_wg_sum_u32(const uint32_t* in, uint32_t len, uint32_t wg) {
  uint32_t* aux = (uint32_t*)malloc(sizeof(uint32_t) * wg);
  for (uint32_t lid = 0; lid < wg; ++lid) {
    uint32_t lsum = 0;
    for (uint32_t i = lid; i < len; i += wg) lsum += in[i];
    aux[lid] = lsum;
  }
  wg_tree_u32(aux, wg);
  uint32_t r = aux[0];
  free(aux);
  return r;
}
/* normalize.cc:13-23 */
float orc_wg_normalize_f32(float* in, uint32_t len, uint32_t wg) {
  float sum = orc_wg_sum_f32(in, len, wg);
  for (uint32_t i = 0; i < len; ++i) in[i] = in[i] / sum;
  return sum;
}

/* ---------------------------------------------------------- update_phi --- */

static inline uint64_t make_edge(uint32_t u, uint32_t v) { /* learner.cc:26-28 */
  return (((uint64_t)u) << 32) | v;
}
static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

/* One mini-batch slot.  THREAD: phi.cc:78-122 with states[0]; WG-NAIVE:
 * phi.cc:214-275 with states[lane], lane = k % wg, draw order k / wg. */
static void phi_for_node(int mode, uint32_t wg, const orc_params* p, const float* beta,
                         const float* pi_all, const float* g_phi, float* phi_vec,
                         const orc_set* set, uint32_t node, const uint32_t* neighbors,
                         uint32_t step_count, orc_rng* states, int disable_noise,
                         float* grads, float* probs, float* aux) {
  const uint32_t K = p->K;
  const float EPSILON = p->epsilon;
  const float* pi = pi_all + (uint64_t)node * K;
  float eps_t = orc_eps_t(p, step_count);
  float phi_sum = g_phi[node];
  for (uint32_t k = 0; k < K; ++k) grads[k] = 0;
  for (uint32_t i = 0; i < p->num_neighbors; ++i) {
    uint32_t neighbor = neighbors[i];
    const float* pi_neighbor = pi_all + (uint64_t)neighbor * K;
    uint64_t edge = make_edge(umin(node, neighbor), umax(node, neighbor));
    int y = orc_set_has(set, edge);
    float e = (y == 1 ? EPSILON : 1.0f - EPSILON);
    float probs_sum = 0;
    if (mode == ORC_MODE_THREAD) {
      for (uint32_t k = 0; k < K; ++k) {
        float beta_k = beta[2 * k + 1];
        float f = (y == 1) ? (beta_k - EPSILON) : (EPSILON - beta_k);
        float probs_k = pi[k] * (pi_neighbor[k] * f + e);
        probs_sum += probs_k;
        probs[k] = probs_k;
      }
    } else {
      for (uint32_t k = 0; k < K; ++k) {
        float beta_k = beta[2 * k + 1];
        float f = (y == 1) ? (beta_k - EPSILON) : (EPSILON - beta_k);
        probs[k] = pi[k] * (pi_neighbor[k] * f + e);
      }
      for (uint32_t lid = 0; lid < wg; ++lid) {
        float ps = 0;
        for (uint32_t k = lid; k < K; k += wg) ps += probs[k];
        aux[lid] = ps;
      }
      wg_tree_f32(aux, wg);
      probs_sum = aux[0];
    }
    for (uint32_t k = 0; k < K; ++k)
      grads[k] += (probs[k] / probs_sum) / (pi[k] * phi_sum) - 1.0f / phi_sum;
  }
  float Nn = (1.0f * p->N) / p->num_neighbors;
  /* noise in the reference's draw order: per lane, ascending k */
  float* noise = probs; /* reuse */
  if (disable_noise) {
    for (uint32_t k = 0; k < K; ++k) noise[k] = 1; /* phi.cc:673-677 */
  } else if (mode == ORC_MODE_THREAD) {
    for (uint32_t k = 0; k < K; ++k) noise[k] = orc_randn(&states[0]);
  } else {
    for (uint32_t lid = 0; lid < wg; ++lid)
      for (uint32_t k = lid; k < K; k += wg) noise[k] = orc_randn(&states[lid]);
  }
  for (uint32_t k = 0; k < K; ++k) {
    float phi_k = pi[k] * phi_sum;
    float phi_vec_k = fabsf(phi_k + eps_t / 2 * (p->alpha - phi_k + Nn * grads[k]) +
                            sqrtf(eps_t * phi_k) * noise[k]);
    phi_vec[k] = fmaxf(phi_vec_k, 1e-24f);
  }
}

/* phi.cc:124-152 / :277-302 kernels; :728-757 launch geometry.
 * THREAD: global = min(ceil(V/wg),65535)*wg work-items, item i serves slots
 * i, i+global, ...; state index i.  WG: G = min(V,65535) groups, group g serves
 * slots g, g+G, ...; state index g*wg + lane. */
void orc_update_phi(int mode, uint32_t wg, const orc_params* p, const float* beta,
                    const float* pi, const float* phi, const orc_set* train,
                    const uint32_t* nodes, const uint32_t* neighbors, uint32_t V,
                    uint32_t step_count, orc_rng* pool, int disable_noise, float* phi_vec) {
  const uint32_t K = p->K;
  uint32_t units, per_unit_states;
  if (mode == ORC_MODE_THREAD) {
    uint32_t g = V / wg + (V % wg ? 1 : 0);
    if (g > MAX_GROUPS) g = MAX_GROUPS;
    units = g * wg;
    per_unit_states = 1;
  } else {
    units = V < MAX_GROUPS ? V : MAX_GROUPS;
    per_unit_states = wg;
  }
#pragma omp parallel
  {
    float* grads = (float*)malloc(sizeof(float) * K);
    float* probs = (float*)malloc(sizeof(float) * K);
    float* aux = (float*)malloc(sizeof(float) * (wg ? wg : 1));
    orc_rng* st = (orc_rng*)malloc(sizeof(orc_rng) * per_unit_states);
#pragma omp for schedule(static)
    for (uint32_t u = 0; u < units; ++u) {
      if (u >= V) continue;
      if (pool) memcpy(st, pool + (uint64_t)u * per_unit_states, sizeof(orc_rng) * per_unit_states);
      for (uint32_t i = u; i < V; i += units)
        phi_for_node(mode, wg, p, beta, pi, phi, phi_vec + (uint64_t)i * K, train, nodes[i],
                     neighbors + (uint64_t)i * p->num_neighbors, step_count, st, disable_noise,
                     grads, probs, aux);
      if (pool) memcpy(pool + (uint64_t)u * per_unit_states, st, sizeof(orc_rng) * per_unit_states);
    }
    free(grads); free(probs); free(aux); free(st);
  }
}

/* phi.cc:154-173 (THREAD: serial sum) / :178-197 (WG: copy, WG_NORMALIZE) */
void orc_update_pi(int mode, uint32_t wg, uint32_t K, float* pi_all, float* g_phi,
                   const float* phi_vec, const uint32_t* nodes, uint32_t V) {
  /* slots are visited in launch order so that a node listed twice ends with the
   * value of its last slot, as any in-order execution of the reference would */
  for (uint32_t i = 0; i < V; ++i) {
    uint32_t n = nodes[i];
    float* pi = pi_all + (uint64_t)n * K;
    const float* phi = phi_vec + (uint64_t)i * K;
    float sum;
    if (mode == ORC_MODE_THREAD) {
      sum = 0;
      for (uint32_t k = 0; k < K; ++k) sum += phi[k];
    } else {
      sum = orc_wg_sum_f32(phi, K, wg);
    }
    for (uint32_t k = 0; k < K; ++k) pi[k] = phi[k] / sum;
    g_phi[n] = sum;
  }
}

/* --------------------------------------------------------- update_beta --- */

void orc_theta_to_beta(uint32_t K, const float* theta, float* beta) {
  /* beta.cc:378-379: CopyTo + Normalizer(slice=2, wg=1) -> normalize.cc:13-32 */
  for (uint32_t k = 0; k < K; ++k) {
    float lsum = 0;
    lsum += theta[2 * k];
    lsum += theta[2 * k + 1];
    beta[2 * k] = theta[2 * k] / lsum;
    beta[2 * k + 1] = theta[2 * k + 1] / lsum;
  }
}

void orc_update_beta(int mode, uint32_t wg, const orc_params* p, float* theta, float* beta,
                     const float* pi_all, const orc_set* train, const uint64_t* edges,
                     uint32_t E_mb, float scale, uint32_t step_count, orc_rng* pool,
                     float* theta_sum, float* grads_out) {
  const uint32_t K = p->K;
  const float EPSILON = p->epsilon;
  /* sum_theta: beta.cc:30-37 */
  for (uint32_t k = 0; k < K; ++k) theta_sum[k] = theta[2 * k] + theta[2 * k + 1];

  /* launch geometry: beta.cc:346-357 */
  uint32_t units; /* work-items (THREAD) or groups (WG) that own a partial */
  if (mode == ORC_MODE_THREAD) {
    uint32_t g = E_mb / wg + (E_mb % wg ? 1 : 0);
    if (g > MAX_GROUPS) g = MAX_GROUPS;
    units = g * wg;
  } else {
    units = E_mb < MAX_GROUPS ? E_mb : MAX_GROUPS;
  }
  uint32_t P = units < E_mb ? units : E_mb; /* partials actually written */
  float* partial = (float*)calloc((size_t)P * 2 * K, sizeof(float));
#pragma omp parallel
  {
    float* probs = (float*)malloc(sizeof(float) * K);
    float* aux = (float*)malloc(sizeof(float) * wg);
#pragma omp for schedule(static)
    for (uint32_t u = 0; u < P; ++u) {
      float* grads = partial + (size_t)u * 2 * K;
      for (uint32_t i = u; i < E_mb; i += units) {
        /* beta.cc:105-136 (THREAD) / :195-226 (WG) */
        uint64_t edge = edges[i];
        uint32_t a = (uint32_t)(edge >> 32), b = (uint32_t)(edge & 0xffffffffu);
        edge = make_edge(umin(a, b), umax(a, b));
        uint32_t y = orc_set_has(train, edge) ? 1 : 0;
        const float* pi_a = pi_all + (uint64_t)a * K;
        const float* pi_b = pi_all + (uint64_t)b * K;
        float pi_sum = 0, probs_sum = 0;
        if (mode == ORC_MODE_THREAD) {
          for (uint32_t k = 0; k < K; ++k) {
            float f = pi_a[k] * pi_b[k];
            pi_sum += f;
            float probs_k = y ? beta[2 * k + 1] * f : (1.0f - beta[2 * k + 1]) * f;
            probs[k] = probs_k;
            probs_sum += probs_k;
          }
        } else {
          for (uint32_t lid = 0; lid < wg; ++lid) aux[lid] = 0;
          float* aux2 = (float*)alloca(sizeof(float) * wg);
          for (uint32_t lid = 0; lid < wg; ++lid) {
            float scratch = 0, ps = 0;
            for (uint32_t k = lid; k < K; k += wg) {
              float f = pi_a[k] * pi_b[k];
              scratch += f;
              float beta_k = beta[2 * k + 1];
              float probs_k = y ? beta_k * f : (1.0f - beta_k) * f;
              probs[k] = probs_k;
              ps += probs_k;
            }
            aux[lid] = scratch;
            aux2[lid] = ps;
          }
          wg_tree_f32(aux, wg);
          pi_sum = aux[0];
          wg_tree_f32(aux2, wg);
          probs_sum = aux2[0];
        }
        float prob_0 = (y ? EPSILON : (1.0f - EPSILON)) * (1.0f - pi_sum);
        probs_sum += prob_0;
        for (uint32_t k = 0; k < K; ++k) {
          float f = probs[k] / probs_sum;
          float one_over_theta_sum = 1.0f / theta_sum[k];
          grads[2 * k] += f * ((1 - y) / theta[2 * k] - one_over_theta_sum);
          grads[2 * k + 1] += f * (y / theta[2 * k + 1] - one_over_theta_sum);
        }
      }
    }
    free(probs); free(aux);
  }
  /* sum_grads: beta.cc:39-49, serial over partials */
  for (uint32_t i = 0; i < 2 * K; ++i) {
    float sum = partial[i];
    for (uint32_t q = 1; q < P; ++q) sum += partial[i + (size_t)q * 2 * K];
    grads_out[i] = sum;
  }
  free(partial);
  /* update_theta: beta.cc:51-82; state index k, two normals per k */
  float eps_t = orc_eps_t(p, step_count);
  for (uint32_t k = 0; k < K; ++k) {
    orc_rng* rseed = &pool[k];
    float r0 = orc_randn(rseed);
    float grads_k = grads_out[2 * k];
    float theta_k = theta[2 * k];
    float f0 = sqrtf(eps_t * theta_k);
    theta_k = fabsf(theta_k + eps_t / 2.0f * (p->eta0 - theta_k + scale * grads_k) + f0 * r0);
    theta[2 * k] = fmaxf(theta_k, 1e-24f);
    float r1 = orc_randn(rseed);
    float grads_2k = grads_out[2 * k + 1];
    float theta_2k = theta[2 * k + 1];
    float f1 = sqrtf(eps_t * theta_2k);
    theta_2k = fabsf(theta_2k + eps_t / 2.0f * (p->eta1 - theta_2k + scale * grads_2k) + f1 * r1);
    theta[2 * k + 1] = fmaxf(theta_2k, 1e-24f);
  }
  orc_theta_to_beta(K, theta, beta);
}

/* ---------------------------------------------------------- perplexity --- */

double orc_perplexity(int mode, uint32_t wg, const orc_params* p, const float* pi_all,
                      const float* beta, const orc_set* heldout, const uint64_t* edges,
                      uint32_t H, float* ppx_per_edge, uint32_t call_count, double* sums_out) {
  const uint32_t K = p->K;
  float* lik = (float*)malloc(sizeof(float) * H);
  uint8_t* is_link = (uint8_t*)malloc(H);
#pragma omp parallel
  {
    float* scratch = (float*)malloc(sizeof(float) * K);
#pragma omp for schedule(static)
    for (uint32_t i = 0; i < H; ++i) {
      uint64_t e = edges[i];
      uint32_t u = (uint32_t)(e >> 32), v = (uint32_t)(e & 0xffffffffu);
      int is_edge = orc_set_has(heldout, e); /* key used as stored: perplexity.cc:45-47 */
      const float* pi_a = pi_all + (uint64_t)u * K;
      const float* pi_b = pi_all + (uint64_t)v * K;
      float s = 0;
      if (mode == ORC_MODE_THREAD) { /* perplexity.cc:16-40 */
        if (is_edge) {
          for (uint32_t k = 0; k < K; ++k) s += pi_a[k] * pi_b[k] * beta[2 * k + 1];
        } else {
          float sum = 0;
          for (uint32_t k = 0; k < K; ++k) {
            float f = pi_a[k] * pi_b[k];
            s += f * (1.0f - beta[2 * k + 1]);
            sum += f;
          }
          s += (1.0f - sum) * (1.0f - p->epsilon);
        }
      } else { /* perplexity.cc:93-126 */
        if (is_edge) {
          for (uint32_t k = 0; k < K; ++k) scratch[k] = pi_a[k] * pi_b[k] * beta[2 * k + 1];
          s = orc_wg_sum_f32(scratch, K, wg);
        } else {
          for (uint32_t k = 0; k < K; ++k) scratch[k] = pi_a[k] * pi_b[k];
          float sum = orc_wg_sum_f32(scratch, K, wg);
          for (uint32_t k = 0; k < K; ++k)
            scratch[k] = pi_a[k] * pi_b[k] * (1.0f - beta[2 * k + 1]);
          s = orc_wg_sum_f32(scratch, K, wg);
          s += (1.0f - sum) * (1.0f - p->epsilon);
        }
      }
      if (s < 1.0e-30f) s = 1.0e-30f;
      /* perplexity.cc:51-64 */
      float ppx = ppx_per_edge[i];
      ppx = (ppx * (call_count - 1) + s) / call_count;
      lik[i] = logf(ppx);
      is_link[i] = (uint8_t)is_edge;
      ppx_per_edge[i] = ppx;
    }
    free(scratch);
  }
  /* perplexity.cu:27-37: four reductions (float sums, uint counts) */
  float link_lik = 0, non_link_lik = 0;
  uint32_t link_count = 0, non_link_count = 0;
  for (uint32_t i = 0; i < H; ++i) {
    if (is_link[i]) { link_lik += lik[i]; ++link_count; }
    else { non_link_lik += lik[i]; ++non_link_count; }
  }
  free(lik); free(is_link);
  if (sums_out) {
    sums_out[0] = link_lik; sums_out[1] = non_link_lik;
    sums_out[2] = link_count; sums_out[3] = non_link_count;
  }
  /* perplexity.cc:264-273 */
  double avg = 0.0;
  if (link_count + non_link_count != 0)
    avg = (link_lik + non_link_lik) / (link_count + non_link_count);
  return -avg;
}

/* ------------------------------------------------------------- pi init --- */

/* random.cc:108-167: pool N*32 seeded {11,113}; G = min(N,65535) groups of 32;
 * group g draws rows g, g+G, ...; lane-strided columns; then
 * normalize.cc:34-52 with wg 32: row /= WG_SUM(row), phi[row] = sum */
void orc_init_pi(uint64_t N, uint32_t K, float eta0, float eta1, float* pi, float* phi) {
  const uint32_t local = 32;
  uint32_t G = N < MAX_GROUPS ? (uint32_t)N : MAX_GROUPS;
#pragma omp parallel for schedule(static)
  for (uint32_t g = 0; g < G; ++g) {
    orc_rng st[32];
    for (uint32_t l = 0; l < local; ++l) {
      uint64_t id = (uint64_t)g * local + l;
      st[l].x = 11 + id;
      st[l].y = 113 + id;
    }
    for (uint64_t row = g; row < N; row += G) {
      float* r = pi + row * K;
      for (uint32_t l = 0; l < local; ++l)
        for (uint32_t j = l; j < K; j += local) r[j] = orc_rand_gamma(&st[l], eta0, eta1);
      phi[row] = orc_wg_normalize_f32(r, K, local);
    }
  }
}
