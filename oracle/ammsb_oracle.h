/*
 * oracle/ammsb_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in plain C, of the reference's SG-MCMC a-MMSB hot path
 * (ielhelw/mcmc-ammsb-gpu).  Every function cites the reference file:line it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit (integer paths)
 * and bit-for-bit / <=1 ulp (fp32 paths, same compiler, -ffp-contract=off) against
 *   (1) the reference's own kernel text and host sources compiled from
 *       /root/reference by oracle/build_ref.py into oracle/_ref/ (tests/test_oracle_vs_ref.py),
 *   (2) golden vectors generated from (1), committed under tests/golden/,
 *   (3) the reference's own exact test properties (cuckoo-test.cc, random-test.cc,
 *       wg-sample-test.cc, wg-sum-test.cc, wg-normalize-test.cc).
 */
#ifndef AMMSB_ORACLE_H_
#define AMMSB_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mcmc/random.h:13 -- ulong2 state, .x = values[0], .y = values[1] */
typedef struct { uint64_t x, y; } orc_rng;

/* Hyper-parameters as they reach the reference kernels (config.cc:66-83).  Float
 * members must already be rounded through "%e" text (orc_round_param). */
typedef struct {
  uint64_t N;
  uint64_t E;
  uint32_t K;
  uint32_t num_neighbors;
  float alpha, a, b, c, epsilon, eta0, eta1;
} orc_params;

/* work-item mapping of a reference launch */
enum { ORC_MODE_THREAD = 0, ORC_MODE_WG = 1 };

/* cuckoo::Set (cuckoo.h:16-67); table layout == Set::Serialize() (cuckoo.cc:211-220) */
typedef struct {
  uint64_t* table;    /* [2][num_bins][4] */
  uint64_t num_bins;  /* N_ */
  uint32_t prime_idx;
  uint64_t count;
} orc_set;

float orc_round_param(float f);                       /* config.cc:57-64 */
float orc_eps_t(const orc_params* p, uint32_t step);  /* learner.cc:41-43 */

/* RNG: random.cc:31-44, random.cl.inc:13-49,221-279,353-395 */
void orc_rng_init(orc_rng* pool, uint64_t n, uint64_t sx, uint64_t sy);
uint64_t orc_rand(orc_rng* s);
float orc_random(orc_rng* s);
int orc_randint(orc_rng* s, int from, int upto);
float orc_randn(orc_rng* s);
float orc_rand_gamma(orc_rng* s, float a, float b);

/* cuckoo: cuckoo.cc:98-220 (host build) and :39-65 (device lookup) */
uint64_t orc_set_bins_for(uint64_t n);
int orc_set_build(const uint64_t* keys, uint64_t n, orc_set* out);
void orc_set_free(orc_set* s);
int orc_set_has(const orc_set* s, uint64_t key);
void orc_set_has_many(const orc_set* s, const uint64_t* keys, uint64_t n, uint8_t* out);

/* neighbor sampler: sample.cc:15-77 kernel, :111-121 launch geometry */
void orc_neighbor_sample(orc_rng* pool, const uint32_t* nodes, uint32_t V,
                         uint32_t N, uint32_t n, uint32_t wg,
                         uint32_t* hash_scratch /* [V*2n] */,
                         uint32_t* out /* [V*n] */);

/* update_phi: phi.cc:78-152 (THREAD), :214-302 (WG-NAIVE), launch :728-757 */
void orc_update_phi(int mode, uint32_t wg, const orc_params* p,
                    const float* beta /* [2K] */, const float* pi /* [N,K] */,
                    const float* phi /* [N] */, const orc_set* train,
                    const uint32_t* nodes, const uint32_t* neighbors /* [V,n] */,
                    uint32_t V, uint32_t step_count, orc_rng* pool,
                    int disable_noise, float* phi_vec /* [V,K] */);

/* update_pi: phi.cc:154-173 (THREAD), :178-197 (WG) */
void orc_update_pi(int mode, uint32_t wg, uint32_t K, float* pi, float* phi,
                   const float* phi_vec, const uint32_t* nodes, uint32_t V);

/* BetaUpdater::operator(): beta.cc:334-384; kernels :30-82, :87-137, :174-234 */
void orc_update_beta(int mode, uint32_t wg, const orc_params* p,
                     float* theta /* [2K] in/out */, float* beta /* [2K] out */,
                     const float* pi, const orc_set* train, const uint64_t* edges,
                     uint32_t E_mb, float scale, uint32_t step_count,
                     orc_rng* pool /* [K] */, float* theta_sum /* [K] out */,
                     float* grads /* [2K] out (summed) */);

/* PerplexityCalculator::operator(): perplexity.cc:14-182, :251-274.
 * sums_out = {link_lik, non_link_lik, link_count, non_link_count}; the four
 * sums are accumulated serially in pair order (library reduce order is
 * unpinned in the reference).  Returns avg (not exp'd). */
double orc_perplexity(int mode, uint32_t wg, const orc_params* p, const float* pi,
                      const float* beta, const orc_set* heldout,
                      const uint64_t* edges, uint32_t H, float* ppx_per_edge,
                      uint32_t call_count, double* sums_out /* [4] or NULL */);

/* pi init: random.cc:108-167 + normalize.cc:34-52 (pool N*32, seed {11,113}) */
void orc_init_pi(uint64_t N, uint32_t K, float eta0, float eta1, float* pi, float* phi);

/* beta = row-normalised theta: random.h:70-79 + normalize.cc:13-32 (slice 2, wg 1) */
void orc_theta_to_beta(uint32_t K, const float* theta, float* beta);

/* WG helpers exposed for the reference's own unit-test properties
 * (wg-sum-test.cc, wg-normalize-test.cc): sum.cc:11-42, normalize.cc:13-23 */
float orc_wg_sum_f32(const float* in, uint32_t len, uint32_t wg);
uint32_t orc_wg_sum_u32(const uint32_t* in, uint32_t len, uint32_t wg);
float orc_wg_normalize_f32(float* inout, uint32_t len, uint32_t wg);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* AMMSB_ORACLE_H_ */
