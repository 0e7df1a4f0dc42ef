// oracle/ref_api.cc -- TEST INFRASTRUCTURE ONLY.
//
// Exposes the reference's OWN code through the orc_* C API of ammsb_oracle.h:
//   - device kernels: the raw-string kernel sources, extracted verbatim by
//     build_ref.py into oracle/_ref/gen/*.inc and compiled here under ref_prelude.h,
//     once with WG_SIZE 1 (the THREAD / EDGE_PER_THREAD variants the reference selects
//     on CPU devices, learner.cc:105-107,113-114) and once with WG_SIZE 32 (WG-NAIVE,
//     the reference's default phi mode, main.cc:71);
//   - host code: mcmc::cuckoo::Set, GenerateSetsFromEdges, Graph, the mini-batch
//     strategies and MakeCompileFlags from the unmodified cuckoo.cc/data.cc/sample.cc/
//     config.cc objects.
// Only launch geometry (a few lines of host code per operator that live in .cc files
// which cannot be compiled without the OpenCL/CUDA backend) is restated here, each with
// its citation.
#include <omp.h>

#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_set>

#include "mcmc/config.h"
#include "mcmc/cuckoo.h"
#include "mcmc/data.h"
#include "mcmc/sample.h"

#include "ammsb_oracle.h"
#include "ref_prelude.h"

// ------------------------------------------------------------ work-item runtime --
namespace refrt {

thread_local WorkItem wi;
KernelConfig kc;

struct Fibers {
  std::vector<ucontext_t> ctx;
  std::vector<std::vector<char>> stacks;
  std::vector<char> done;
  ucontext_t main;
  size_t cur = 0;
  const std::function<void()>* fn = nullptr;
};
static thread_local Fibers* fibers = nullptr;

static void trampoline() {
  Fibers* f = fibers;
  (*f->fn)();
  f->done[f->cur] = 1;
  swapcontext(&f->ctx[f->cur], &f->main);
}

void wg_barrier() {
  Fibers* f = fibers;
  if (f) swapcontext(&f->ctx[f->cur], &f->main);
}

static void run_group_fibers(size_t grp, size_t ngrp, size_t lsize, const std::function<void()>& fn) {
  static thread_local Fibers store;
  Fibers* f = &store;
  if (f->ctx.size() < lsize) {
    f->ctx.resize(lsize);
    f->stacks.resize(lsize);
    for (auto& s : f->stacks)
      if (s.empty()) s.resize(1 << 20);
    f->done.resize(lsize);
  }
  f->fn = &fn;
  fibers = f;
  for (size_t l = 0; l < lsize; ++l) {
    getcontext(&f->ctx[l]);
    f->ctx[l].uc_stack.ss_sp = f->stacks[l].data();
    f->ctx[l].uc_stack.ss_size = f->stacks[l].size();
    f->ctx[l].uc_link = &f->main;
    makecontext(&f->ctx[l], trampoline, 0);
    f->done[l] = 0;
  }
  size_t remaining = lsize;
  while (remaining) {
    for (size_t l = 0; l < lsize; ++l) {
      if (f->done[l]) continue;
      f->cur = l;
      wi = WorkItem{grp * lsize + l, ngrp * lsize, l, lsize, grp, ngrp};
      swapcontext(&f->main, &f->ctx[l]);
      if (f->done[l]) --remaining;
    }
  }
  fibers = nullptr;
}

void launch(size_t ngroups, size_t lsize, bool with_barriers, const std::function<void()>& fn) {
#pragma omp parallel for schedule(static)
  for (size_t g = 0; g < ngroups; ++g) {
    if (with_barriers && lsize > 1) {
      run_group_fibers(g, ngroups, lsize, fn);
    } else {
      for (size_t l = 0; l < lsize; ++l) {
        wi = WorkItem{g * lsize + l, ngroups * lsize, l, lsize, g, ngroups};
        fn();
      }
    }
  }
}

}  // namespace refrt

// ------------------------------------------------------- the reference's kernels --
typedef unsigned int uint;
typedef unsigned long ulong;

#define REF_COMMON_KERNELS                     \
  struct ulong2 {                              \
    ulong x, y;                                \
  };                                           \
  typedef ulong uint64_t;                      \
  typedef uint uint32_t;                       \
  using std::max;                              \
  using std::min;

namespace ref_thread {
REF_COMMON_KERNELS
#define WG_SIZE 1
#include "random_types.inc"
#include "random_impl.inc"
#include "random_source.inc"
#include "set_types.inc"
#include "set_header.inc"
#include "set_source.inc"
#include "rpm.inc"
#include "base_funcs.inc"
#include "sum.inc"
#include "normalize.inc"
#include "sampler.inc"
#include "gamma.inc"
#include "phi_vec.inc"
#include "phi_thread.inc"
#include "beta_base.inc"
#include "beta_thread.inc"
#include "ppx_thread.inc"
#undef WG_SIZE
}  // namespace ref_thread

namespace ref_wg {
REF_COMMON_KERNELS
#define WG_SIZE 32
#include "random_types.inc"
#include "random_impl.inc"
#include "random_source.inc"
#include "set_types.inc"
#include "set_header.inc"
#include "set_source.inc"
#include "rpm.inc"
#include "base_funcs.inc"
#include "sum.inc"
#include "normalize.inc"
#include "gamma.inc"
#include "phi_vec.inc"
#include "phi_wg.inc"
#include "pi_wg.inc"
#include "ppx_wg.inc"
#undef WG_SIZE
}  // namespace ref_wg

// the single-letter configuration macros must not leak into the API code below
#undef K
#undef N
#undef E
#undef NUM_NEIGHBORS
#undef ALPHA
#undef EPSILON
#undef EPS_A
#undef EPS_B
#undef EPS_C
#undef ETA0
#undef ETA1
#undef Float
#undef MAX
#undef LOG
#undef EXP
#undef POW
#undef SQRT
#undef FABS

static const uint32_t kMaxGroups = 65535;  // GetMaxGroups(), types.cc:537
using refrt::kc;
using refrt::launch;

static void set_cfg(const orc_params* p, int disable_noise = 0) {
  kc.K = (int)p->K;
  kc.N = (int)p->N;
  kc.E = (int)p->E;
  kc.NUM_NEIGHBORS = (int)p->num_neighbors;
  kc.ALPHA = p->alpha;
  kc.EPS_A = p->a;
  kc.EPS_B = p->b;
  kc.EPS_C = p->c;
  kc.EPSILON = p->epsilon;
  kc.ETA0 = p->eta0;
  kc.ETA1 = p->eta1;
  kc.disable_noise = disable_noise;
}

// device-side handle structs, filled by the reference's own init kernels
#define DEFINE_HANDLES(NS)                                                                       \
  struct Handles_##NS {                                                                          \
    NS::floatRowPartitionedMatrix pm;                                                            \
    NS::Random rnd;                                                                              \
    NS::Set set;                                                                                 \
    void make_pm(const float* pi, uint32_t rows, uint32_t cols) {                                \
      refrt::wi = refrt::WorkItem{0, 1, 0, 1, 0, 1};                                             \
      /* one block (partitioned-alloc.h:92-118) */                                               \
      NS::floatRowPartitionedMatrix_init(&pm, rows, 1, rows, cols);                              \
      NS::floatRowPartitionedMatrix_set(&pm, (void*)pi, 0);                                      \
    }                                                                                            \
    void make_rnd(orc_rng* pool, uint64_t n) {                                                   \
      rnd.base_ = reinterpret_cast<NS::random_seed_t*>(pool);                                    \
      rnd.num_seeds = n;                                                                         \
    }                                                                                            \
    void make_set(const orc_set* s) {                                                            \
      refrt::wi = refrt::WorkItem{0, 1, 0, 1, 0, 1};                                             \
      NS::SetInit(&set, reinterpret_cast<NS::uint64_t*>(s->table), s->num_bins, s->prime_idx);   \
    }                                                                                            \
  };
DEFINE_HANDLES(ref_thread)
DEFINE_HANDLES(ref_wg)

extern "C" {

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// MakeCompileFlags / float_to_string (config.cc:57-83) from the reference's own object
float orc_round_param(float f) {
  mcmc::Config cfg;
  cfg.alpha = f;
  for (const std::string& s : mcmc::MakeCompileFlags(cfg)) {
    if (s.rfind("-DALPHA=", 0) == 0) return strtof(s.c_str() + 8, nullptr);
  }
  abort();
}

float orc_eps_t(const orc_params* p, uint32_t step) {
  set_cfg(p);
  return ref_thread::get_eps_t(step);
}

void orc_rng_init(orc_rng* pool, uint64_t n, uint64_t sx, uint64_t sy) {
  // OpenClRandom::SetSeed launches RandomInit with {1},{1} (random.cc:59-69)
  ref_thread::Random r;
  ref_thread::ulong2 seed{sx, sy};
  refrt::wi = refrt::WorkItem{0, 1, 0, 1, 0, 1};
  ref_thread::RandomInit(&r, (int)n, seed, reinterpret_cast<ref_thread::ulong2*>(pool));
}
uint64_t orc_rand(orc_rng* s) { return ref_thread::rand(reinterpret_cast<ref_thread::ulong2*>(s)); }
float orc_random(orc_rng* s) { return ref_thread::random(reinterpret_cast<ref_thread::ulong2*>(s)); }
int orc_randint(orc_rng* s, int from, int upto) {
  return ref_thread::randint(reinterpret_cast<ref_thread::ulong2*>(s), from, upto);
}
float orc_randn(orc_rng* s) { return ref_thread::randn(reinterpret_cast<ref_thread::ulong2*>(s)); }
float orc_rand_gamma(orc_rng* s, float a, float b) {
  return ref_thread::rand_gamma(reinterpret_cast<ref_thread::ulong2*>(s), a, b);
}

// ---- cuckoo: the reference's host class builds, its device function looks up ----
uint64_t orc_set_bins_for(uint64_t n) { return mcmc::cuckoo::Set(n).BinsPerBucket(); }

int orc_set_build(const uint64_t* keys, uint64_t n, orc_set* out) {
  std::vector<mcmc::Edge> v(keys, keys + n);
  mcmc::cuckoo::Set set(n);
  bool ok = set.SetContents(v.begin(), v.end());
  std::vector<mcmc::Edge> ser = set.Serialize();
  out->num_bins = set.BinsPerBucket();
  out->prime_idx = set.PrimeIdx();
  out->count = set.Size();
  out->table = (uint64_t*)malloc(sizeof(uint64_t) * ser.size());
  memcpy(out->table, ser.data(), sizeof(uint64_t) * ser.size());
  // host Set::Has must agree with the device lookup on every inserted key
  if (ok)
    for (uint64_t i = 0; i < n; ++i)
      if (!set.Has(keys[i])) abort();
  return ok ? 1 : 0;
}
void orc_set_free(orc_set* s) {
  free(s->table);
  s->table = nullptr;
}
int orc_set_has(const orc_set* s, uint64_t key) {
  Handles_ref_thread h;
  h.make_set(s);
  return ref_thread::Set_HasEdge(&h.set, key) ? 1 : 0;
}
void orc_set_has_many(const orc_set* s, const uint64_t* keys, uint64_t n, uint8_t* out) {
  Handles_ref_thread h;
  h.make_set(s);
  for (uint64_t i = 0; i < n; ++i) out[i] = ref_thread::Set_HasEdge(&h.set, keys[i]) ? 1 : 0;
}

// ---- neighbor sampler (launch: sample.cc:111-121) ----
void orc_neighbor_sample(orc_rng* pool, const uint32_t* nodes, uint32_t V, uint32_t N_, uint32_t n,
                         uint32_t wg, uint32_t* hash, uint32_t* out) {
  Handles_ref_thread h;
  h.make_rnd(pool, 0);
  uint32_t global = std::min(V / wg + (V % wg ? 1 : 0), kMaxGroups / wg);
  uint32_t capacity = 2 * n;
  launch(global, wg, false, [&]() {
    ref_thread::generate_random_int_kernel(V, (uint*)nodes, hash, N_, n, capacity, out, &h.rnd);
  });
}

// ---- update_phi / update_pi (launch: phi.cc:728-763) ----
void orc_update_phi(int mode, uint32_t wg, const orc_params* p, const float* beta, const float* pi,
                    const float* phi, const orc_set* train, const uint32_t* nodes,
                    const uint32_t* neighbors, uint32_t V, uint32_t step_count, orc_rng* pool,
                    int disable_noise, float* phi_vec) {
  set_cfg(p, disable_noise);
  orc_rng dummy{0, 0};
  if (mode == ORC_MODE_THREAD) {
    Handles_ref_thread h;
    h.make_pm(pi, (uint32_t)p->N, p->K);
    h.make_set(train);
    uint32_t groups = std::min(V / wg + (V % wg ? 1 : 0), kMaxGroups);
    uint64_t global = (uint64_t)groups * wg;
    std::vector<orc_rng> tmp;
    if (!pool) { tmp.assign(global, dummy); pool = tmp.data(); }
    h.make_rnd(pool, global);
    std::vector<float> grads(global * p->K), probs(global * p->K);  // phi.cc:639-646
    launch(groups, wg, false, [&]() {
      ref_thread::update_phi((float*)beta, &h.pm, (float*)phi, phi_vec, &h.set, (uint*)nodes,
                             (uint*)neighbors, V, step_count, grads.data(), probs.data(), &h.rnd);
    });
  } else {
    if (wg != 32) abort();  // the WG text is compiled for WG_SIZE 32 only
    Handles_ref_wg h;
    h.make_pm(pi, (uint32_t)p->N, p->K);
    h.make_set(train);
    uint32_t groups = std::min(V, kMaxGroups);
    std::vector<orc_rng> tmp;
    if (!pool) { tmp.assign((uint64_t)groups * wg, dummy); pool = tmp.data(); }
    h.make_rnd(pool, (uint64_t)groups * wg);
    launch(groups, wg, true, [&]() {
      ref_wg::update_phi((float*)beta, &h.pm, (float*)phi, phi_vec, &h.set, (uint*)nodes,
                         (uint*)neighbors, V, step_count, &h.rnd);
    });
  }
}

void orc_update_pi(int mode, uint32_t wg, uint32_t K_, float* pi, float* phi, const float* phi_vec,
                   const uint32_t* nodes, uint32_t V) {
  kc.K = (int)K_;
  uint32_t rows = 0;
  for (uint32_t i = 0; i < V; ++i) rows = std::max(rows, nodes[i] + 1);
  if (mode == ORC_MODE_THREAD) {
    Handles_ref_thread h;
    h.make_pm(pi, rows, K_);
    uint32_t groups = std::min(V / wg + (V % wg ? 1 : 0), kMaxGroups);
    launch(groups, wg, false,
           [&]() { ref_thread::update_pi(&h.pm, (float*)phi_vec, phi, (uint*)nodes, V); });
  } else {
    if (wg != 32) abort();
    Handles_ref_wg h;
    h.make_pm(pi, rows, K_);
    uint32_t groups = std::min(V, kMaxGroups);
    launch(groups, wg, true,
           [&]() { ref_wg::update_pi(&h.pm, (float*)phi_vec, phi, (uint*)nodes, V); });
  }
}

// ---- BetaUpdater::operator() (beta.cc:334-384), EDGE_PER_THREAD kernels ----
void orc_theta_to_beta(uint32_t K_, const float* theta, float* beta) {
  memcpy(beta, theta, sizeof(float) * 2 * K_);  // theta_.CopyTo(beta_), beta.cc:378
  // Normalizer(slice=2, wg=1) (beta.cc:249, normalize.h:39-46)
  uint32_t groups = std::min(K_, kMaxGroups);
  launch(groups, 1, false, [&]() { ref_thread::WG_NORMALIZE_KERNEL_float(beta, K_, 2); });
}

void orc_update_beta(int mode, uint32_t wg, const orc_params* p, float* theta, float* beta,
                     const float* pi, const orc_set* train, const uint64_t* edges, uint32_t E_mb,
                     float scale, uint32_t step_count, orc_rng* pool, float* theta_sum,
                     float* grads_out) {
  if (mode != ORC_MODE_THREAD) abort();  // lgrads[2*K] needs a compile-time K (beta.cc:183)
  set_cfg(p);
  const uint32_t Kk = p->K;
  Handles_ref_thread h;
  h.make_pm(pi, (uint32_t)p->N, Kk);
  h.make_set(train);
  h.make_rnd(pool, Kk);
  const uint32_t w32 = 32;
  const uint32_t kwg = (Kk / w32 + (Kk % w32 ? 1 : 0)) * w32;
  launch(kwg / w32, w32, false, [&]() { ref_thread::sum_theta(theta, theta_sum); });
  uint32_t global = std::min(E_mb / wg + (E_mb % wg ? 1 : 0), kMaxGroups) * wg;
  uint32_t num_partials = std::min(global, E_mb);
  std::vector<float> probs((size_t)std::max(E_mb, 1u) * Kk), grads((size_t)std::max(num_partials, 1u) * 2 * Kk);
  launch(global / wg, wg, false, [&]() {
    ref_thread::calculate_grads_partial(theta, theta_sum, beta, &h.pm, &h.set, (ulong*)edges, E_mb,
                                        probs.data(), grads.data());
  });
  launch(2 * kwg / w32, w32, false, [&]() { ref_thread::sum_grads(grads.data(), num_partials); });
  launch(kwg / w32, w32, false,
         [&]() { ref_thread::update_theta(theta, grads.data(), step_count, scale, &h.rnd); });
  memcpy(grads_out, grads.data(), sizeof(float) * 2 * Kk);
  orc_theta_to_beta(Kk, theta, beta);
}

// ---- PerplexityCalculator::operator() (perplexity.cc:184-274) ----
double orc_perplexity(int mode, uint32_t wg, const orc_params* p, const float* pi, const float* beta,
                      const orc_set* heldout, const uint64_t* edges, uint32_t H, float* ppx_per_edge,
                      uint32_t call_count, double* sums_out) {
  set_cfg(p);
  std::vector<float> ll(H), nl(H);
  std::vector<uint32_t> lc(H), nc(H);
  if (mode == ORC_MODE_THREAD) {
    Handles_ref_thread h;
    h.make_pm(pi, (uint32_t)p->N, p->K);
    h.make_set(heldout);
    uint32_t groups = std::min(H / wg + (H % wg ? 1 : 0), kMaxGroups);
    launch(groups, wg, false, [&]() {
      ref_thread::calculate_ppx_partial_for_edge((ulong*)edges, H, &h.pm, (float*)beta, &h.set,
                                                 ppx_per_edge, ll.data(), nl.data(), lc.data(),
                                                 nc.data(), call_count);
    });
  } else {
    if (wg != 32) abort();
    Handles_ref_wg h;
    h.make_pm(pi, (uint32_t)p->N, p->K);
    h.make_set(heldout);
    uint32_t groups = std::min(H, kMaxGroups);
    std::vector<float> scratch((size_t)groups * p->K);
    launch(groups, wg, true, [&]() {
      ref_wg::calculate_ppx_partial_for_edge((ulong*)edges, H, &h.pm, (float*)beta, &h.set,
                                             ppx_per_edge, ll.data(), nl.data(), lc.data(), nc.data(),
                                             call_count, scratch.data());
    });
  }
  // perplexity.cu:27-37: four library reductions; order unpinned -> serial here
  float link_lik = 0, non_link_lik = 0;
  uint32_t link_count = 0, non_link_count = 0;
  for (uint32_t i = 0; i < H; ++i) {
    link_lik += ll[i];
    non_link_lik += nl[i];
    link_count += lc[i];
    non_link_count += nc[i];
  }
  if (sums_out) {
    sums_out[0] = link_lik;
    sums_out[1] = non_link_lik;
    sums_out[2] = link_count;
    sums_out[3] = non_link_count;
  }
  double avg = 0.0;  // perplexity.cc:264-268
  if (link_count + non_link_count != 0) avg = (link_lik + non_link_lik) / (link_count + non_link_count);
  return -avg;
}

// ---- RandomGammaAndNormalize (random.cc:131-167) ----
void orc_init_pi(uint64_t N_, uint32_t K_, float eta0, float eta1, float* pi, float* phi) {
  const uint32_t local = 32;
  std::vector<orc_rng> pool(N_ * 32);
  orc_rng_init(pool.data(), pool.size(), 11, 113);
  Handles_ref_wg h;
  h.make_pm(pi, (uint32_t)N_, K_);
  h.make_rnd(pool.data(), pool.size());
  uint32_t groups = std::min((uint32_t)N_, kMaxGroups);
  launch(groups, local, false, [&]() { ref_wg::generate_gamma(&h.pm, &h.rnd, eta0, eta1); });
  launch(groups, local, true, [&]() { ref_wg::WG_NORMALIZE_PARTITIONED_KERNEL_float(&h.pm, phi); });
}

// ---- work-group helpers (wg-sum-test.cc / wg-normalize-test.cc entry points) ----
float orc_wg_sum_f32(const float* in, uint32_t len, uint32_t wg) {
  if (wg != 32) abort();
  float out = 0;
  launch(1, wg, true, [&]() { ref_wg::WG_SUM_KERNEL_float((float*)in, &out, 1, len); });
  return out;
}
uint32_t orc_wg_sum_u32(const uint32_t*, uint32_t, uint32_t) { abort(); }
float orc_wg_normalize_f32(float* inout, uint32_t len, uint32_t wg) {
  float s = orc_wg_sum_f32(inout, len, wg);
  launch(1, wg, true, [&]() { ref_wg::WG_NORMALIZE_KERNEL_float(inout, 1, len); });
  return s;
}

// ---- host logic of the reference, for golden vectors (tests/golden/) ----

// GenerateSetsFromEdges (data.cc:80-128) after srand(seed); returns counts, fills arrays
int ref_generate_sets(uint64_t N_, const uint64_t* edges, uint64_t n, double heldout_ratio,
                      unsigned srand_seed, uint64_t* training_out, uint64_t* n_training,
                      uint64_t* heldout_out, uint64_t* n_heldout) {
  std::vector<mcmc::Edge> vals(edges, edges + n), tr, he;
  std::unique_ptr<mcmc::Set> ts, hs;
  srand(srand_seed);
  if (!mcmc::GenerateSetsFromEdges(N_, vals, heldout_ratio, &tr, &he, &ts, &hs)) return 0;
  memcpy(training_out, tr.data(), 8 * tr.size());
  memcpy(heldout_out, he.data(), 8 * he.size());
  *n_training = tr.size();
  *n_heldout = he.size();
  return 1;
}

// GetUniqueEdgesFromFile (data.cc:36-78) after srand(seed): SNAP text -> renumbered, sorted,
// de-duplicated, shuffled edge list.  Returns the edge count (-1: failure, -2: edges_out too small).
int64_t ref_unique_edges_from_file(const char* path, unsigned srand_seed, uint64_t* count_vertices,
                                   uint64_t* edges_out, uint64_t cap) {
  std::vector<mcmc::Edge> vals;
  srand(srand_seed);
  if (!mcmc::GetUniqueEdgesFromFile(path, count_vertices, &vals)) return -1;
  if (vals.size() > cap) return -2;
  memcpy(edges_out, vals.data(), 8 * vals.size());
  return (int64_t)vals.size();
}

// one mini-batch of the given strategy + ExtractNodesFromMiniBatch (learner.cc:162-173)
struct RefSamplerCtx {
  mcmc::Config cfg;
};

void* ref_sampler_create(uint64_t N_, uint64_t E_, const uint64_t* training, uint64_t n_training,
                         const uint64_t* heldout_links, uint64_t n_heldout, uint64_t mini_batch) {
  RefSamplerCtx* c = new RefSamplerCtx();
  c->cfg.N = N_;
  c->cfg.E = E_;
  c->cfg.mini_batch_size = mini_batch;
  c->cfg.training_edges.assign(training, training + n_training);
  c->cfg.training.reset(new mcmc::Set(n_training));
  if (!c->cfg.training->SetContents(c->cfg.training_edges.begin(), c->cfg.training_edges.end())) abort();
  std::vector<mcmc::Edge> he(heldout_links, heldout_links + n_heldout);
  c->cfg.heldout.reset(new mcmc::Set(n_heldout));
  if (!c->cfg.heldout->SetContents(he.begin(), he.end())) abort();
  c->cfg.trainingGraph.reset(new mcmc::Graph(N_, c->cfg.training_edges));
  return c;
}
void ref_sampler_destroy(void* h) { delete (RefSamplerCtx*)h; }

// operator>>(SampleStrategy) (sample.cc:135-156) / operator>>(PhiUpdaterMode) (config.cc:119-137):
// the enum value for a token, -1 when the reference throws
int ref_parse_token(int kind, const char* token) {
  std::istringstream in(token);
  try {
    if (kind == 0) {
      mcmc::SampleStrategy s;
      in >> s;
      return (int)s;
    }
    mcmc::PhiUpdaterMode m;
    in >> m;
    return (int)m;
  } catch (...) {
    return -1;
  }
}

// operator<<(Config) (config.cc:85-117) followed by the MakeCompileFlags list (config.cc:66-83), one
// flag per line after a "flags:" line.  v[] = heldout_ratio, alpha, a, b, c, epsilon, eta0, eta1, K, m, n,
// N, E, ppx_wg, phi_wg, beta_wg, strategy, phi_mode, phi_vector_width, probs/grads/pi shared;
// seeds = phi, beta, neighbor (x, y each).  h (may be NULL): a sampler context whose sets are printed.
int ref_config_print(void* h, const double* v, const uint64_t* seeds, char* buf, uint64_t len) {
  mcmc::Config local;
  mcmc::Config& c = h ? ((RefSamplerCtx*)h)->cfg : local;
  c.heldout_ratio = v[0]; c.alpha = v[1]; c.a = v[2]; c.b = v[3]; c.c = v[4]; c.epsilon = v[5];
  c.eta0 = v[6]; c.eta1 = v[7];
  c.K = (uint64_t)v[8]; c.mini_batch_size = (uint64_t)v[9]; c.num_node_sample = (uint64_t)v[10];
  c.N = (uint64_t)v[11]; c.E = (uint64_t)v[12];
  c.ppx_wg_size = (uint32_t)v[13]; c.phi_wg_size = (uint32_t)v[14]; c.beta_wg_size = (uint32_t)v[15];
  c.strategy = (mcmc::SampleStrategy)(int)v[16];
  c.phi_mode = (mcmc::PhiUpdaterMode)(int)v[17];
  c.phi_vector_width = (uint32_t)v[18];
  c.phi_probs_shared = v[19] != 0; c.phi_grads_shared = v[20] != 0; c.phi_pi_shared = v[21] != 0;
  c.phi_seed = {seeds[0], seeds[1]};
  c.beta_seed = {seeds[2], seeds[3]};
  c.neighbor_seed = {seeds[4], seeds[5]};
  std::ostringstream o;
  o << c << "flags:\n";
  for (const std::string& f : mcmc::MakeCompileFlags(c)) o << f << "\n";
  const std::string s = o.str();
  if (s.size() + 1 > len) return 1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}
uint64_t ref_sampler_max_fan_out(void* h) { return ((RefSamplerCtx*)h)->cfg.trainingGraph->MaxFanOut(); }

// strategy numbering = enum SampleStrategy (sample.h:94-101)
float ref_sample(void* h, int strategy, unsigned* seed, uint64_t* edges_out, uint64_t* n_edges,
                 uint32_t* nodes_out, uint64_t* n_nodes) {
  RefSamplerCtx* c = (RefSamplerCtx*)h;
  std::vector<mcmc::Edge> edges;
  float w = 0;
  switch (strategy) {
    case mcmc::Node: w = mcmc::sampleNode(c->cfg, &edges, seed); break;
    case mcmc::NodeLink: w = mcmc::sampleNodeLink(c->cfg, &edges, seed); break;
    case mcmc::NodeNonLink: w = mcmc::sampleNodeNonLink(c->cfg, &edges, seed); break;
    case mcmc::BFLink: w = mcmc::sampleBreadthFirstLink(c->cfg, &edges, seed); break;
    case mcmc::BFNonLink: w = mcmc::sampleBreadthFirstNonLink(c->cfg, &edges, seed); break;
    case mcmc::BF: w = mcmc::sampleBreadthFirst(c->cfg, &edges, seed); break;
    default: abort();
  }
  // ExtractNodesFromMiniBatch, learner.cc:162-173 (learner.cc itself needs the device
  // backend to compile; these ten lines are restated with the same container)
  std::unordered_set<mcmc::Vertex> nodes;
  for (auto e : edges) {
    mcmc::Vertex u, v;
    std::tie(u, v) = mcmc::Vertices(e);
    nodes.insert(u);
    nodes.insert(v);
  }
  std::vector<mcmc::Vertex> nodes_vec;
  nodes_vec.insert(nodes_vec.begin(), nodes.begin(), nodes.end());
  memcpy(edges_out, edges.data(), 8 * edges.size());
  *n_edges = edges.size();
  memcpy(nodes_out, nodes_vec.data(), 4 * nodes_vec.size());
  *n_nodes = nodes_vec.size();
  return w;
}

}  // extern "C"
