// capi.cc -- C wrappers over the C++ host API (mcmc::Config / Learner / data / sampling)
// so that the Python test and bench harness can drive it with ctypes.  Not part of the
// drop-in surface: a C++ user links libmcmc.so and includes mcmc/learner.h directly.
#include <cstring>
#include <fstream>
#include <functional>
#include <random>
#include <sstream>

#include "mcmc/learner.h"
#include "mcmc/sharded_learner.h"
#include "mcmc/serialize.h"
#include "mcmc/std_order_set.h"
#include <unordered_set>

using namespace mcmc;

namespace {
thread_local std::string g_err;
template <class F>
int Guard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}
}  // namespace

extern "C" {

const char* mcmc_last_error() { return g_err.c_str(); }

// ---- Config ----
void* mcmc_config_create() { return new Config(); }
void mcmc_config_destroy(void* c) { delete static_cast<Config*>(c); }

int mcmc_config_set(void* vc, const char* key, double v) {
  Config* c = static_cast<Config*>(vc);
  const std::string k(key);
  if (k == "heldout_ratio") c->heldout_ratio = v;
  else if (k == "alpha") c->alpha = v;
  else if (k == "a") c->a = v;
  else if (k == "b") c->b = v;
  else if (k == "c") c->c = v;
  else if (k == "epsilon") c->epsilon = v;
  else if (k == "eta0") c->eta0 = v;
  else if (k == "eta1") c->eta1 = v;
  else if (k == "K") c->K = static_cast<uint64_t>(v);
  else if (k == "mini_batch_size") c->mini_batch_size = static_cast<uint64_t>(v);
  else if (k == "num_node_sample") c->num_node_sample = static_cast<uint64_t>(v);
  else if (k == "ppx_interval") c->ppx_interval = static_cast<uint32_t>(v);
  else if (k == "neighbor_sampler_wg_size") c->neighbor_sampler_wg_size = static_cast<uint32_t>(v);
  else if (k == "phi_wg_size") c->phi_wg_size = static_cast<uint32_t>(v);
  else if (k == "phi_disable_noise") c->phi_disable_noise = v != 0;
  else if (k == "strategy") c->strategy = static_cast<SampleStrategy>(static_cast<int>(v));
  else if (k == "phi_mode") c->phi_mode = static_cast<PhiUpdaterMode>(static_cast<int>(v));
  else if (k == "phi_strict") c->phi_strict = v != 0;
  else if (k == "calc_train_ppx") c->calc_train_ppx = v != 0;
  else if (k == "training_ppx_ratio") c->training_ppx_ratio = v;
  else if (k == "stage_timers") c->stage_timers = v != 0;
  else if (k == "device_sampler") c->device_sampler = v != 0;
  else if (k == "N") c->N = static_cast<uint64_t>(v);
  else if (k == "E") c->E = static_cast<uint64_t>(v);
  else if (k == "ppx_wg_size") c->ppx_wg_size = static_cast<uint32_t>(v);
  else if (k == "beta_wg_size") c->beta_wg_size = static_cast<uint32_t>(v);
  else if (k == "phi_vector_width") c->phi_vector_width = static_cast<uint32_t>(v);
  else if (k == "phi_probs_shared") c->phi_probs_shared = v != 0;
  else if (k == "phi_grads_shared") c->phi_grads_shared = v != 0;
  else if (k == "phi_pi_shared") c->phi_pi_shared = v != 0;
  else { g_err = "unknown config key " + k; return 1; }
  return 0;
}
int mcmc_config_set_seed(void* vc, const char* key, uint64_t x, uint64_t y) {
  Config* c = static_cast<Config*>(vc);
  const std::string k(key);
  if (k == "phi_seed") c->phi_seed = {x, y};
  else if (k == "beta_seed") c->beta_seed = {x, y};
  else if (k == "neighbor_seed") c->neighbor_seed = {x, y};
  else { g_err = "unknown seed " + k; return 1; }
  return 0;
}

// main.cc:101-154 minus the file I/O: split, build sets and graphs, alpha = 1/K if 0
int mcmc_config_set_graph(void* vc, uint64_t N, const uint64_t* edges, uint64_t n, unsigned srand_seed) {
  Config* c = static_cast<Config*>(vc);
  return Guard([&] {
    std::vector<Edge> vals(edges, edges + n);
    c->N = N;
    c->training_edges.clear();
    c->heldout_edges.clear();
    srand(srand_seed);
    if (!GenerateSetsFromEdges(N, vals, c->heldout_ratio, &c->training_edges, &c->heldout_edges, &c->training,
                               &c->heldout))
      throw std::runtime_error("Failed to generate training/heldout sets");
    c->trainingGraph.reset(new Graph(N, c->training_edges));
    c->heldoutGraph.reset(new Graph(N, c->heldout_edges));
    if (c->alpha == 0) c->alpha = static_cast<Float>(1) / c->K;
    c->E = vals.size();
  });
}
uint64_t mcmc_config_num_training(void* vc) { return static_cast<Config*>(vc)->training_edges.size(); }
uint64_t mcmc_config_num_heldout(void* vc) { return static_cast<Config*>(vc)->heldout_edges.size(); }
void mcmc_config_get_edges(void* vc, uint64_t* training, uint64_t* heldout) {
  Config* c = static_cast<Config*>(vc);
  if (training) std::memcpy(training, c->training_edges.data(), 8 * c->training_edges.size());
  if (heldout) std::memcpy(heldout, c->heldout_edges.data(), 8 * c->heldout_edges.size());
}
uint64_t mcmc_config_max_fan_out(void* vc) { return static_cast<Config*>(vc)->trainingGraph->MaxFanOut(); }
uint64_t mcmc_config_max_nodes(void* vc) { return MaxMiniBatchNodes(*static_cast<Config*>(vc)); }
uint64_t mcmc_config_max_edges(void* vc) { return MaxMiniBatchEdges(*static_cast<Config*>(vc)); }
int mcmc_config_print(void* vc, char* buf, size_t len) {
  std::ostringstream o;
  o << *static_cast<Config*>(vc);
  std::strncpy(buf, o.str().c_str(), len - 1);
  buf[len - 1] = 0;
  return 0;
}
// operator>>(SampleStrategy) / operator>>(PhiUpdaterMode): the enum value, -1 when the parser throws
int mcmc_parse_token(int kind, const char* token) {
  std::istringstream in(token);
  try {
    if (kind == 0) {
      SampleStrategy s;
      in >> s;
      return static_cast<int>(s);
    }
    PhiUpdaterMode m;
    in >> m;
    return static_cast<int>(m);
  } catch (const std::exception&) {
    return -1;
  }
}
// operator<<(Config), then "flags:" and the MakeCompileFlags list one per line
int mcmc_config_print_with_flags(void* vc, char* buf, size_t len) {
  std::ostringstream o;
  o << *static_cast<Config*>(vc) << "flags:\n";
  for (const std::string& f : MakeCompileFlags(*static_cast<Config*>(vc))) o << f << "\n";
  const std::string s = o.str();
  if (s.size() + 1 > len) return 1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return 0;
}
void mcmc_config_params(void* vc, ammsb_params* p) { *p = MakeParams(*static_cast<Config*>(vc)); }
// the training/held-out cuckoo tables as the device sees them
uint64_t mcmc_config_set_info(void* vc, int which, uint64_t* bins, uint32_t* prime) {
  Config* c = static_cast<Config*>(vc);
  Set* s = which == 0 ? c->training.get() : c->heldout.get();
  *bins = s->BinsPerBucket();
  *prime = s->PrimeIdx();
  return s->Capacity();
}
void mcmc_config_set_table(void* vc, int which, uint64_t* table) {
  Config* c = static_cast<Config*>(vc);
  Set* s = which == 0 ? c->training.get() : c->heldout.get();
  std::vector<Edge> t = s->Serialize();
  std::memcpy(table, t.data(), 8 * t.size());
}

// one host mini-batch with the given strategy (enum order of sample.h)
float mcmc_sample(void* vc, int strategy, unsigned* seed, uint64_t* edges_out, uint64_t* n_edges,
                  uint32_t* nodes_out, uint64_t* n_nodes) {
  Config* c = static_cast<Config*>(vc);
  std::vector<Edge> edges;
  float w = 0;
  switch (strategy) {
    case Node: w = sampleNode(*c, &edges, seed); break;
    case NodeLink: w = sampleNodeLink(*c, &edges, seed); break;
    case NodeNonLink: w = sampleNodeNonLink(*c, &edges, seed); break;
    case BFLink: w = sampleBreadthFirstLink(*c, &edges, seed); break;
    case BFNonLink: w = sampleBreadthFirstNonLink(*c, &edges, seed); break;
    default: w = sampleBreadthFirst(*c, &edges, seed); break;
  }
  std::vector<Vertex> nodes;
  ExtractNodesFromMiniBatch(edges, &nodes);
  std::memcpy(edges_out, edges.data(), 8 * edges.size());
  std::memcpy(nodes_out, nodes.data(), 4 * nodes.size());
  *n_edges = edges.size();
  *n_nodes = nodes.size();
  return w;
}

// standalone host cuckoo build (parity of the table image with the oracle/reference)
int mcmc_host_set_build(const uint64_t* keys, uint64_t n, uint64_t* table_out, uint64_t table_cap,
                        uint64_t* bins, uint32_t* prime, uint64_t* size) {
  std::vector<Edge> v(keys, keys + n);
  Set s(n);
  const bool ok = s.SetContents(v.begin(), v.end());
  *bins = s.BinsPerBucket();
  *prime = s.PrimeIdx();
  *size = s.Size();
  std::vector<Edge> t = s.Serialize();
  if (t.size() <= table_cap) std::memcpy(table_out, t.data(), 8 * t.size());
  return ok ? 1 : 0;
}
uint64_t mcmc_host_set_bins(uint64_t n) { return Set(n).BinsPerBucket(); }

// test hook: iteration order of StdOrderSet vs std::unordered_set for one insert sequence
// (width 8: Edge keys, width 4: Vertex keys); returns the number of distinct keys
uint64_t mcmc_test_set_order(const uint64_t* keys, uint64_t n, int width, uint64_t* out_std, uint64_t* out_flat) {
  std::vector<uint64_t> a, b;
  if (width == 8) {
    std::unordered_set<Edge> ref;
    StdOrderSet<Edge> flat;
    for (uint64_t i = 0; i < n; ++i) {
      const bool x = ref.insert(keys[i]).second, y = flat.Insert(keys[i]);
      if (x != y) return ~0ull;
    }
    a.assign(ref.begin(), ref.end());
    flat.EmitTo(&b);
  } else {
    std::unordered_set<Vertex> ref;
    StdOrderSet<Vertex> flat;
    for (uint64_t i = 0; i < n; ++i) {
      const bool x = ref.insert(static_cast<Vertex>(keys[i])).second, y = flat.Insert(static_cast<Vertex>(keys[i]));
      if (x != y) return ~0ull;
    }
    a.assign(ref.begin(), ref.end());
    flat.EmitTo(&b);
  }
  std::memcpy(out_std, a.data(), 8 * a.size());
  std::memcpy(out_flat, b.data(), 8 * b.size());
  return a.size() == b.size() ? a.size() : ~0ull;
}

// data.h entry points for the harness: the SNAP text loader (after srand(seed), as main.cc leaves
// libc's generator) and the gzip dataset dump of main.cc:109-143
int64_t mcmc_unique_edges_from_file(const char* path, unsigned srand_seed, uint64_t* count_vertices,
                                    uint64_t* edges_out, uint64_t cap) {
  std::vector<Edge> vals;
  srand(srand_seed);
  if (!GetUniqueEdgesFromFile(path, count_vertices, &vals)) return -1;
  if (vals.size() > cap) return -2;
  std::memcpy(edges_out, vals.data(), 8 * vals.size());
  return static_cast<int64_t>(vals.size());
}
int mcmc_dump_dataset(const char* path, uint64_t N, float heldout_ratio, const uint64_t* edges, uint64_t n) {
  return DumpDataset(path, N, heldout_ratio, std::vector<Edge>(edges, edges + n)) ? 0 : 1;
}
int64_t mcmc_load_dataset(const char* path, uint64_t* N, float* heldout_ratio, uint64_t* edges_out, uint64_t cap) {
  std::vector<Edge> vals;
  if (!LoadDataset(path, N, heldout_ratio, &vals)) return -1;
  if (vals.size() > cap) return -2;
  std::memcpy(edges_out, vals.data(), 8 * vals.size());
  return static_cast<int64_t>(vals.size());
}

// Graph::NeighborsOf (data.h) of the training (which = 0) or held-out graph; returns the degree
int64_t mcmc_config_neighbors(void* vc, int which, uint32_t u, uint32_t* out, uint64_t cap) {
  Config* c = static_cast<Config*>(vc);
  const Graph* g = which == 0 ? c->trainingGraph.get() : c->heldoutGraph.get();
  if (g == nullptr) return -1;
  const std::vector<Vertex>& adj = g->NeighborsOf(u);
  for (size_t i = 0; i < adj.size() && i < cap; ++i) out[i] = adj[i];
  return static_cast<int64_t>(adj.size());
}

// test hooks: exact remainder without a divide, and the per-endpoint view of a cuckoo set that
// the non-link strategy filters with (which: 0 training, 1 held-out)
void mcmc_test_fastmod(const uint64_t* a, const uint64_t* d, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = FastMod64(d[i]).Mod(a[i]);
}
uint64_t mcmc_config_set_has(void* vc, int which, const uint64_t* keys, uint64_t n, uint8_t* out) {
  Config* c = static_cast<Config*>(vc);
  const Set* s = which == 0 ? c->training.get() : c->heldout.get();
  for (uint64_t i = 0; i < n; ++i) out[i] = s->Has(keys[i]) ? 1 : 0;
  return n;
}
int64_t mcmc_config_partners(void* vc, int which, uint32_t u, uint32_t* out, uint64_t cap) {
  Config* c = static_cast<Config*>(vc);
  const Set* s = which == 0 ? c->training.get() : c->heldout.get();
  const Set::Partners* idx = s->PartnerIndex();
  if (idx == nullptr) return -1;
  uint64_t n = 0;
  for (const Vertex* v = idx->begin(u); v != idx->end(u); ++v, ++n)
    if (n < cap) out[n] = *v;
  return static_cast<int64_t>(n);
}

// test hooks: one checkpoint record (uint64 length + proto2 message, serialize.h) of each kind,
// written from / parsed into plain arrays.  kind: 0 BetaProperties, 1 PhiProperties,
// 2 PerplexityProperties, 3 SampleStorage, 4 LearnerProperties, 5 VectorStorage, 6 RpmProperties.
// ints/dbls hold the integer / double fields in field order; b1/b2 the bytes fields.
uint64_t mcmc_test_serialize(int kind, const uint64_t* ints, const double* dbls, const char* b1, uint64_t n1,
                             const char* b2, uint64_t n2, char* out, uint64_t cap) {
  std::ostringstream o;
  bool ok = false;
  switch (kind) {
    case 0: {
      BetaProperties m;
      m.count_calls = static_cast<uint32_t>(ints[0]);
      m.theta_sum_time = dbls[0]; m.grads_partial_time = dbls[1]; m.grads_sum_time = dbls[2];
      m.update_theta_time = dbls[3]; m.normalize_time = dbls[4];
      ok = SerializeMessage(&o, m);
      break;
    }
    case 1: {
      PhiProperties m;
      m.count_calls = static_cast<uint32_t>(ints[0]);
      m.update_phi_time = dbls[0]; m.update_pi_time = dbls[1];
      ok = SerializeMessage(&o, m);
      break;
    }
    case 2: {
      PerplexityProperties m;
      m.count_calls = static_cast<uint32_t>(ints[0]);
      m.ppx_time = dbls[0]; m.accumulate_time = dbls[1];
      ok = SerializeMessage(&o, m);
      break;
    }
    case 3: {
      SampleStorage m;
      m.edges.assign(b1, n1);
      m.nodes_vec.assign(b2, n2);
      m.seed = static_cast<uint32_t>(ints[0]);
      ok = SerializeMessage(&o, m);
      break;
    }
    case 4: {
      LearnerProperties m;
      m.stepCount = static_cast<uint32_t>(ints[0]);
      m.time = ints[1]; m.samplingTime = ints[2];
      m.phase = static_cast<int32_t>(static_cast<int64_t>(ints[3]));
      m.weight = dbls[0];
      ok = SerializeMessage(&o, m);
      break;
    }
    case 5: ok = SerializeBytes(&o, b1, n1); break;
    case 6:
      ok = WriteRpmProperties(&o, static_cast<uint32_t>(ints[0]), static_cast<uint32_t>(ints[1]),
                              static_cast<uint32_t>(ints[2]));
      break;
  }
  const std::string s = o.str();
  if (!ok || s.size() > cap) return ~0ull;
  std::memcpy(out, s.data(), s.size());
  return s.size();
}
int mcmc_test_parse(int kind, const char* bytes, uint64_t n, uint64_t* ints, double* dbls, char* b1, uint64_t* n1,
                    char* b2, uint64_t* n2) {
  std::istringstream in(std::string(bytes, n));
  switch (kind) {
    case 0: {
      BetaProperties m;
      if (!ParseMessage(&in, &m)) return 1;
      ints[0] = m.count_calls;
      dbls[0] = m.theta_sum_time; dbls[1] = m.grads_partial_time; dbls[2] = m.grads_sum_time;
      dbls[3] = m.update_theta_time; dbls[4] = m.normalize_time;
      return 0;
    }
    case 1: {
      PhiProperties m;
      if (!ParseMessage(&in, &m)) return 1;
      ints[0] = m.count_calls; dbls[0] = m.update_phi_time; dbls[1] = m.update_pi_time;
      return 0;
    }
    case 2: {
      PerplexityProperties m;
      if (!ParseMessage(&in, &m)) return 1;
      ints[0] = m.count_calls; dbls[0] = m.ppx_time; dbls[1] = m.accumulate_time;
      return 0;
    }
    case 3: {
      SampleStorage m;
      if (!ParseMessage(&in, &m) || m.edges.size() > *n1 || m.nodes_vec.size() > *n2) return 1;
      std::memcpy(b1, m.edges.data(), m.edges.size());
      std::memcpy(b2, m.nodes_vec.data(), m.nodes_vec.size());
      *n1 = m.edges.size(); *n2 = m.nodes_vec.size(); ints[0] = m.seed;
      return 0;
    }
    case 4: {
      LearnerProperties m;
      if (!ParseMessage(&in, &m)) return 1;
      ints[0] = m.stepCount; ints[1] = m.time; ints[2] = m.samplingTime;
      ints[3] = static_cast<uint64_t>(static_cast<int64_t>(m.phase));
      dbls[0] = m.weight;
      return 0;
    }
    case 5: return ParseBytes(&in, b1, *n1) ? 0 : 1;  // *n1 = expected size (mismatch fails)
    case 6: {
      uint32_t r = 0, c = 0, b = 0;
      if (!ReadRpmProperties(&in, &r, &c, &b)) return 1;
      ints[0] = r; ints[1] = c; ints[2] = b;
      return 0;
    }
  }
  return 1;
}

// theta init stream of Learner::Learner (host mt19937 + gamma_distribution)
void mcmc_init_theta_host(uint32_t K, float eta0, float eta1, float* theta_out) {
  std::mt19937 engine(6342455113);
  std::gamma_distribution<Float> dist(eta0, eta1);
  auto gamma = std::bind(dist, engine);
  std::vector<Float> host(2 * K);
  std::generate(host.begin(), host.end(), gamma);
  std::memcpy(theta_out, host.data(), 4 * host.size());
}

// ---- ShardedLearner (several GPUs of one box, column-sharded pi) ----
void* mcmc_sharded_create(void* vc, const int* devices, int n) {
  ShardedLearner* l = nullptr;
  const int rc = Guard([&] { l = new ShardedLearner(*static_cast<Config*>(vc), std::vector<int>(devices, devices + n)); });
  return rc ? nullptr : l;
}
void mcmc_sharded_destroy(void* v) { delete static_cast<ShardedLearner*>(v); }
int mcmc_sharded_run(void* v, uint32_t iters) {
  return Guard([&] { static_cast<ShardedLearner*>(v)->Run(iters); });
}
int mcmc_sharded_heldout_perplexity(void* v, float* out) {
  return Guard([&] { *out = static_cast<ShardedLearner*>(v)->HeldoutPerplexity(); });
}
uint64_t mcmc_sharded_edges_processed(void* v) { return static_cast<ShardedLearner*>(v)->EdgesProcessed(); }
// pi [N][K], phi [N]; theta / beta [world][2K]: every rank's copy (they must agree)
int mcmc_sharded_read(void* v, float* pi, float* phi, float* beta, float* theta, uint64_t N, uint64_t K) {
  return Guard([&] {
    ShardedLearner* l = static_cast<ShardedLearner*>(v);
    if (pi) l->ReadPi(0, N, pi);
    if (phi) l->ReadPhi(0, N, phi);
    for (uint32_t r = 0; r < l->World(); ++r)
      if (beta || theta) l->ReadTheta(r, theta ? theta + 2 * K * r : nullptr, beta ? beta + 2 * K * r : nullptr);
  });
}

// ---- Learner ----
struct LearnerBox {
  clcuda::Context ctx;
  clcuda::Queue queue;
  std::unique_ptr<Learner> learner;
};

void* mcmc_learner_create(void* vc, int device) {
  LearnerBox* b = nullptr;
  const int rc = Guard([&] {
    b = new LearnerBox();
    clcuda::Device dev(device);
    b->ctx = clcuda::Context(dev);
    b->queue = clcuda::Queue(b->ctx, dev);
    b->learner.reset(new Learner(*static_cast<Config*>(vc), b->queue));
  });
  if (rc) {
    delete b;
    return nullptr;
  }
  return b;
}
void mcmc_learner_destroy(void* vb) { delete static_cast<LearnerBox*>(vb); }
int mcmc_learner_run(void* vb, uint32_t iters) {
  return Guard([&] { static_cast<LearnerBox*>(vb)->learner->Run(iters); });
}
int mcmc_learner_heldout_perplexity(void* vb, float* out) {
  return Guard([&] { *out = static_cast<LearnerBox*>(vb)->learner->HeldoutPerplexity(); });
}
int mcmc_learner_training_perplexity(void* vb, float* out) {
  return Guard([&] { *out = static_cast<LearnerBox*>(vb)->learner->TrainingPerplexity(); });
}
uint64_t mcmc_learner_train_ppx_edges(void* vb, uint64_t* out) {
  const std::vector<Edge>& e = static_cast<LearnerBox*>(vb)->learner->TrainingPerplexityEdges();
  if (out) std::memcpy(out, e.data(), 8 * e.size());
  return e.size();
}
int mcmc_learner_print_stats(void* vb) {
  return Guard([&] { static_cast<LearnerBox*>(vb)->learner->PrintStats(); });
}
int mcmc_learner_read(void* vb, float* pi /* [N,K] or null */, float* phi, float* beta, float* theta,
                      uint64_t N) {
  return Guard([&] {
    Learner* l = static_cast<LearnerBox*>(vb)->learner.get();
    if (pi) l->ReadPi(0, N, pi);
    if (phi) l->ReadPhi(phi);
    if (beta) l->ReadBeta(beta);
    if (theta) l->ReadTheta(theta);
  });
}
int mcmc_learner_mirror_beta(void* vb, float* pinned_host) {
  return Guard([&] { static_cast<LearnerBox*>(vb)->learner->MirrorBetaTo(pinned_host); });
}
uint64_t mcmc_learner_h2d_bytes(void* vb) { return static_cast<LearnerBox*>(vb)->learner->BytesH2D(); }
uint64_t mcmc_learner_edges_processed(void* vb) { return static_cast<LearnerBox*>(vb)->learner->EdgesProcessed(); }
// the mini-batch the next Run() iteration will consume
int mcmc_learner_peek(void* vb, uint64_t* edges, uint64_t* n_edges, uint32_t* nodes, uint64_t* n_nodes,
                      uint32_t* neighbors /* [n_nodes, n] */, uint32_t n, float* weight) {
  return Guard([&] {
    LearnerBox* b = static_cast<LearnerBox*>(vb);
    const SampleSlot& s = b->learner->PeekNextSample();
    *weight = s.weight;
    *n_edges = s.edges.size();
    *n_nodes = s.nodes_vec.size();
    std::memcpy(edges, s.edges.data(), 8 * s.edges.size());
    std::memcpy(nodes, s.nodes_vec.data(), 4 * s.nodes_vec.size());
    s.neighbors.Read(b->queue, s.nodes_vec.size() * n, neighbors);
  });
}
int mcmc_learner_serialize(void* vb, const char* path) {
  return Guard([&] {
    std::ofstream out(path, std::ios::binary);
    if (!static_cast<LearnerBox*>(vb)->learner->Serialize(&out)) throw std::runtime_error("Serialize failed");
  });
}
int mcmc_learner_parse(void* vb, const char* path) {
  return Guard([&] {
    std::ifstream in(path, std::ios::binary);
    if (!static_cast<LearnerBox*>(vb)->learner->Parse(&in)) throw std::runtime_error("Parse failed");
  });
}

}  // extern "C"
