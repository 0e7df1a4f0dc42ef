#include "mcmc/sharded_learner.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <map>
#include <random>

using namespace std::chrono;

namespace mcmc {

namespace {
typedef Float (*SamplerFn)(const Config&, std::vector<Edge>*, unsigned int*);
SamplerFn PickStrategy(SampleStrategy s) {
  switch (s) {
    case NodeLink: return sampleNodeLink;
    case NodeNonLink: return sampleNodeNonLink;
    case Node: return sampleNode;
    case BFLink: return sampleBreadthFirstLink;
    case BFNonLink: return sampleBreadthFirstNonLink;
    case BF: return sampleBreadthFirst;
  }
  throw std::invalid_argument("Unknown sample strategy");
}
}  // namespace

// the ranks that live on one device: one stream, one copy of the replicated inputs, and per rank the
// column shard with its RNG pools (same sizes and seeds as the one-GPU operators: a rank only
// touches the states of its own lanes / columns)
struct ShardedLearner::Group {
  int device = 0;
  ammsb_ctx* ctx = nullptr;
  std::vector<ammsb_cols*> cols;
  std::vector<ammsb_rng*> phi_pool, beta_pool, nb_pool[2];
  ammsb_set *training = nullptr, *heldout = nullptr;
  uint64_t* d_edges[2] = {nullptr, nullptr};
  Vertex* d_nodes[2] = {nullptr, nullptr};
  uint64_t* d_heldout = nullptr;
  // pinned staging of a mini-batch slot and the event that says its copy has left the host
  uint64_t* h_edges[2] = {nullptr, nullptr};
  Vertex* h_nodes[2] = {nullptr, nullptr};
  ammsb_event* staged[2] = {nullptr, nullptr};

  ~Group() {
    for (ammsb_cols* c : cols) ammsb_cols_destroy(c);
    for (auto* v : {&phi_pool, &beta_pool, &nb_pool[0], &nb_pool[1]})
      for (ammsb_rng* r : *v) ammsb_rng_destroy(r);
    if (training) ammsb_set_destroy(training);
    if (heldout) ammsb_set_destroy(heldout);
    for (int b = 0; b < 2; ++b) {
      if (d_edges[b]) ammsb_free(ctx, d_edges[b]);
      if (d_nodes[b]) ammsb_free(ctx, d_nodes[b]);
      if (h_edges[b]) ammsb_host_free(h_edges[b]);
      if (h_nodes[b]) ammsb_host_free(h_nodes[b]);
      if (staged[b]) ammsb_event_destroy(staged[b]);
    }
    if (d_heldout) ammsb_free(ctx, d_heldout);
    if (ctx) ammsb_ctx_destroy(ctx);
  }
};

ShardedLearner::ShardedLearner(const Config& cfg, const std::vector<int>& devices)
    : cfg_(cfg), params_(MakeParams(cfg)), opts_(MakePhiOpts(cfg)), sampler_(PickStrategy(cfg.strategy)) {
  const uint32_t G = static_cast<uint32_t>(devices.size());
  if (G != 2 && G != 4 && G != 8) throw BackendError("ShardedLearner: 2, 4 or 8 ranks");
  if (cfg.phi_mode == PHI_NODE_PER_THREAD || cfg.phi_wg_size != 32)
    throw BackendError("ShardedLearner reproduces the work-group launch with phi_wg_size 32");
  const uint64_t max_nodes = MaxMiniBatchNodes(cfg), max_edges = MaxMiniBatchEdges(cfg);
  const uint32_t n = static_cast<uint32_t>(cfg.num_node_sample);
  ranks_.assign(G, nullptr);
  std::map<int, Group*> by_device;
  const std::vector<Edge> train_image = cfg.training->Serialize(), heldout_image = cfg.heldout->Serialize();
  for (uint32_t r = 0; r < G; ++r) {
    Group*& g = by_device[devices[r]];
    if (g == nullptr) {
      groups_.emplace_back(new Group());
      g = groups_.back().get();
      g->device = devices[r];
      AmmsbCheck(ammsb_ctx_create(devices[r], &g->ctx));
      AmmsbCheck(ammsb_set_create(g->ctx, train_image.data(), cfg.training->BinsPerBucket(), cfg.training->PrimeIdx(),
                                  &g->training));
      AmmsbCheck(ammsb_set_create(g->ctx, heldout_image.data(), cfg.heldout->BinsPerBucket(), cfg.heldout->PrimeIdx(),
                                  &g->heldout));
      for (int b = 0; b < 2; ++b) {
        AmmsbCheck(ammsb_malloc(g->ctx, sizeof(Edge) * max_edges, reinterpret_cast<void**>(&g->d_edges[b])));
        AmmsbCheck(ammsb_malloc(g->ctx, sizeof(Vertex) * max_nodes, reinterpret_cast<void**>(&g->d_nodes[b])));
        AmmsbCheck(ammsb_host_alloc(sizeof(Edge) * max_edges, reinterpret_cast<void**>(&g->h_edges[b])));
        AmmsbCheck(ammsb_host_alloc(sizeof(Vertex) * max_nodes, reinterpret_cast<void**>(&g->h_nodes[b])));
        AmmsbCheck(ammsb_event_create(g->ctx, &g->staged[b]));
      }
      const size_t H = cfg.heldout_edges.size();
      AmmsbCheck(ammsb_malloc(g->ctx, sizeof(Edge) * std::max<size_t>(H, 1), reinterpret_cast<void**>(&g->d_heldout)));
      if (H) AmmsbCheck(ammsb_h2d(g->ctx, g->d_heldout, cfg.heldout_edges.data(), sizeof(Edge) * H));
    }
    ammsb_cols* c = nullptr;
    AmmsbCheck(ammsb_cols_create(g->ctx, cfg.N, static_cast<uint32_t>(cfg.K), G, r, n, static_cast<uint32_t>(max_nodes),
                                 static_cast<uint32_t>(max_edges), cfg.heldout_edges.size(), &c));
    g->cols.push_back(c);
    ranks_[r] = c;
    ammsb_rng* pool = nullptr;
    // phi.cc:625-629, beta.cc (K states), sample.cc:86-97 (one pool per sampler stream)
    AmmsbCheck(ammsb_rng_create(g->ctx, std::min<uint64_t>(max_nodes, 65535) * 32, cfg.phi_seed[0], cfg.phi_seed[1], &pool));
    g->phi_pool.push_back(pool);
    AmmsbCheck(ammsb_rng_create(g->ctx, cfg.K, cfg.beta_seed[0], cfg.beta_seed[1], &pool));
    g->beta_pool.push_back(pool);
    for (int s = 0; s < 2; ++s) {
      AmmsbCheck(ammsb_rng_create(g->ctx, max_nodes, cfg.neighbor_seed[0], cfg.neighbor_seed[1], &pool));
      g->nb_pool[s].push_back(pool);
    }
  }
  // every rank maps every other rank's mailbox (peer access inside the process)
  for (uint32_t a = 0; a < G; ++a)
    for (uint32_t b = 0; b < G; ++b)
      if (a != b) AmmsbCheck(ammsb_cols_attach_local(ranks_[a], b, ranks_[b]));
  // the two sampler streams draw their seeds as the reference's Samples do (sample.cc:132)
  seeds_[0] = rand();
  seeds_[1] = rand();
  // theta ~ Gamma(eta0, eta1) on the host, fixed seed; beta = theta row-normalised (learner.cc:149-153)
  std::mt19937 engine(6342455113);
  std::gamma_distribution<Float> gamma_distribution(cfg.eta0, cfg.eta1);
  auto gamma = std::bind(gamma_distribution, engine);
  std::vector<Float> theta(2 * cfg.K), beta(2 * cfg.K);
  std::generate(theta.begin(), theta.end(), gamma);
  for (uint64_t k = 0; k < cfg.K; ++k) {
    const Float sum = (Float(0) + theta[2 * k]) + theta[2 * k + 1];  // normalize.cc:13-32, rows of 2
    beta[2 * k] = theta[2 * k] / sum;
    beta[2 * k + 1] = theta[2 * k + 1] / sum;
  }
  for (ammsb_cols* c : ranks_) {
    AmmsbCheck(ammsb_cols_init_pi(c, cfg.eta0, cfg.eta1));  // learner.cc:154-155
    AmmsbCheck(ammsb_cols_write_theta(c, theta.data(), beta.data()));
  }
  for (auto& g : groups_) AmmsbCheck(ammsb_ctx_sync(g->ctx));
  for (int s = 0; s < 2; ++s) producers_[s] = std::thread(&ShardedLearner::Producer, this, s);
}

ShardedLearner::~ShardedLearner() {
  {
    std::unique_lock<std::mutex> lock(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (std::thread& t : producers_)
    if (t.joinable()) t.join();
  for (auto& g : groups_) ammsb_ctx_sync(g->ctx);
}

void ShardedLearner::Producer(int stream) {
  std::unique_lock<std::mutex> lock(mu_);
  for (;;) {
    cv_.wait(lock, [&] { return stop_ || (ready_[stream].size() < kAhead && !error_); });
    if (stop_) return;
    lock.unlock();
    MiniBatch mb;
    std::exception_ptr err;
    try {
      mb = Draw(stream);
    } catch (...) {
      err = std::current_exception();
    }
    lock.lock();
    if (err) error_ = err; else ready_[stream].push_back(std::move(mb));
    cv_.notify_all();
  }
}

ShardedLearner::MiniBatch ShardedLearner::Next(int stream) {
  std::unique_lock<std::mutex> lock(mu_);
  cv_.wait(lock, [&] { return !ready_[stream].empty() || error_; });
  if (ready_[stream].empty()) {
    std::exception_ptr e = error_;
    error_ = nullptr;
    std::rethrow_exception(e);
  }
  MiniBatch mb = std::move(ready_[stream].front());
  ready_[stream].pop_front();
  cv_.notify_all();
  return mb;
}

ShardedLearner::MiniBatch ShardedLearner::Draw(int stream) {
  MiniBatch mb;
  mb.weight = sampler_(cfg_, &mb.edges, &seeds_[stream]);
  ExtractNodesFromMiniBatch(mb.edges, &mb.nodes);
  if (mb.nodes.empty()) throw BackendError("mini-batch size = 0!");
  if (mb.edges.size() > MaxMiniBatchEdges(cfg_) || mb.nodes.size() > MaxMiniBatchNodes(cfg_))
    throw BackendError("mini-batch exceeds the device buffers");
  return mb;
}

void ShardedLearner::Upload(const MiniBatch& mb, int slot) {
  for (auto& g : groups_) {
    AmmsbCheck(ammsb_event_sync(g->staged[slot]));  // the copy of two mini-batches ago has left the staging buffer
    std::copy(mb.edges.begin(), mb.edges.end(), g->h_edges[slot]);
    std::copy(mb.nodes.begin(), mb.nodes.end(), g->h_nodes[slot]);
    AmmsbCheck(ammsb_h2d_async(g->ctx, g->d_edges[slot], g->h_edges[slot], sizeof(Edge) * mb.edges.size()));
    AmmsbCheck(ammsb_h2d_async(g->ctx, g->d_nodes[slot], g->h_nodes[slot], sizeof(Vertex) * mb.nodes.size()));
    AmmsbCheck(ammsb_event_record(g->ctx, g->staged[slot]));
  }
}

void ShardedLearner::Run(uint32_t max_iters, sig_atomic_t* signaled) {
  const auto t1 = high_resolution_clock::now();
  for (uint32_t i = 0; i < max_iters && (signaled == nullptr || !*signaled); ++i) {
    const auto ts = high_resolution_clock::now();
    // mini-batch t comes from sampler stream `phase_` (learner.cc:216-232)
    MiniBatch mb = Next(phase_);
    samplingTime_ += duration_cast<nanoseconds>(high_resolution_clock::now() - ts).count();
    ++stepCount_;
    const int slot = stepCount_ & 1;
    const uint32_t V = static_cast<uint32_t>(mb.nodes.size()), E_mb = static_cast<uint32_t>(mb.edges.size());
    Upload(mb, slot);
    // a stage is launched on every device before the next one: the ranks' kernels meet in their mailboxes
    for (auto& g : groups_)
      AmmsbCheck(ammsb_cols_neighbor_sample(g->ctx, g->cols.data(), static_cast<uint32_t>(g->cols.size()),
                                            g->d_nodes[slot], V, cfg_.neighbor_sampler_wg_size, stepCount_,
                                            g->nb_pool[phase_].data()));
    for (auto& g : groups_)
      AmmsbCheck(ammsb_cols_update_phi(g->ctx, g->cols.data(), static_cast<uint32_t>(g->cols.size()), &params_, &opts_,
                                       g->training, g->d_nodes[slot], nullptr, V, stepCount_, g->phi_pool.data()));
    for (auto& g : groups_)
      AmmsbCheck(ammsb_cols_update_pi(g->ctx, g->cols.data(), static_cast<uint32_t>(g->cols.size()), g->d_nodes[slot], V,
                                      stepCount_));
    for (auto& g : groups_)
      AmmsbCheck(ammsb_cols_update_beta(g->ctx, g->cols.data(), static_cast<uint32_t>(g->cols.size()), &params_,
                                        g->training, g->d_edges[slot], E_mb, mb.weight, stepCount_,
                                        g->beta_pool.data()));
    edgesProcessed_ += E_mb;
    phase_ = 1 - phase_;
  }
  for (auto& g : groups_) AmmsbCheck(ammsb_ctx_sync(g->ctx));
  for (ammsb_cols* c : ranks_) {
    uint32_t timed_out = 0;
    AmmsbCheck(ammsb_cols_check(c, &timed_out));
    if (timed_out) throw BackendError("a rank gave up waiting for a peer's partial sums");
  }
  time_ += duration_cast<nanoseconds>(high_resolution_clock::now() - t1).count();
}

Float ShardedLearner::HeldoutPerplexity() {
  ++ppxCalls_;
  const uint32_t H = static_cast<uint32_t>(cfg_.heldout_edges.size());
  for (auto& g : groups_)
    AmmsbCheck(ammsb_cols_perplexity(g->ctx, g->cols.data(), static_cast<uint32_t>(g->cols.size()), &params_, g->heldout,
                                     g->d_heldout, H, ppxCalls_, nullptr, nullptr));
  double avg = 0;
  AmmsbCheck(ammsb_cols_perplexity_result(groups_[0]->ctx, groups_[0]->cols[0], nullptr, &avg));
  for (auto& g : groups_) AmmsbCheck(ammsb_ctx_sync(g->ctx));
  return std::exp(static_cast<Float>(avg));
}

void ShardedLearner::PrintStats() {
  std::cerr << "TOTAL    : " << time_ / 1.0e9 << std::endl;
  std::cerr << "SAMPLING : " << samplingTime_ / 1.0e9 << std::endl;
  std::cerr << "ITERATIONS  : " << stepCount_ << ", MINI-BATCH EDGES: " << edgesProcessed_ << ", RANKS: " << World()
            << std::endl;
}

void ShardedLearner::ReadPi(uint64_t row0, uint64_t nrows, Float* rows) {
  for (ammsb_cols* c : ranks_) AmmsbCheck(ammsb_cols_read_pi(c, row0, nrows, rows));  // each fills its own columns
}

void ShardedLearner::ReadPhi(uint64_t row0, uint64_t nrows, Float* sums) {
  AmmsbCheck(ammsb_cols_read_phi(ranks_[0], row0, nrows, sums));
}

void ShardedLearner::ReadTheta(uint32_t rank, Float* theta, Float* beta) {
  if (rank >= ranks_.size()) throw BackendError("rank out of range");
  AmmsbCheck(ammsb_cols_read_theta(ranks_[rank], theta, beta));
}

}  // namespace mcmc
