#include "mcmc/sample.h"

#include <algorithm>
#include <chrono>
#include <cctype>
#include <queue>
#include <unordered_set>

#include "mcmc/config.h"
#include "mcmc/serialize.h"
#include "mcmc/std_order_set.h"

namespace mcmc {

namespace {

// Bump allocator for the short-lived hash sets of one mini-batch.  The allocator of a
// std::unordered_set changes neither its bucket-count sequence nor its iteration order (the
// output contract of the strategies); it only removes a malloc/free pair per element and keeps
// the nodes contiguous.  Memory is recycled per thread when the last set of a scope dies.
class Arena {
 public:
  void* Allocate(size_t bytes) {
    bytes = (bytes + 15) & ~size_t(15);
    if (blocks_.empty() || used_ + bytes > blocks_[cur_].size()) NextBlock(bytes);
    void* p = blocks_[cur_].data() + used_;
    used_ += bytes;
    return p;
  }
  void Enter() { ++live_; }
  void Leave() {
    if (--live_ == 0) {
      cur_ = 0;
      used_ = 0;
    }
  }

 private:
  void NextBlock(size_t bytes) {
    while (!blocks_.empty() && cur_ + 1 < blocks_.size()) {
      ++cur_;
      used_ = 0;
      if (bytes <= blocks_[cur_].size()) return;
    }
    blocks_.emplace_back(std::max<size_t>(bytes, size_t(1) << 20));
    cur_ = blocks_.size() - 1;
    used_ = 0;
  }
  std::vector<std::vector<char>> blocks_;
  size_t cur_ = 0, used_ = 0;
  int live_ = 0;
};
thread_local Arena t_arena;

struct ArenaScope {
  ArenaScope() { t_arena.Enter(); }
  ~ArenaScope() { t_arena.Leave(); }
};

template <class T>
struct ArenaAlloc {
  typedef T value_type;
  ArenaAlloc() {}
  template <class U>
  ArenaAlloc(const ArenaAlloc<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(t_arena.Allocate(n * sizeof(T))); }
  void deallocate(T*, size_t) {}
  template <class U>
  bool operator==(const ArenaAlloc<U>&) const { return true; }
  template <class U>
  bool operator!=(const ArenaAlloc<U>&) const { return false; }
};
// same hash, equality and growth policy as the reference's std::unordered_set<T>
template <class T>
using HashSet = std::unordered_set<T, std::hash<T>, std::equal_to<T>, ArenaAlloc<T>>;

inline Edge Canonical(Vertex u, Vertex v) { return MakeEdge(std::min(u, v), std::max(u, v)); }
inline Vertex DrawVertex(const Config& cfg, unsigned int* seed) { return rand_r(seed) % cfg.N; }
template <class SetT>
inline void Emit(const SetT& picked, std::vector<Edge>* edges) {
  edges->insert(edges->begin(), picked.begin(), picked.end());  // std::unordered_set order is the contract
}
bool SameNoCase(const std::string& a, const char* b) {
  size_t i = 0;
  for (; i < a.size() && b[i]; ++i)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return i == a.size() && !b[i];
}
}  // namespace

uint64_t MaxMiniBatchNodes(const Config& cfg) {
  return std::max<uint64_t>(2 * cfg.mini_batch_size, 1 + cfg.trainingGraph->MaxFanOut());
}
uint64_t MaxMiniBatchEdges(const Config& cfg) {
  return std::max<uint64_t>(cfg.mini_batch_size, cfg.trainingGraph->MaxFanOut());
}

// ---- strategies (reference sample.cc:177-302) ----

// Pick unseen vertices u until one has training neighbors; the mini-batch is every training
// edge of u.  Scale N.
Float sampleNodeLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  ArenaScope scope;
  HashSet<Vertex> tried;
  HashSet<Edge> picked;
  while (picked.empty()) {
    const Vertex u = DrawVertex(cfg, seed);
    if (!tried.insert(u).second) continue;
    for (Vertex v : cfg.trainingGraph->NeighborsOf(u)) picked.insert(Canonical(u, v));
  }
  Emit(picked, edges);
  return static_cast<Float>(cfg.N);
}

// One vertex u; draw v until (u,v) is in neither the held-out nor the training set, m
// distinct pairs.  (u == v is not excluded and no v is ever blacklisted -- kept as is.)
// Scale 2E/m.
//
// The reference consumes one rand_r draw per candidate whatever its fate and asks both cuckoo
// sets about every candidate (two random 32-byte bins each: four cache misses per draw).  All
// candidates of a mini-batch share the endpoint u, so here the sets are asked once, up front:
// the stored keys that involve u (Set::PartnerIndex(), a handful) are marked as refused in the
// very table that de-duplicates the picks, and a candidate then costs its rand_r draw plus the
// one probe it needed anyway.  Same answers, same draw count, same stream position.  Sets too
// large to index fall back to block-wise hashing with prefetch.
Float sampleNodeNonLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  // std::unordered_set<Edge> order, flat storage; bound to a reference once: in a shared
  // library every direct use of a thread_local goes through the TLS wrapper
  thread_local StdOrderSet<Edge> tls_picked;
  StdOrderSet<Edge>& picked = tls_picked;
  picked.Clear();
  const Vertex u = DrawVertex(cfg, seed);
  const size_t m = cfg.mini_batch_size;
  const uint32_t n_vertices = static_cast<uint32_t>(cfg.N);
  const bool narrow = cfg.N <= 0xffffffffull;  // rand_r() % N in 32 bits is the same number
  const Set::Partners* in_heldout = cfg.heldout ? cfg.heldout->PartnerIndex() : nullptr;
  const Set::Partners* in_training = cfg.training->PartnerIndex();
  unsigned int s = *seed;
  // All candidates share u, so a candidate is identified by its other endpoint v: when a bit per
  // vertex is smaller than the hash table the picks would need (N/8 bytes against 16-32 bytes per
  // pick: every graph but the very largest), "seen before or refused" is one bit test -- 40 KB that
  // stay in L1/L2 at the DBLP shape, where a table for 131072 picks is 4 MB of cache misses
  // (measured: 11.5 -> 1.6 ms for the picks of an m = 131072 mini-batch).  The bitmap is all-zero
  // between calls: the bits set here are cleared again one by one.
  const bool by_vertex = narrow && in_training != nullptr && (in_heldout != nullptr || !cfg.heldout) &&
                         cfg.N / 8 <= 32 * m;
  if (by_vertex) {
    thread_local std::vector<uint64_t> tls_seen;
    std::vector<uint64_t>& seen = tls_seen;
    if (seen.size() < (cfg.N + 63) / 64) seen.assign((cfg.N + 63) / 64, 0);
    for (const Set::Partners* idx : {in_heldout, in_training}) {
      if (idx == nullptr) continue;
      for (const Vertex* v = idx->begin(u); v != idx->end(u); ++v) seen[*v >> 6] |= uint64_t(1) << (*v & 63);
    }
    while (picked.size() < m) {
      const uint32_t v = static_cast<uint32_t>(rand_r(&s)) % n_vertices;
      const uint64_t bit = uint64_t(1) << (v & 63);
      uint64_t& word = seen[v >> 6];
      if (word & bit) continue;  // refused (a stored pair) or picked before
      word |= bit;
      picked.AppendUnique(Canonical(u, v));
    }
    for (Edge e : picked.InsertionOrder()) {
      const Vertex a = static_cast<Vertex>(e >> 32), b = static_cast<Vertex>(e & 0xffffffffu);
      const Vertex v = a == u ? b : a;
      seen[v >> 6] &= ~(uint64_t(1) << (v & 63));
    }
    for (const Set::Partners* idx : {in_heldout, in_training}) {
      if (idx == nullptr) continue;
      for (const Vertex* v = idx->begin(u); v != idx->end(u); ++v) seen[*v >> 6] &= ~(uint64_t(1) << (*v & 63));
    }
  } else if (narrow && in_training != nullptr && (in_heldout != nullptr || !cfg.heldout)) {
    picked.Reserve(m + 64);  // the m picks and the handful of blocked keys: no regrowth
    for (const Set::Partners* idx : {in_heldout, in_training}) {
      if (idx == nullptr) continue;
      for (const Vertex* v = idx->begin(u); v != idx->end(u); ++v) picked.Block(Canonical(u, *v));
    }
    while (picked.size() < m) picked.Insert(Canonical(u, static_cast<uint32_t>(rand_r(&s)) % n_vertices));
  } else {
    const int kBlock = 32;
    Edge cand[kBlock];
    unsigned int after[kBlock];
    size_t hb[kBlock][2], tb[kBlock][2];
    while (picked.size() < m) {
      const int take = static_cast<int>(std::min<size_t>(kBlock, m - picked.size()));
      for (int i = 0; i < take; ++i) {
        cand[i] = Canonical(u, rand_r(&s) % cfg.N);
        after[i] = s;
        cfg.heldout->Locate(cand[i], hb[i]);
        cfg.training->Locate(cand[i], tb[i]);
      }
      // at most `take` insertions can happen, so the block never overshoots m mid-way
      for (int i = 0; i < take; ++i) {
        if (cfg.heldout->HasAt(cand[i], hb[i]) || cfg.training->HasAt(cand[i], tb[i])) continue;
        picked.Insert(cand[i]);
      }
      s = after[take - 1];
    }
  }
  *seed = s;
  picked.EmitTo(edges);
  return (2 * cfg.E) / static_cast<Float>(cfg.mini_batch_size);
}

Float sampleNode(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  return (rand_r(seed) % 2) ? sampleNodeLink(cfg, edges, seed) : sampleNodeNonLink(cfg, edges, seed);
}

// Breadth-first over training links from random roots until m edges.  Scale E/m.
Float sampleBreadthFirstLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  ArenaScope scope;
  HashSet<Vertex> visited;
  std::queue<Vertex> frontier;
  HashSet<Edge> picked;
  while (picked.size() < cfg.mini_batch_size) {
    if (frontier.empty()) {
      Vertex root;
      do {
        root = DrawVertex(cfg, seed);
      } while (visited.count(root));
      frontier.push(root);
    }
    const Vertex u = frontier.front();
    frontier.pop();
    if (!visited.insert(u).second) continue;
    for (Vertex v : cfg.trainingGraph->NeighborsOf(u)) {
      if (picked.size() >= cfg.mini_batch_size) break;
      frontier.push(v);
      picked.insert(Canonical(u, v));
    }
  }
  Emit(picked, edges);
  return static_cast<Float>(cfg.E) / cfg.mini_batch_size;
}

// Breadth-first where each visited vertex contributes up to 32 random non-neighbors.
// Scale (N(N-1)/2 - E)/m.
Float sampleBreadthFirstNonLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  ArenaScope scope;
  HashSet<Vertex> visited;
  std::queue<Vertex> frontier;
  HashSet<Edge> picked;
  while (picked.size() < cfg.mini_batch_size) {
    if (frontier.empty()) {
      Vertex root;
      do {
        root = DrawVertex(cfg, seed);
      } while (visited.count(root));
      frontier.push(root);
    }
    const Vertex u = frontier.front();
    frontier.pop();
    if (!visited.insert(u).second) continue;
    const std::vector<Vertex>& adj = cfg.trainingGraph->NeighborsOf(u);
    for (uint32_t i = 0; i < 32 && picked.size() < cfg.mini_batch_size; ++i) {
      Vertex v;
      do {
        v = DrawVertex(cfg, seed);
      } while (u == v || std::find(adj.begin(), adj.end(), v) != adj.end());
      frontier.push(v);
      picked.insert(Canonical(u, v));
    }
  }
  Emit(picked, edges);
  return static_cast<Float>((cfg.N * (cfg.N - 1) / 2.0 - cfg.E) / cfg.mini_batch_size);
}

Float sampleBreadthFirst(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  return (rand_r(seed) % 2) ? sampleBreadthFirstLink(cfg, edges, seed) : sampleBreadthFirstNonLink(cfg, edges, seed);
}

void ExtractNodesFromMiniBatch(const std::vector<Edge>& edges, std::vector<Vertex>* nodes_vec) {
  thread_local StdOrderSet<Vertex> tls_nodes;  // std::unordered_set<Vertex> order (learner.cc:164-172)
  StdOrderSet<Vertex>& nodes = tls_nodes;
  nodes.Clear();
  // A Node-strategy mini-batch is a star: every edge contains one vertex u, and the edges are
  // distinct, so the other endpoints are distinct too -- the insert sequence of first occurrences is
  // known without a single hash probe: a0, b0, then the other endpoint of every further edge (an
  // edge (u, u) adds nothing).  Falls back to the general path at the first edge that is not on u.
  if (edges.size() >= 2) {
    const Vertex a0 = static_cast<Vertex>(edges[0] >> 32), b0 = static_cast<Vertex>(edges[0] & 0xffffffffu);
    const Vertex a1 = static_cast<Vertex>(edges[1] >> 32), b1 = static_cast<Vertex>(edges[1] & 0xffffffffu);
    const Vertex u = (a0 == a1 || a0 == b1) ? a0 : b0;
    bool star = true;
    nodes.AppendUnique(a0);
    if (b0 != a0) nodes.AppendUnique(b0);
    for (size_t i = 1; i < edges.size(); ++i) {
      const Vertex a = static_cast<Vertex>(edges[i] >> 32), b = static_cast<Vertex>(edges[i] & 0xffffffffu);
      Vertex other;
      if (a == u) other = b;
      else if (b == u) other = a;
      else {
        star = false;
        break;
      }
      // a == u is checked first, so the insert order inside the edge (a, then b) does not matter:
      // only `other` can be new
      if (other != u) nodes.AppendUnique(other);
    }
    if (star) {
      nodes_vec->clear();
      nodes.EmitTo(nodes_vec);
      return;
    }
    nodes.Clear();
  }
  if (edges.size() > 1024) nodes.Reserve(edges.size() + 1);
  // An endpoint shared with the previous edge is already in the set, and re-inserting a present
  // key never changes the set: skip it.  Node-strategy mini-batches share one endpoint
  // throughout, so this halves the lookups.
  Vertex p0 = 0, p1 = 0;
  bool have_prev = false;
  for (Edge e : edges) {
    const Vertex a = static_cast<Vertex>(e >> 32), b = static_cast<Vertex>(e & 0xffffffffu);
    if (!(have_prev && (a == p0 || a == p1))) nodes.Insert(a);
    if (!(have_prev && (b == p0 || b == p1))) nodes.Insert(b);
    p0 = a;
    p1 = b;
    have_prev = true;
  }
  nodes_vec->clear();
  nodes.EmitTo(nodes_vec);
}

std::istream& operator>>(std::istream& in, SampleStrategy& strategy) {
  std::string token;
  in >> token;
  static const struct { const char* name; SampleStrategy s; } kNames[] = {
      {"NodeLink", NodeLink}, {"NodeNonLink", NodeNonLink}, {"Node", Node},
      {"BFLink", BFLink},     {"BFNonLink", BFNonLink},     {"BF", BF}};
  for (const auto& n : kNames) {
    if (SameNoCase(token, n.name)) {
      strategy = n.s;
      return in;
    }
  }
  throw std::invalid_argument("Invalid SampleStrategy");
}

std::string to_string(const SampleStrategy& s) {
  switch (s) {
    case NodeLink: return "NodeLink";
    case NodeNonLink: return "NodeNonLink";
    case Node: return "Node";
    case BFLink: return "BFLink";
    case BFNonLink: return "BFNonLink";
    case BF: return "BF";
  }
  throw std::invalid_argument("Invalid strategy");
}

// ---- device neighbor sampler ----

NeighborSampler::NeighborSampler(const Config& cfg, clcuda::Queue queue)
    : cfg_(cfg),
      capacity_(2 * cfg.num_node_sample),
      local_(cfg.neighbor_sampler_wg_size),
      queue_(queue),
      hash_(queue_.GetContext(), MaxMiniBatchNodes(cfg) * capacity_),
      data_(queue_.GetContext(), MaxMiniBatchNodes(cfg) * cfg.num_node_sample),
      randFactory_(random::OpenClRandomFactory::New(queue_)),
      rand_(randFactory_->CreateRandom(MaxMiniBatchNodes(cfg) * capacity_,
                                       random::random_seed_t{cfg.neighbor_seed[0], cfg.neighbor_seed[1]})) {}

void NeighborSampler::operator()(uint32_t num_samples, clcuda::Buffer<Vertex>* nodes) {
  AmmsbCheck(ammsb_neighbor_sample(queue_(), rand_->Get(), nodes->data(), num_samples,
                                   static_cast<uint32_t>(cfg_.N), static_cast<uint32_t>(cfg_.num_node_sample),
                                   local_, data_.data(), export_hash_ ? hash_.data() : nullptr));
  queue_.Finish();
}

uint32_t NeighborSampler::DataSizePerSample() { return cfg_.num_node_sample; }

bool NeighborSampler::Serialize(std::ostream* out) {
  return rand_->Serialize(out) && ::mcmc::Serialize(out, &data_, &queue_);
}
bool NeighborSampler::Parse(std::istream* in) { return rand_->Parse(in) && ::mcmc::Parse(in, &data_, &queue_); }

void NeighborSampler::operator()(uint32_t num_samples, clcuda::Buffer<Vertex>* nodes, clcuda::Buffer<Vertex>* out) {
  AmmsbCheck(ammsb_neighbor_sample(queue_(), rand_->Get(), nodes->data(), num_samples,
                                   static_cast<uint32_t>(cfg_.N), static_cast<uint32_t>(cfg_.num_node_sample),
                                   local_, data_.data(), export_hash_ ? hash_.data() : nullptr));
  data_.CopyTo(queue_, static_cast<size_t>(num_samples) * cfg_.num_node_sample, *out);
  queue_.Finish();
}

SampleSlot::SampleSlot(const Config& cfg, const clcuda::Context& ctx)
    : dev_edges(ctx, MaxMiniBatchEdges(cfg)),
      dev_nodes(ctx, MaxMiniBatchNodes(cfg)),
      neighbors(ctx, MaxMiniBatchNodes(cfg) * cfg.num_node_sample) {}

Sample::Sample(const Config& cfg, clcuda::Queue q)
    : queue(q.GetContext(), q.GetDevice()),
      seed(rand()),
      neighbor_sampler(cfg, clcuda::Queue(q.GetContext(), q.GetDevice())),
      cfg_(cfg) {
  for (uint64_t i = 0; i < kRing; ++i) ring.emplace_back(new SampleSlot(cfg, q.GetContext()));
}

Sample::~Sample() {
  {
    std::unique_lock<std::mutex> lock(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  if (a_.joinable()) a_.join();
  if (b_.joinable()) b_.join();
  if (c_.joinable()) c_.join();
  if (dev_sampler_ != nullptr) ammsb_sampler_destroy(dev_sampler_);
  if (dev_order_ != nullptr) ammsb_orderset_destroy(dev_order_);
}

namespace {
// mcmc::Graph (data.cc:12-25) flattened: offsets[u] .. offsets[u + 1] index u's neighbors in the
// order NeighborsOf(u) lists them -- the order sampleNodeLink inserts the edges in
std::vector<uint64_t> GraphOffsets(const Config& cfg) {
  std::vector<uint64_t> off(cfg.N + 1, 0);
  for (uint64_t u = 0; u < cfg.N; ++u) off[u + 1] = off[u] + cfg.trainingGraph->NeighborsOf(static_cast<Vertex>(u)).size();
  return off;
}
std::vector<Vertex> GraphAdjacency(const Config& cfg, const std::vector<uint64_t>& off) {
  std::vector<Vertex> adj(std::max<uint64_t>(off.back(), 1));
  for (uint64_t u = 0; u < cfg.N; ++u) {
    const std::vector<Vertex>& nb = cfg.trainingGraph->NeighborsOf(static_cast<Vertex>(u));
    std::copy(nb.begin(), nb.end(), adj.begin() + off[u]);
  }
  return adj;
}
}  // namespace

DeviceStrategyData::DeviceStrategyData(const Config& cfg, clcuda::Queue queue, ammsb_set* tr, ammsb_set* he)
    : offsets(queue.GetContext(), cfg.N + 1),
      adjacency(queue.GetContext(), std::max<uint64_t>(2 * cfg.training_edges.size(), 1)),
      degree(cfg.N),
      training(tr),
      heldout(he) {
  if (!cfg.trainingGraph) throw BackendError("the device mini-batch strategies need Config::trainingGraph");
  const std::vector<uint64_t> off = GraphOffsets(cfg);
  const std::vector<Vertex> adj = GraphAdjacency(cfg, off);
  for (uint64_t u = 0; u < cfg.N; ++u) degree[u] = static_cast<uint32_t>(off[u + 1] - off[u]);
  offsets.Write(queue, off.size(), off.data());
  adjacency.Write(queue, off.back(), adj.data());
  queue.Finish();
}

void Sample::StartOnDevice(SampleStrategy which, std::shared_ptr<DeviceStrategyData> data, SamplerStats* stats) {
  if (which != Node && which != NodeLink && which != NodeNonLink)
    throw BackendError("device_sampler: only the Node strategies are drawn on the device (the breadth-first "
                       "strategies are queue-driven and stay on the host)");
  dev_ = data;
  dev_strategy_ = which;
  stats_ = stats;
  AmmsbCheck(ammsb_sampler_create(queue(), cfg_.N, static_cast<uint32_t>(cfg_.mini_batch_size), &dev_sampler_));
  AmmsbCheck(ammsb_orderset_create(queue(), static_cast<uint32_t>(2 * MaxMiniBatchEdges(cfg_) + 2), &dev_order_));
  a_ = std::thread(&Sample::StageA, this);
  c_ = std::thread(&Sample::StageC, this);
}

Float Sample::DrawOnDevice(SampleSlot* slot) {
  bool link = dev_strategy_ == NodeLink;
  if (dev_strategy_ == Node) link = (rand_r(&seed) % 2) != 0;  // sampleNode, sample.cc:295-302
  uint32_t num_edges = 0, num_nodes = 0;
  Float weight;
  if (link) {  // sampleNodeLink: vertices until one has training neighbors (sample.cc:253-272)
    Vertex u;
    do {
      u = static_cast<Vertex>(rand_r(&seed) % cfg_.N);
    } while (dev_->degree[u] == 0);
    num_edges = dev_->degree[u];
    AmmsbCheck(ammsb_minibatch_link(queue(), u, num_edges, dev_->offsets.data(), dev_->adjacency.data(),
                                    slot->dev_edges.data(), slot->dev_nodes.data()));
    weight = static_cast<Float>(cfg_.N);
  } else {  // sampleNodeNonLink (sample.cc:274-293)
    const Vertex u = static_cast<Vertex>(rand_r(&seed) % cfg_.N);
    AmmsbCheck(ammsb_minibatch_nonlink(dev_sampler_, queue(), u, &seed, dev_->training, dev_->heldout,
                                       slot->dev_edges.data(), slot->dev_nodes.data(), &num_edges, &num_nodes));
    weight = (2 * cfg_.E) / static_cast<Float>(cfg_.mini_batch_size);
  }
  AmmsbCheck(ammsb_minibatch_finish(dev_order_, queue(), slot->dev_edges.data(), num_edges, slot->dev_nodes.data(),
                                    &num_nodes));
  slot->edges.resize(num_edges);
  slot->nodes_vec.resize(num_nodes);
  slot->dev_edges.Read(queue, num_edges, slot->edges.data());
  slot->dev_nodes.Read(queue, num_nodes, slot->nodes_vec.data());
  if (slot->nodes_vec.empty()) throw BackendError("mini-batch size = 0!");
  return weight;
}

void Sample::Start(Strategy strategy, SamplerStats* stats) {
  strategy_ = strategy;
  stats_ = stats;
  a_ = std::thread(&Sample::StageA, this);
  b_ = std::thread(&Sample::StageB, this);
  c_ = std::thread(&Sample::StageC, this);
}

namespace {
inline uint64_t NowNs() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}
}  // namespace

// stage A: the host strategy (sample.cc:177-302 of the reference), the only consumer of `seed`
void Sample::StageA() {
  std::unique_lock<std::mutex> lock(mu_);
  for (;;) {
    cv_.wait(lock, [this] { return stop_ || (drawn_ < allowed_ && drawn_ - consumed_ < kRing && !error_); });
    if (stop_) return;
    SampleSlot& slot = *ring[drawn_ % kRing];
    a_busy_ = true;
    lock.unlock();
    std::exception_ptr err;
    const uint64_t t0 = NowNs();
    try {
      if (dev_) {
        slot.weight = DrawOnDevice(&slot);
      } else {
        slot.edges.clear();
        slot.weight = strategy_(cfg_, &slot.edges, &seed);
      }
    } catch (...) {
      err = std::current_exception();
    }
    stats_->strategy += NowNs() - t0;
    lock.lock();
    a_busy_ = false;
    if (err) {
      error_ = err;
    } else {
      ++drawn_;
      if (dev_) ++extracted_;  // the device strategy delivers the nodes as well: no stage B
    }
    cv_.notify_all();
  }
}

// stage B: ExtractNodesFromMiniBatch (learner.cc:162-173, 177-178)
void Sample::StageB() {
  std::unique_lock<std::mutex> lock(mu_);
  for (;;) {
    cv_.wait(lock, [this] { return stop_ || (extracted_ < drawn_ && !error_); });
    if (stop_) return;
    SampleSlot& slot = *ring[extracted_ % kRing];
    lock.unlock();
    std::exception_ptr err;
    try {
      const uint64_t t0 = NowNs();
      ExtractNodesFromMiniBatch(slot.edges, &slot.nodes_vec);
      stats_->extract += NowNs() - t0;
      if (slot.nodes_vec.empty()) throw BackendError("mini-batch size = 0!");
    } catch (...) {
      err = std::current_exception();
    }
    lock.lock();
    if (err) error_ = err; else ++extracted_;
    cv_.notify_all();
  }
}

// stage C: Buffer::Write x2 + NeighborSampler (learner.cc:179-193)
void Sample::StageC() {
  std::unique_lock<std::mutex> lock(mu_);
  for (;;) {
    cv_.wait(lock, [this] { return stop_ || (ready_ < extracted_ && !error_); });
    if (stop_) return;
    SampleSlot& slot = *ring[ready_ % kRing];
    lock.unlock();
    std::exception_ptr err;
    try {
      const uint64_t t1 = NowNs();
      if (slot.edges.size() > slot.dev_edges.GetSize() / sizeof(Edge) ||
          slot.nodes_vec.size() > slot.dev_nodes.GetSize() / sizeof(Vertex))
        throw BackendError("mini-batch exceeds the device buffers");
      if (!dev_) {  // a device-drawn mini-batch is in the slot's device buffers already
        slot.dev_edges.Write(queue, slot.edges.size(), slot.edges.data());
        slot.dev_nodes.Write(queue, slot.nodes_vec.size(), slot.nodes_vec.data());
      }
      const uint64_t t2 = NowNs();
      neighbor_sampler(static_cast<uint32_t>(slot.nodes_vec.size()), &slot.dev_nodes, &slot.neighbors);
      const uint64_t t3 = NowNs();
      stats_->copy += t2 - t1;
      stats_->neighbors += t3 - t2;
      stats_->h2d_bytes += dev_ ? 16 : slot.edges.size() * sizeof(Edge) + slot.nodes_vec.size() * sizeof(Vertex);
    } catch (...) {
      err = std::current_exception();
    }
    lock.lock();
    if (err) error_ = err; else ++ready_;
    cv_.notify_all();
  }
}

void Sample::Allow(uint64_t more) {
  std::unique_lock<std::mutex> lock(mu_);
  if (consumed_ + more > allowed_) allowed_ = consumed_ + more;
  cv_.notify_all();
}

SampleSlot& Sample::WaitReady(uint64_t skip) {
  std::unique_lock<std::mutex> lock(mu_);
  const uint64_t want = consumed_ + skip;  // index of the mini-batch to hand out
  if (allowed_ < want + 1) {
    allowed_ = want + 1;
    cv_.notify_all();
  }
  cv_.wait(lock, [&] { return ready_ > want || error_; });
  if (ready_ <= want) {
    std::exception_ptr e = error_;
    error_ = nullptr;
    std::rethrow_exception(e);
  }
  return *ring[want % kRing];
}

void Sample::Release() {
  std::unique_lock<std::mutex> lock(mu_);
  ++consumed_;
  cv_.notify_all();
}

uint64_t Sample::Quiesce() {
  std::unique_lock<std::mutex> lock(mu_);
  // a draw that stage A has already started cannot be cancelled: let it land, start no other
  allowed_ = std::min(allowed_, drawn_ + (a_busy_ ? 1 : 0));
  cv_.wait(lock, [this] { return error_ || (!a_busy_ && ready_ == drawn_); });
  allowed_ = drawn_;
  return ready_ - consumed_;
}

SampleSlot& Sample::Latest() {
  std::unique_lock<std::mutex> lock(mu_);
  return *ring[(drawn_ + kRing - 1) % kRing];
}

// The reference's record for a Sample (sample.h:62-75): SampleStorage{edges, nodes_vec, seed},
// dev_edges, dev_nodes, then the NeighborSampler (RNG pool, neighbor data) -- of the most
// recently drawn mini-batch of this stream.  Call Quiesce() first.
bool Sample::Serialize(std::ostream* out) {
  SampleSlot& slot = Latest();
  SampleStorage s;
  s.edges.assign(reinterpret_cast<const char*>(slot.edges.data()), slot.edges.size() * sizeof(Edge));
  s.nodes_vec.assign(reinterpret_cast<const char*>(slot.nodes_vec.data()), slot.nodes_vec.size() * sizeof(Vertex));
  s.seed = seed;
  return SerializeMessage(out, s) && ::mcmc::Serialize(out, &slot.dev_edges, &queue) &&
         ::mcmc::Serialize(out, &slot.dev_nodes, &queue) && neighbor_sampler.Serialize(out);
}

bool Sample::Parse(std::istream* in, bool pending) {
  Quiesce();
  std::unique_lock<std::mutex> lock(mu_);
  // restart the counters: the parsed mini-batch lands in slot 0
  drawn_ = extracted_ = ready_ = 1;
  consumed_ = pending ? 0 : 1;
  allowed_ = drawn_;
  SampleSlot& slot = *ring[0];
  SampleStorage s;
  if (!ParseMessage(in, &s)) return false;
  // the mini-batch must fit the buffers this Sample was built with (file content is not trusted)
  if (s.edges.size() % sizeof(Edge) || s.nodes_vec.size() % sizeof(Vertex) ||
      s.edges.size() > slot.dev_edges.GetSize() || s.nodes_vec.size() > slot.dev_nodes.GetSize())
    return false;
  if (!(::mcmc::Parse(in, &slot.dev_edges, &queue) &&
        ::mcmc::Parse(in, &slot.dev_nodes, &queue) && neighbor_sampler.Parse(in)))
    return false;
  slot.edges.resize(s.edges.size() / sizeof(Edge));
  std::memcpy(slot.edges.data(), s.edges.data(), slot.edges.size() * sizeof(Edge));
  slot.nodes_vec.resize(s.nodes_vec.size() / sizeof(Vertex));
  std::memcpy(slot.nodes_vec.data(), s.nodes_vec.data(), slot.nodes_vec.size() * sizeof(Vertex));
  seed = s.seed;
  // neighbor data of the parsed mini-batch: the sampler's buffer holds it (sample.h:30-36)
  neighbor_sampler.GetData().CopyTo(queue, slot.nodes_vec.size() * neighbor_sampler.DataSizePerSample(),
                                    slot.neighbors);
  queue.Finish();
  return true;
}

}  // namespace mcmc
