#include "mcmc/sample.h"

#include <algorithm>
#include <cctype>
#include <queue>
#include <unordered_set>

#include "mcmc/config.h"
#include "mcmc/serialize.h"

namespace mcmc {

namespace {
inline Edge Canonical(Vertex u, Vertex v) { return MakeEdge(std::min(u, v), std::max(u, v)); }
inline Vertex DrawVertex(const Config& cfg, unsigned int* seed) { return rand_r(seed) % cfg.N; }
inline void Emit(const std::unordered_set<Edge>& picked, std::vector<Edge>* edges) {
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
  std::unordered_set<Vertex> tried;
  std::unordered_set<Edge> picked;
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
Float sampleNodeNonLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  std::unordered_set<Edge> picked;
  const Vertex u = DrawVertex(cfg, seed);
  while (picked.size() < cfg.mini_batch_size) {
    Edge e;
    do {
      e = Canonical(u, DrawVertex(cfg, seed));
    } while (cfg.heldout->Has(e) || cfg.training->Has(e));
    picked.insert(e);
  }
  Emit(picked, edges);
  return (2 * cfg.E) / static_cast<Float>(cfg.mini_batch_size);
}

Float sampleNode(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  return (rand_r(seed) % 2) ? sampleNodeLink(cfg, edges, seed) : sampleNodeNonLink(cfg, edges, seed);
}

// Breadth-first over training links from random roots until m edges.  Scale E/m.
Float sampleBreadthFirstLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed) {
  std::unordered_set<Vertex> visited;
  std::queue<Vertex> frontier;
  std::unordered_set<Edge> picked;
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
  std::unordered_set<Vertex> visited;
  std::queue<Vertex> frontier;
  std::unordered_set<Edge> picked;
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
  std::unordered_set<Vertex> nodes;
  for (Edge e : edges) {
    nodes.insert(static_cast<Vertex>(e >> 32));
    nodes.insert(static_cast<Vertex>(e & 0xffffffffu));
  }
  nodes_vec->clear();
  nodes_vec->insert(nodes_vec->begin(), nodes.begin(), nodes.end());
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

Sample::Sample(const Config& cfg, clcuda::Queue q)
    : queue(q.GetContext(), q.GetDevice()),
      dev_edges(q.GetContext(), MaxMiniBatchEdges(cfg)),
      dev_nodes(q.GetContext(), MaxMiniBatchNodes(cfg)),
      seed(rand()),
      neighbor_sampler(cfg, clcuda::Queue(q.GetContext(), q.GetDevice())) {}

bool Sample::Serialize(std::ostream* out) {
  SampleStorage s;
  s.edges.assign(reinterpret_cast<const char*>(edges.data()), edges.size() * sizeof(Edge));
  s.nodes_vec.assign(reinterpret_cast<const char*>(nodes_vec.data()), nodes_vec.size() * sizeof(Vertex));
  s.seed = seed;
  return SerializeMessage(out, s) && ::mcmc::Serialize(out, &dev_edges, &queue) &&
         ::mcmc::Serialize(out, &dev_nodes, &queue) && neighbor_sampler.Serialize(out);
}

bool Sample::Parse(std::istream* in) {
  SampleStorage s;
  if (!(ParseMessage(in, &s) && ::mcmc::Parse(in, &dev_edges, &queue) && ::mcmc::Parse(in, &dev_nodes, &queue) &&
        neighbor_sampler.Parse(in)))
    return false;
  edges.resize(s.edges.size() / sizeof(Edge));
  std::memcpy(edges.data(), s.edges.data(), edges.size() * sizeof(Edge));
  nodes_vec.resize(s.nodes_vec.size() / sizeof(Vertex));
  std::memcpy(nodes_vec.data(), s.nodes_vec.data(), nodes_vec.size() * sizeof(Vertex));
  seed = s.seed;
  return true;
}

}  // namespace mcmc
