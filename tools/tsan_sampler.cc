// Host mini-batch strategies from several threads at once, for ThreadSanitizer (development tool;
// tools/tsan_host.sh builds and runs it).  What the Learner's sampler threads share is read-only
// (Config, the two cuckoo sets, the graph) except the lazily built partner indexes of the sets;
// everything else (ordered-set scratch) is thread-local.  Each thread's output must equal what a
// single thread produces for the same seed.
#include <cstdio>
#include <random>
#include <thread>
#include <unordered_set>

#include "mcmc/config.h"
#include "mcmc/sample.h"

using namespace mcmc;

int main() {
  Config cfg;
  cfg.mini_batch_size = 2048;
  cfg.heldout_ratio = 0.1;
  const uint64_t N = 20000, E = 120000;
  std::mt19937_64 g(1);
  std::unordered_set<Edge> seen;
  std::vector<Edge> edges;
  while (edges.size() < E) {
    const Vertex u = g() % N, v = g() % N;
    if (u == v) continue;
    const Edge e = MakeEdge(std::min(u, v), std::max(u, v));
    if (seen.insert(e).second) edges.push_back(e);
  }
  cfg.N = N;
  cfg.E = E;
  srand(1);
  if (!GenerateSetsFromEdges(N, edges, cfg.heldout_ratio, &cfg.training_edges, &cfg.heldout_edges, &cfg.training,
                             &cfg.heldout))
    return 2;
  cfg.trainingGraph.reset(new Graph(N, cfg.training_edges));
  const int kThreads = 6, kRounds = 60;
  auto run = [&](unsigned seed, uint64_t* digest) {
    std::vector<Edge> mb;
    std::vector<Vertex> nodes;
    uint64_t h = 1469598103934665603ull;
    for (int r = 0; r < kRounds; ++r) {
      mb.clear();
      sampleNode(cfg, &mb, &seed);
      ExtractNodesFromMiniBatch(mb, &nodes);
      for (Edge e : mb) h = (h ^ e) * 1099511628211ull;
      for (Vertex v : nodes) h = (h ^ v) * 1099511628211ull;
    }
    *digest = h ^ seed;
  };
  uint64_t parallel[kThreads], serial[kThreads];
  std::vector<std::thread> threads;
  for (int t = 0; t < kThreads; ++t) threads.emplace_back(run, 100u + t, &parallel[t]);  // first calls race to build the indexes
  for (auto& th : threads) th.join();
  for (int t = 0; t < kThreads; ++t) run(100u + t, &serial[t]);
  for (int t = 0; t < kThreads; ++t)
    if (parallel[t] != serial[t]) {
      std::printf("thread %d: output differs from the single-threaded run\n", t);
      return 1;
    }
  std::printf("tsan sampler: %d threads x %d mini-batches, outputs equal the serial run\n", kThreads, kRounds);
  return 0;
}
