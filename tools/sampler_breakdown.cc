// Host mini-batch sampler cost (development tool):
//   g++ -O2 -std=c++17 -I mcmc-ammsb-gpu_b200/host -I include tools/sampler_breakdown.cc \
//       -L mcmc-ammsb-gpu_b200 -lmcmc -lammsb -Wl,-rpath,$PWD/mcmc-ammsb-gpu_b200 -o /tmp/sb
// DBLP-shaped graph, m = 16384: time of one non-link mini-batch (strategy, node extraction) and
// of the pieces a strategy is made of.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <random>
#include <unordered_set>

#include "mcmc/config.h"
#include "mcmc/sample.h"
#include "mcmc/std_order_set.h"
using namespace mcmc;
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }

int main(int argc, char** argv) {
  Config cfg;
  cfg.K = 1024;
  cfg.mini_batch_size = argc > 1 ? atoi(argv[1]) : 16384;
  cfg.num_node_sample = 32;
  cfg.heldout_ratio = 0.1;
  const uint64_t N = 317080, E = 1049866;
  std::mt19937_64 g(1);
  std::unordered_set<Edge> es;
  std::vector<Edge> edges;
  while (es.size() < E) {
    const Vertex u = g() % N, v = g() % N;
    if (u == v) continue;
    const Edge e = MakeEdge(std::min(u, v), std::max(u, v));
    if (es.insert(e).second) edges.push_back(e);
  }
  cfg.N = N;
  srand(1);
  GenerateSetsFromEdges(N, edges, cfg.heldout_ratio, &cfg.training_edges, &cfg.heldout_edges, &cfg.training, &cfg.heldout);
  cfg.trainingGraph.reset(new Graph(N, cfg.training_edges));
  cfg.E = E;
  const int R = 200, M = static_cast<int>(cfg.mini_batch_size);
  std::vector<Edge> cand(M), out;
  std::vector<Vertex> nodes;
  unsigned seed = 1;
  double t_strategy = 0, t_extract = 0, t_rand = 0, t_loc = 0, t_ins = 0, t_emit = 0, t_std = 0;
  uint64_t sink = 0;
  sampleNodeNonLink(cfg, &out, &seed);  // builds the partner indexes
  for (int r = 0; r < R; ++r) {
    out.clear();
    const auto t0 = clk::now();
    sampleNodeNonLink(cfg, &out, &seed);
    const auto t1 = clk::now();
    ExtractNodesFromMiniBatch(out, &nodes);
    const auto t2 = clk::now();
    t_strategy += ms(t0, t1);
    t_extract += ms(t1, t2);
    sink += out[3] + nodes[5];
  }
  StdOrderSet<Edge> fs;
  size_t hb[2], tb[2];
  for (int r = 0; r < R / 4; ++r) {
    const Vertex u = rand_r(&seed) % N;
    const auto t0 = clk::now();
    for (int i = 0; i < M; ++i) cand[i] = MakeEdge(std::min<Vertex>(u, rand_r(&seed) % N), std::max<Vertex>(u, rand_r(&seed) % N));
    const auto t1 = clk::now();
    for (int i = 0; i < M; ++i) {
      cfg.heldout->Locate(cand[i], hb);
      cfg.training->Locate(cand[i], tb);
      sink += hb[0] + tb[1];
    }
    const auto t2 = clk::now();
    fs.Clear();
    for (int i = 0; i < M; ++i) fs.Insert(cand[i]);
    const auto t3 = clk::now();
    out.clear();
    fs.EmitTo(&out);
    const auto t4 = clk::now();
    {
      std::unordered_set<Edge> s;
      for (int i = 0; i < M; ++i) s.insert(cand[i]);
      sink += s.size();
    }
    const auto t5 = clk::now();
    t_rand += ms(t0, t1) / 2, t_loc += ms(t1, t2), t_ins += ms(t2, t3), t_emit += ms(t3, t4), t_std += ms(t4, t5);
  }
  printf("m=%d  sampleNodeNonLink %.3f ms  ExtractNodesFromMiniBatch %.3f ms\n", M, t_strategy / R, t_extract / R);
  printf("pieces: rand_r %.3f  4 cuckoo bins hashed+prefetched %.3f  StdOrderSet insert %.3f + emit %.3f  (std::unordered_set insert %.3f) ms  [%lu]\n",
         t_rand / (R / 4), t_loc / (R / 4), t_ins / (R / 4), t_emit / (R / 4), t_std / (R / 4), (unsigned long)sink);
}
