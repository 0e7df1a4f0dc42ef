// Host mini-batch sampler cost breakdown (development tool): g++ -O2 -std=c++17 -I mcmc-ammsb-gpu_b200/host -I include tools/sampler_breakdown.cc -L mcmc-ammsb-gpu_b200 -lmcmc -lammsb
#include <chrono>
#include <random>
#include <algorithm>
#include <unordered_set>
#include "mcmc/config.h"
#include "mcmc/sample.h"
#include "mcmc/std_order_set.h"
using namespace mcmc;
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a, clk::time_point b){return std::chrono::duration<double,std::milli>(b-a).count();}
int main() {
  Config cfg; cfg.K=1024; cfg.mini_batch_size=16384; cfg.num_node_sample=32; cfg.heldout_ratio=0.1;
  uint64_t N=317080, E=1049866; std::mt19937_64 g(1);
  std::unordered_set<Edge> es; std::vector<Edge> edges;
  while (es.size()<E){ Vertex u=g()%N, v=g()%N; if(u==v) continue; Edge e=MakeEdge(std::min(u,v),std::max(u,v)); if(es.insert(e).second) edges.push_back(e);}
  cfg.N=N; srand(1);
  GenerateSetsFromEdges(N, edges, cfg.heldout_ratio, &cfg.training_edges, &cfg.heldout_edges, &cfg.training, &cfg.heldout);
  cfg.trainingGraph.reset(new Graph(N, cfg.training_edges)); cfg.E=E;
  const int R=50, M=16384;
  std::vector<Edge> cand(M); unsigned seed=1; double t_rand=0,t_loc=0,t_has=0,t_ins=0,t_std=0,t_emit=0,t_ext=0;
  StdOrderSet<Edge> fs; std::vector<Edge> out; std::vector<Vertex> nodes; size_t hb[2],tb[2]; uint64_t sink=0;
  for(int r=0;r<R;++r){
    Vertex u = rand_r(&seed)%N;
    auto t0=clk::now();
    for(int i=0;i<M;++i){ Vertex v=rand_r(&seed)%N; cand[i]=MakeEdge(std::min(u,v),std::max(u,v)); }
    auto t1=clk::now();
    for(int i=0;i<M;++i){ cfg.heldout->Locate(cand[i],hb); cfg.training->Locate(cand[i],tb); sink+=hb[0]+tb[1]; }
    auto t2=clk::now();
    for(int i=0;i<M;++i){ sink += cfg.heldout->Has(cand[i]) || cfg.training->Has(cand[i]); }
    auto t3=clk::now();
    fs.Clear(); for(int i=0;i<M;++i) fs.Insert(cand[i]);
    auto t4=clk::now();
    { std::unordered_set<Edge> s; for(int i=0;i<M;++i) s.insert(cand[i]); sink+=s.size(); }
    auto t5=clk::now();
    out.clear(); fs.EmitTo(&out);
    auto t6=clk::now();
    ExtractNodesFromMiniBatch(out,&nodes);
    auto t7=clk::now();
    t_rand+=ms(t0,t1); t_loc+=ms(t1,t2); t_has+=ms(t2,t3); t_ins+=ms(t3,t4); t_std+=ms(t4,t5); t_emit+=ms(t5,t6); t_ext+=ms(t6,t7);
  }
  printf("rand %.3f  locate(4 hashes+prefetch) %.3f  has(again, cached) %.3f  flat insert %.3f  std insert %.3f  emit %.3f  extract %.3f ms  [%lu]\n",
    t_rand/R,t_loc/R,t_has/R,t_ins/R,t_std/R,t_emit/R,t_ext/R,(unsigned long)sink);
}
