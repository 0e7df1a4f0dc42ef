// A caller written against the REFERENCE's public C++ surface only -- the calls main.cc:17-172 and
// serialize-test.cc:119-133 make (mcmc/learner.h:18-36, config.h:25-102, data.h:16-52, types.h:31-74,
// sample.h:94-101, the clcuda handles of types.h:29) -- compiled and linked against libmcmc.so by
// tests/test_dropin_api.py.  If a name, a field or a signature drifts from the reference's, this
// stops compiling.  (It is only run when a GPU is present.)
#include <csignal>
#include <fstream>
#include <iostream>
#include <sstream>

#include "mcmc/learner.h"

static sig_atomic_t signaled = 0;
static void handler(int) { signaled = 1; }

int main(int argc, char** argv) {
  mcmc::Config cfg;
  // every field main.cc binds an option to (main.cc:43-81)
  cfg.heldout_ratio = 0.01;
  cfg.alpha = 0;
  cfg.a = 0.0315;
  cfg.b = 1024;
  cfg.c = 0.5;
  cfg.epsilon = 1e-7;
  cfg.eta0 = 1;
  cfg.eta1 = 1;
  cfg.K = 32;
  cfg.mini_batch_size = 32;
  cfg.num_node_sample = 32;
  cfg.ppx_wg_size = 32;
  cfg.ppx_interval = 100;
  cfg.phi_wg_size = 32;
  cfg.beta_wg_size = 32;
  cfg.neighbor_sampler_wg_size = 32;
  cfg.phi_seed = {42, 43};
  cfg.beta_seed = {44, 45};
  cfg.neighbor_seed = {56, 57};
  cfg.phi_disable_noise = false;
  cfg.phi_probs_shared = cfg.phi_grads_shared = cfg.phi_pi_shared = true;
  cfg.phi_vector_width = 1;
  cfg.sum_grads_vector_width = 1;
  std::istringstream("Node") >> cfg.strategy;
  std::istringstream("WG-NAIVE") >> cfg.phi_mode;
  const mcmc::SampleStrategy all[] = {mcmc::Node, mcmc::NodeLink, mcmc::NodeNonLink, mcmc::BFLink, mcmc::BFNonLink, mcmc::BF};
  const mcmc::PhiUpdaterMode modes[] = {mcmc::PHI_NODE_PER_THREAD, mcmc::PHI_NODE_PER_WORKGROUP_NAIVE,
                                        mcmc::PHI_NODE_PER_WORKGROUP_SHARED, mcmc::PHI_NODE_PER_WORKGROUP_CODE_GEN};
  (void)all;
  (void)modes;
  if (argc < 2) {
    std::cerr << "usage: " << argv[0] << " snap-file [checkpoint]" << std::endl;
    return 2;
  }
  // main.cc:17-20,99-101
  mcmc::clcuda::Platform platform((size_t)0);
  mcmc::clcuda::Device dev(platform, 0);
  mcmc::clcuda::Context context(dev);
  mcmc::clcuda::Queue queue(context, dev);
  std::cerr << dev.Name() << " " << dev.Vendor() << " " << dev.Type() << " " << dev.Version() << std::endl;
  // main.cc:101-154
  std::vector<mcmc::Edge> unique_edges;
  if (!mcmc::GetUniqueEdgesFromFile(argv[1], &cfg.N, &unique_edges) ||
      !mcmc::GenerateSetsFromEdges(cfg.N, unique_edges, cfg.heldout_ratio, &cfg.training_edges, &cfg.heldout_edges,
                                   &cfg.training, &cfg.heldout)) {
    std::cerr << "Failed to generate sets from file" << std::endl;
    return 1;
  }
  cfg.trainingGraph.reset(new mcmc::Graph(cfg.N, cfg.training_edges));
  cfg.heldoutGraph.reset(new mcmc::Graph(cfg.N, cfg.heldout_edges));
  if (cfg.alpha == 0) cfg.alpha = static_cast<mcmc::Float>(1) / cfg.K;
  cfg.E = unique_edges.size();
  std::cerr << "max fan out " << cfg.trainingGraph->MaxFanOut() << " / " << cfg.heldoutGraph->MaxFanOut() << "\n" << cfg;
  mcmc::Vertex u, v;
  std::tie(u, v) = mcmc::Vertices(unique_edges[0]);
  if (mcmc::MakeEdge(u, v) != unique_edges[0] || !(cfg.training->Has(unique_edges[0]) || cfg.heldout->Has(unique_edges[0])))
    return 1;
  for (const std::string& f : mcmc::MakeCompileFlags(cfg)) std::cerr << f << " ";
  std::cerr << std::endl;
  signal(SIGINT, handler);
  // main.cc:160-170
  mcmc::Learner learner(cfg, queue);
  std::cout << "ppx[0] = " << learner.HeldoutPerplexity() << std::endl;
  for (uint32_t i = 0; i < 20 && !signaled; i += 10) {
    learner.Run(10, &signaled);
    std::cout << "ppx[" << i + 10 << "] = " << learner.HeldoutPerplexity() << std::endl;
  }
  learner.PrintStats();
  // serialize-test.cc:119-133
  if (argc > 2) {
    {
      std::ofstream out(argv[2], std::ofstream::binary);
      if (!learner.Serialize(&out)) return 1;
    }
    mcmc::Learner resumed(cfg, queue);
    std::ifstream in(argv[2], std::ifstream::binary);
    if (!resumed.Parse(&in)) return 1;
    learner.Run(10);
    resumed.Run(10);
    const mcmc::Float a = learner.HeldoutPerplexity(), b = resumed.HeldoutPerplexity();
    std::cout << "resumed " << b << " vs " << a << std::endl;
    if (a != b) return 3;  // ASSERT_EQ(ppx, ppx2)
  }
  return 0;
}
