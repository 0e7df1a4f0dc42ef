// main.cc -- command-line driver with the reference's flags (main.cc:43-81) and run loop
// (main.cc:160-170): load or generate the data set, build a Learner, alternate
// Run(ppx_interval) and HeldoutPerplexity().
#include <signal.h>

#include <cstring>
#include <iostream>
#include <map>
#include <sstream>

#include "mcmc/learner.h"
#include "mcmc/sharded_learner.h"

static sig_atomic_t signaled = 0;
static void handler(int) { signaled = 1; }

template <class T>
static bool Take(std::map<std::string, std::string>& args, const std::string& key, T* out) {
  auto it = args.find(key);
  if (it == args.end()) return false;
  std::istringstream in(it->second);
  in >> *out;
  args.erase(it);
  return true;
}

int main(int argc, char** argv) {
  std::map<std::string, std::string> args;
  static const std::map<std::string, std::string> kShort = {
      {"-f", "--file"}, {"-r", "--heldout-ratio"}, {"-a", "--a"}, {"-b", "--b"}, {"-c", "--c"},
      {"-e", "--epsilon"}, {"-k", "--k"}, {"-m", "--mini_batch"}, {"-n", "--neighbors"},
      {"-i", "--ppx-interval"}, {"-x", "--max-iters"}, {"-s", "--sample"}};
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-h" || a == "--help") {
      std::cout << "usage: ammsb-main --file GRAPH | --load-data 1 --load-file F [--dump-data 1 --dump-file F]\n"
                   "  -r heldout-ratio --alpha -a -b -c -e epsilon --eta0 --eta1 -k K -m mini_batch -n neighbors\n"
                   "  --ppx-wg --ppx-interval/-i --phi-wg --beta-wg --max-iters/-x --sample/-s STRATEGY\n"
                   "  --sampler-wg --phi-seed a,b --beta-seed a,b --neighbor-seed a,b --phi-mode MODE\n"
                   "  --phi-disable-noise 0|1 --phi-strict 0|1 --device D --device-sampler 0|1\n"
                   "  --devices 0-7 | 0,1,2,3   (2, 4 or 8 GPUs of one box, pi column-sharded)\n";
      return 1;
    }
    auto s = kShort.find(a);
    if (s != kShort.end()) a = s->second;
    std::string v;
    const size_t eq = a.find('=');
    if (eq != std::string::npos) {
      v = a.substr(eq + 1);
      a = a.substr(0, eq);
    } else if (i + 1 < argc) {
      v = argv[++i];
    }
    args[a] = v;
  }
  mcmc::Config cfg;
  // CLI defaults differ from the struct defaults (reference main.cc:50-70)
  cfg.alpha = 0;
  cfg.beta_seed = {44, 45};
  cfg.neighbor_seed = {56, 57};
  std::string filename, loadFile, dumpFile;
  uint32_t max_iters = 100;
  bool dumpDataset = false, loadDataset = false;
  int device = 0;
  std::string devices;  // "--devices 0,1,2,3" or "0-7": the iteration over several GPUs (column-sharded pi)
  Take(args, "--devices", &devices);
  Take(args, "--file", &filename);
  Take(args, "--heldout-ratio", &cfg.heldout_ratio);
  Take(args, "--alpha", &cfg.alpha);
  Take(args, "--a", &cfg.a);
  Take(args, "--b", &cfg.b);
  Take(args, "--c", &cfg.c);
  Take(args, "--epsilon", &cfg.epsilon);
  Take(args, "--eta0", &cfg.eta0);
  Take(args, "--eta1", &cfg.eta1);
  Take(args, "--k", &cfg.K);
  Take(args, "--mini_batch", &cfg.mini_batch_size);
  Take(args, "--neighbors", &cfg.num_node_sample);
  Take(args, "--ppx-wg", &cfg.ppx_wg_size);
  Take(args, "--ppx-interval", &cfg.ppx_interval);
  Take(args, "--phi-wg", &cfg.phi_wg_size);
  Take(args, "--beta-wg", &cfg.beta_wg_size);
  Take(args, "--max-iters", &max_iters);
  Take(args, "--sample", &cfg.strategy);
  Take(args, "--sampler-wg", &cfg.neighbor_sampler_wg_size);
  Take(args, "--phi-seed", &cfg.phi_seed);
  Take(args, "--beta-seed", &cfg.beta_seed);
  Take(args, "--neighbor-seed", &cfg.neighbor_seed);
  Take(args, "--phi-mode", &cfg.phi_mode);
  Take(args, "--phi-probs-shared", &cfg.phi_probs_shared);
  Take(args, "--phi-grads-shared", &cfg.phi_grads_shared);
  Take(args, "--phi-pi-shared", &cfg.phi_pi_shared);
  Take(args, "--phi-vwidth", &cfg.phi_vector_width);
  Take(args, "--beta-sum-grads-vwidth", &cfg.sum_grads_vector_width);
  Take(args, "--phi-disable-noise", &cfg.phi_disable_noise);
  Take(args, "--phi-strict", &cfg.phi_strict);
  Take(args, "--train-ppx", &cfg.calc_train_ppx);
  Take(args, "--train-ppx-ratio", &cfg.training_ppx_ratio);
  Take(args, "--stage-timers", &cfg.stage_timers);
  Take(args, "--device-sampler", &cfg.device_sampler);
  Take(args, "--dump-data", &dumpDataset);
  Take(args, "--dump-file", &dumpFile);
  Take(args, "--load-data", &loadDataset);
  Take(args, "--load-file", &loadFile);
  Take(args, "--device", &device);
  if (!args.empty()) {
    std::cerr << "unknown option " << args.begin()->first << std::endl;
    return 2;
  }
  if (!loadDataset && filename.empty()) { std::cerr << "--file is required" << std::endl; return 2; }
  if (loadDataset && loadFile.empty()) { std::cerr << "load-file is required with load-data" << std::endl; return 2; }
  if (dumpDataset && dumpFile.empty()) { std::cerr << "dump-file is required with dump-data" << std::endl; return 2; }

  std::vector<mcmc::Edge> unique_edges;
  if (!loadDataset) {
    if (!mcmc::GetUniqueEdgesFromFile(filename, &cfg.N, &unique_edges)) return 3;
    if (dumpDataset) return mcmc::DumpDataset(dumpFile, cfg.N, cfg.heldout_ratio, unique_edges) ? 0 : 3;
  } else if (!mcmc::LoadDataset(loadFile, &cfg.N, &cfg.heldout_ratio, &unique_edges)) {
    return 3;
  }
  if (!mcmc::GenerateSetsFromEdges(cfg.N, unique_edges, cfg.heldout_ratio, &cfg.training_edges,
                                   &cfg.heldout_edges, &cfg.training, &cfg.heldout)) {
    std::cerr << "Failed to generate training/heldout sets" << std::endl;
    return 3;
  }
  cfg.trainingGraph.reset(new mcmc::Graph(cfg.N, cfg.training_edges));
  cfg.heldoutGraph.reset(new mcmc::Graph(cfg.N, cfg.heldout_edges));
  if (cfg.alpha == 0) cfg.alpha = static_cast<mcmc::Float>(1) / cfg.K;
  cfg.E = unique_edges.size();
  std::cerr << "Loaded " << (loadDataset ? loadFile : filename)
            << " (training max fan out = " << cfg.trainingGraph->MaxFanOut()
            << ", heldout max fan out = " << cfg.heldoutGraph->MaxFanOut() << ")\n" << cfg;
  signal(SIGINT, handler);
  try {
    if (!devices.empty()) {
      std::vector<int> list;
      const size_t dash = devices.find('-');
      if (dash != std::string::npos) {
        for (int d = std::stoi(devices.substr(0, dash)); d <= std::stoi(devices.substr(dash + 1)); ++d) list.push_back(d);
      } else {
        std::stringstream ss(devices);
        for (std::string tok; std::getline(ss, tok, ',');) list.push_back(std::stoi(tok));
      }
      mcmc::ShardedLearner learner(cfg, list);
      std::cerr << "Devices: " << list.size() << " ranks, pi column-sharded" << std::endl;
      std::cerr << "ppx[0] = " << learner.HeldoutPerplexity() << std::endl;
      for (uint64_t i = 0; i < max_iters && !signaled; i += cfg.ppx_interval) {
        const uint64_t step = std::min<uint64_t>(max_iters - i, cfg.ppx_interval);
        learner.Run(step, &signaled);
        if (!signaled) std::cerr << "ppx[" << i + step << "] = " << learner.HeldoutPerplexity() << std::endl;
      }
      if (signaled) std::cerr << "FORCED TERMINATE" << std::endl;
      learner.PrintStats();
      return 0;
    }
    mcmc::clcuda::Device dev(device);
    mcmc::clcuda::Context context(dev);
    mcmc::clcuda::Queue queue(context, dev);
    std::cerr << "Device: " << dev.Name() << " (" << dev.Version() << ")" << std::endl;
    mcmc::Learner learner(cfg, queue);
    std::cerr << "ppx[0] = " << learner.HeldoutPerplexity() << std::endl;
    for (uint64_t i = 0; i < max_iters && !signaled; i += cfg.ppx_interval) {
      const uint64_t step = std::min<uint64_t>(max_iters - i, cfg.ppx_interval);
      learner.Run(step, &signaled);
      if (!signaled) std::cerr << "ppx[" << i + step << "] = " << learner.HeldoutPerplexity() << std::endl;
    }
    if (signaled) std::cerr << "FORCED TERMINATE" << std::endl;
    learner.PrintStats();
  } catch (const std::exception& e) {
    std::cerr << "FATAL: " << e.what() << std::endl;
    return 4;
  }
  return 0;
}
