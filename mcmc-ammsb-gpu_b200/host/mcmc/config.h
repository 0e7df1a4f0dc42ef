// mcmc/config.h -- run configuration of the SG-MCMC learner.
// Field names, types and defaults are those of the reference's struct Config
// (mcmc/config.h:25-102) so that code written against it keeps compiling.  Where the
// reference bakes hyper-parameters into JIT compile flags (MakeCompileFlags,
// config.cc:66-83), MakeParams() produces the ammsb_params block the sm_100a kernels
// take as an argument -- with the same "%e" text rounding of every Float.
#ifndef MCMC_B200_CONFIG_H_
#define MCMC_B200_CONFIG_H_

#include <memory>
#include <string>
#include <vector>

#include "mcmc/data.h"
#include "mcmc/sample.h"

namespace mcmc {

enum PhiUpdaterMode {
  PHI_NODE_PER_THREAD,
  PHI_NODE_PER_WORKGROUP_NAIVE,
  PHI_NODE_PER_WORKGROUP_SHARED,
  PHI_NODE_PER_WORKGROUP_CODE_GEN
};
std::istream& operator>>(std::istream& in, PhiUpdaterMode& mode);
std::string to_string(const PhiUpdaterMode& mode);

struct Config {
  // model
  // the reference's MCMC_CALC_TRAIN_PPX build option (config.h:26-28) as a run-time switch:
  // Learner::TrainingPerplexity() over a sampled subset of the training edges
  bool calc_train_ppx = false;
  Float training_ppx_ratio = 0.01;
  Float heldout_ratio = 0.01;
  Float alpha = 0.001;
  Float a = 0.0315, b = 1024, c = 0.5;  // step size eps_t = a (1 + t/b)^-c
  Float epsilon = 1e-7;
  Float eta0 = 1, eta1 = 1;
  uint64_t K = 32;
  uint64_t mini_batch_size = 32;
  uint64_t num_node_sample = 32;
  uint64_t N = 0;
  uint64_t E = 0;
  // data (owned by the caller's Config; the Learner keeps a reference to it)
  std::vector<Edge> training_edges;
  std::vector<Edge> heldout_edges;
  std::unique_ptr<mcmc::Set> training;
  std::unique_ptr<mcmc::Set> heldout;
  std::unique_ptr<mcmc::Graph> trainingGraph;
  std::unique_ptr<mcmc::Graph> heldoutGraph;
  // launch geometry of the reference; here they select which reference launch's RNG-state
  // mapping and summation order is reproduced (they do not size CUDA blocks)
  uint32_t ppx_wg_size = 32;
  uint32_t ppx_interval = 100;
  uint32_t neighbor_sampler_wg_size = 32;
  uint32_t phi_wg_size = 32;
  uint32_t beta_wg_size = 32;
  // seeds
  ulong2 phi_seed = {42, 43};
  ulong2 beta_seed = {113, 117};
  ulong2 neighbor_seed = {3337, 54351};
  bool phi_disable_noise = false;
  SampleStrategy strategy = Node;
  PhiUpdaterMode phi_mode = PHI_NODE_PER_WORKGROUP_NAIVE;
  // accepted for compatibility; placement/vectorisation are fixed by the sm_100a kernels
  bool phi_probs_shared = true;
  bool phi_grads_shared = true;
  bool phi_pi_shared = true;
  uint32_t phi_vector_width = 1;
  uint32_t sum_grads_vector_width = 1;
  // B200 additions
  bool phi_strict = false;     // run the IEEE reference-association kernel instead of the fast one
  bool stage_timers = false;   // per-kernel timing with a sync after every stage (reference behaviour)
  // draw the Node / NodeLink / NodeNonLink mini-batches on the device (same mini-batches, element
  // for element; takes the host strategy and the H2D copies out of the iteration)
  bool device_sampler = false;
};

std::ostream& operator<<(std::ostream& out, const Config& cfg);

// "-DNAME=value" list exactly as the reference would pass to its JIT (kept for logging
// and for callers that inspect it)
std::vector<std::string> MakeCompileFlags(const Config& cfg);
// the same information as a kernel-argument block
ammsb_params MakeParams(const Config& cfg);
ammsb_phi_opts MakePhiOpts(const Config& cfg);

const std::string& GetSourceGuard();

}  // namespace mcmc

#endif  // MCMC_B200_CONFIG_H_
