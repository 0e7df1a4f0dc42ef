// mcmc/phi.h -- PhiUpdater: update_phi + update_pi for one mini-batch.
// Same call surface as the reference (phi.h:12-62): construct once, then
// operator()(mini_batch_nodes, neighbors, num_nodes) per iteration.
#ifndef MCMC_B200_PHI_H_
#define MCMC_B200_PHI_H_

#include "mcmc/config.h"
#include "mcmc/partitioned-alloc.h"
#include "mcmc/random.h"

namespace mcmc {

class PhiUpdater {
 public:
  typedef PhiUpdaterMode Mode;

  // compileFlags/baseFuncs are what the reference feeds its JIT; accepted and ignored.
  PhiUpdater(const Config& cfg, clcuda::Queue queue, clcuda::Buffer<Float>& beta, RowPartitionedMatrix<Float>* pi,
             clcuda::Buffer<Float>& phi, OpenClSet* trainingSet,
             const std::vector<std::string>& compileFlags = std::vector<std::string>(),
             const std::string& baseFuncs = std::string());

  void operator()(clcuda::Buffer<Vertex>& mini_batch_nodes,  // [<= max(2m, 1+MaxFanOut)]
                  clcuda::Buffer<Vertex>& neighbors,         // [nodes, num_node_sample]
                  uint32_t num_mini_batch_nodes);

  double UpdatePhiTime() const { return t_update_phi_; }
  double UpdatePiTime() const { return t_update_pi_; }
  clcuda::Buffer<Float>& GetPhiVec() { return phi_vec_; }
  random::OpenClRandom* GetRandom() { return rand_.get(); }

  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);

 private:
  const Config& cfg_;
  clcuda::Queue queue_;
  clcuda::Buffer<Float>& beta_;      // [K,2]
  RowPartitionedMatrix<Float>* pi_;  // [N,K] (+ phi row sums inside the store)
  clcuda::Buffer<Float>& phi_;       // [N]   kept in step with the store for API compatibility
  clcuda::Buffer<Float> phi_vec_;    // [max nodes, K]
  clcuda::Buffer<Float> phi_sum_;    // [max nodes]
  OpenClSet* trainingSet_;
  std::shared_ptr<random::OpenClRandomFactory> randFactory_;
  std::unique_ptr<random::OpenClRandom> rand_;
  ammsb_params params_;
  ammsb_phi_opts opts_;
  uint32_t count_calls_;
  double t_update_phi_, t_update_pi_;
};

}  // namespace mcmc

#endif  // MCMC_B200_PHI_H_
