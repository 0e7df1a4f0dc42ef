// mcmc/beta.h -- BetaUpdater: theta gradient over the mini-batch edges, Langevin step on
// theta, beta = row-normalised theta.  Call surface of the reference's beta.h:17-73.
#ifndef MCMC_B200_BETA_H_
#define MCMC_B200_BETA_H_

#include "mcmc/config.h"
#include "mcmc/partitioned-alloc.h"
#include "mcmc/random.h"

namespace mcmc {

class BetaUpdater {
 public:
  enum Mode { EDGE_PER_THREAD, EDGE_PER_WORKGROUP };  // accepted; one kernel serves both

  BetaUpdater(Mode mode, const Config& cfg, clcuda::Queue queue, clcuda::Buffer<Float>& theta,
              clcuda::Buffer<Float>& beta, RowPartitionedMatrix<Float>* pi, OpenClSet* trainingSet,
              const std::vector<std::string>& compileFlags = std::vector<std::string>(),
              const std::string& baseFuncs = std::string());

  void operator()(clcuda::Buffer<Edge>* edges, uint32_t num_edges, Float scale);

  clcuda::Buffer<Float>& GetThetaSum() { return theta_sum_; }
  clcuda::Buffer<Float>& GetGrads() { return grads_; }
  random::OpenClRandom* GetRandom() { return rand_.get(); }

  double ThetaSumTime() const { return 0; }  // fused into the gradient kernels
  double GradsPartialTime() const { return t_grads_; }
  double GradsSumTime() const { return 0; }
  double UpdateThetaTime() const { return t_update_theta_; }
  double NormalizeTime() const { return 0; }

  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);

 private:
  const Config& cfg_;
  clcuda::Queue queue_;
  clcuda::Buffer<Float>& theta_;  // [K,2]
  clcuda::Buffer<Float>& beta_;   // [K,2]
  RowPartitionedMatrix<Float>* pi_;
  OpenClSet* trainingSet_;
  std::shared_ptr<random::OpenClRandomFactory> randFactory_;
  std::unique_ptr<random::OpenClRandom> rand_;
  ammsb_params params_;
  uint32_t count_calls_;
  clcuda::Buffer<Float> theta_sum_;  // [K]
  clcuda::Buffer<Float> grads_;      // [K,2]
  clcuda::Buffer<char> workspace_;
  double t_grads_, t_update_theta_;
};

}  // namespace mcmc

#endif  // MCMC_B200_BETA_H_
