// mcmc/perplexity.h -- PerplexityCalculator: averaged held-out log-likelihood.
// Call surface of the reference's perplexity.h:20-116.
#ifndef MCMC_B200_PERPLEXITY_H_
#define MCMC_B200_PERPLEXITY_H_

#include "mcmc/config.h"
#include "mcmc/partitioned-alloc.h"

namespace mcmc {

class PerplexityCalculator {
 public:
  enum Mode { EDGE_PER_THREAD, EDGE_PER_WORKGROUP };  // accepted; one kernel serves both

  PerplexityCalculator(Mode mode, const Config& cfg, clcuda::Queue queue, clcuda::Buffer<Float>& beta,
                       RowPartitionedMatrix<Float>* pi, clcuda::Buffer<Edge>& edges, OpenClSet* edgeSet,
                       const std::vector<std::string>& compileFlags = std::vector<std::string>(),
                       const std::string& baseFuncs = std::string());

  // -(sum of log-likelihoods)/(number of pairs); the Learner exponentiates
  Float operator()();

  uint64_t LinkCount() const { return static_cast<uint64_t>(sums_[2]); }
  uint64_t NonLinkCount() const { return static_cast<uint64_t>(sums_[3]); }
  double LinkLikelihood() const { return sums_[0]; }
  double NonLinkLikelihood() const { return sums_[1]; }
  double PerplexityTime() const { return t_ppx_; }
  double AccumulateTime() const { return 0; }  // the four reductions are fused into the kernel

  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);

 private:
  clcuda::Queue queue_;
  clcuda::Buffer<Float>& beta_;
  RowPartitionedMatrix<Float>* pi_;
  clcuda::Buffer<Edge>& edges_;
  OpenClSet* edgeSet_;
  clcuda::Buffer<Float> ppx_per_edge_;  // running mean per held-out pair
  clcuda::Buffer<char> workspace_;
  ammsb_params params_;
  uint32_t count_calls_;
  double sums_[4];
  double t_ppx_;
};

}  // namespace mcmc

#endif  // MCMC_B200_PERPLEXITY_H_
