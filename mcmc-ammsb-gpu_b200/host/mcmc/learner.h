// mcmc/learner.h -- the SG-MCMC learner for the a-MMSB.
// Drop-in for the reference's mcmc::Learner (learner.h:18-88): same constructor, Run,
// HeldoutPerplexity, PrintStats, Serialize, Parse.  Owns all device state; one
// iteration = (async, double-buffered) mini-batch + neighbor sampling -> update_phi ->
// update_pi -> update_beta/theta, all on sm_100a kernels through the C ABI.
#ifndef MCMC_B200_LEARNER_H_
#define MCMC_B200_LEARNER_H_

#include <signal.h>

#include <atomic>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <ostream>

#include "mcmc/beta.h"
#include "mcmc/config.h"
#include "mcmc/data.h"
#include "mcmc/perplexity.h"
#include "mcmc/phi.h"

namespace mcmc {

class Learner {
 public:
  // cfg is held by reference (as in the reference): it must outlive the Learner
  Learner(const Config& cfg, clcuda::Queue queue);
  ~Learner();

  void Run(uint32_t max_iters, sig_atomic_t* signaled = nullptr);
  Float HeldoutPerplexity();
  // exp(-mean log-likelihood) over cfg.training_ppx_ratio of the training links plus the
  // matching number of random non-links (reference learner.cc:47-75,204-212); needs
  // cfg.calc_train_ppx
  Float TrainingPerplexity();
  const std::vector<Edge>& TrainingPerplexityEdges() const { return trainingPerplexityEdges_; }
  void PrintStats();
  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);

  static const std::string GetBaseFuncs() { return std::string(); }  // no JIT prelude

  // state access for tests / tools (not in the reference API)
  void ReadPi(uint64_t row0, uint64_t nrows, Float* host) { pi_->ReadRows(row0, nrows, host); }
  void ReadPhi(Float* host) { phi_.Read(queue_, cfg_.N, host); }
  void ReadBeta(Float* host) { beta_.Read(queue_, 2 * cfg_.K, host); }
  void ReadTheta(Float* host) { theta_.Read(queue_, 2 * cfg_.K, host); }
  uint32_t StepCount() const { return stepCount_; }
  uint64_t EdgesProcessed() const { return edgesProcessed_; }
  uint64_t BytesH2D() const { return stats_.h2d_bytes; }  // mini-batch edges + nodes copied to the device so far
  // when set (pinned host memory, 2K floats), every iteration ends with a device->host copy
  // of beta into it: the per-iteration result a monitoring caller reads
  void MirrorBetaTo(Float* pinned_host) { betaMirror_ = pinned_host; }
  // the mini-batch the next iteration will consume (joins the sampler thread)
  const SampleSlot& PeekNextSample();
  Float PeekNextWeight() { return PeekNextSample().weight; }
  Sample& SampleStream(int i) { return *samples_[i]; }

 private:
  Float SampleMiniBatch(std::vector<Edge>* edges, unsigned int* seed);

  const Config& cfg_;
  clcuda::Queue queue_;
  clcuda::Buffer<Float> beta_;   // [K,2]
  clcuda::Buffer<Float> theta_;  // [K,2]
  std::shared_ptr<RowPartitionedMatrixFactory<Float>> allocFactory_;
  std::unique_ptr<RowPartitionedMatrix<Float>> pi_;  // [N,K]
  clcuda::Buffer<Float> phi_;                        // [N]
  std::shared_ptr<OpenClSetFactory> setFactory_;
  std::unique_ptr<OpenClSet> trainingSet_;
  std::unique_ptr<OpenClSet> heldoutSet_;
  clcuda::Buffer<Edge> trainingEdges_;
  clcuda::Buffer<Edge> heldoutEdges_;
  std::vector<std::string> compileFlags_;
  std::vector<Edge> trainingPerplexityEdges_;
  std::unique_ptr<clcuda::Buffer<Edge>> devTrainingPerplexityEdges_;
  std::unique_ptr<PerplexityCalculator> trainingPerplexity_;
  PerplexityCalculator heldoutPerplexity_;
  PhiUpdater phiUpdater_;
  BetaUpdater betaUpdater_;
  Float (*sampler_)(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
  uint32_t stepCount_;
  uint64_t time_;
  uint64_t samplingTime_;
  uint64_t edgesProcessed_;
  SamplerStats stats_;  // sampler-thread time by stage, bytes copied
  uint64_t tKernelsHost_ = 0, tDrain_ = 0;  // main thread: launching / waiting for the GPU (ns)
  std::unique_ptr<Sample> samples_[2];  // two sampler streams, alternating (learner.h:64)
  static const int kInFlight = 3;       // iterations enqueued before the host waits for the oldest
  ammsb_event* iterDone_[kInFlight];
  int phase_;
  Float* betaMirror_;
};

}  // namespace mcmc

#endif  // MCMC_B200_LEARNER_H_
