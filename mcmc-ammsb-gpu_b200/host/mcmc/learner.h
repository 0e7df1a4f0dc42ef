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

// One persistent host thread per Sample buffer, with the interface of the std::future the
// reference gets from std::async (learner.cc:216-229): Launch() = std::async, get()/wait()/
// valid() as std::future.  Persistent so that the thread-local sampling arena and the CUDA
// context binding are paid once, not per iteration.
class SamplerThread {
 public:
  SamplerThread();
  ~SamplerThread();
  void Launch(std::function<Float()> task);
  bool valid() const { return valid_; }
  void wait();
  Float get();          // waits, rethrows a failure of the task, invalidates
  void Reset();         // wait and discard

 private:
  void Loop();
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<Float()> task_;
  bool has_task_ = false, done_ = false, stop_ = false, valid_ = false;
  Float result_ = 0;
  std::exception_ptr error_;
  std::thread thread_;
};

class Learner {
 public:
  // cfg is held by reference (as in the reference): it must outlive the Learner
  Learner(const Config& cfg, clcuda::Queue queue);
  ~Learner();

  void Run(uint32_t max_iters, sig_atomic_t* signaled = nullptr);
  Float HeldoutPerplexity();
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
  uint64_t BytesH2D() const { return h2dBytes_; }  // mini-batch edges + nodes copied to the device so far
  // when set (pinned host memory, 2K floats), every iteration ends with a device->host copy
  // of beta into it: the per-iteration result a monitoring caller reads
  void MirrorBetaTo(Float* pinned_host) { betaMirror_ = pinned_host; }
  // the mini-batch the next iteration will consume (joins the sampler thread)
  const Sample& PeekNextSample();
  Float PeekNextWeight() { PeekNextSample(); return pendingWeight_[phase_]; }

 private:
  Float SampleMiniBatch(std::vector<Edge>* edges, unsigned int* seed);
  void LaunchSampler(int buffer);
  Float DoSample(Sample* sample);

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
  PerplexityCalculator heldoutPerplexity_;
  PhiUpdater phiUpdater_;
  BetaUpdater betaUpdater_;
  Float (*sampler_)(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
  uint32_t stepCount_;
  uint64_t time_;
  uint64_t samplingTime_;
  uint64_t edgesProcessed_;
  std::atomic<uint64_t> h2dBytes_;
  // sampler-thread time by stage, ns (host strategy, node extraction, H2D copies, neighbor kernel)
  std::atomic<uint64_t> tStrategy_{0}, tExtract_{0}, tCopy_{0}, tNeighbor_{0}, tKernelsHost_{0}, tDrain_{0};
  Sample samples_[2];
  SamplerThread futures_[2];
  Float pendingWeight_[2];
  bool pendingValid_[2];
  int phase_;
  Float* betaMirror_;
};

}  // namespace mcmc

#endif  // MCMC_B200_LEARNER_H_
