// mcmc/sharded_learner.h -- mcmc::Learner's iteration (learner.cc:214-250) over SEVERAL GPUs of one
// box, behind the same Config: the multi-GPU successor of the reference's only scale mechanism,
// RowPartitionedMatrix (partitioned-alloc.h:14-141).
//
// pi is COLUMN-SHARDED (csrc/cols.cu): rank g of G (2, 4 or 8) holds the columns of the reference
// work-items l = g (mod G) of every row; every row read is local HBM and only 4-byte partial sums
// cross NVLink, inside the kernels, completed in the reference's WG_SUM tree order -- update_phi /
// update_pi results are those of the one-GPU kernels bit for bit, for every G.
//
// One host thread drives all ranks: the kernels of a stage are launched rank after rank
// (asynchronously, each on its device's stream) and meet through their mailboxes.  `devices` names
// the device of every rank; ranks that share a device are computed by ONE cooperative launch (the
// emulation the one-GPU tests use), so `{0, 0}` runs the two-rank protocol on a single GPU and
// `{0, 1, ..., 7}` is the production layout.
#ifndef MCMC_B200_SHARDED_LEARNER_H_
#define MCMC_B200_SHARDED_LEARNER_H_

#include <condition_variable>
#include <csignal>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "mcmc/config.h"
#include "mcmc/sample.h"
#include "mcmc/types.h"

namespace mcmc {

class ShardedLearner {
 public:
  ShardedLearner(const Config& cfg, const std::vector<int>& devices);
  ~ShardedLearner();
  ShardedLearner(const ShardedLearner&) = delete;
  ShardedLearner& operator=(const ShardedLearner&) = delete;

  void Run(uint32_t max_iters, sig_atomic_t* signaled = nullptr);
  Float HeldoutPerplexity();
  void PrintStats();

  uint32_t World() const { return static_cast<uint32_t>(ranks_.size()); }
  uint64_t EdgesProcessed() const { return edgesProcessed_; }
  // state in the reference's layout, assembled from the ranks' columns
  void ReadPi(uint64_t row0, uint64_t nrows, Float* rows);        // [nrows][K]
  void ReadPhi(uint64_t row0, uint64_t nrows, Float* sums);       // [nrows]
  void ReadTheta(uint32_t rank, Float* theta, Float* beta);       // [2K] each, as rank `rank` holds them

 private:
  struct MiniBatch {
    std::vector<Edge> edges;
    std::vector<Vertex> nodes;
    Float weight = 0;
  };
  struct Group;  // the ranks of one device
  MiniBatch Draw(int stream);
  void Producer(int stream);
  MiniBatch Next(int stream);
  void Upload(const MiniBatch& mb, int slot);

  const Config& cfg_;
  ammsb_params params_;
  ammsb_phi_opts opts_;
  std::vector<ammsb_cols*> ranks_;             // by rank
  std::vector<std::unique_ptr<Group>> groups_;  // by device
  Float (*sampler_)(const Config&, std::vector<Edge>*, unsigned int*);
  // the two sampler streams (the reference's Samples: own seed each, learner.cc:216-232), each drawn
  // by its own thread a few mini-batches ahead; a stream's mini-batches are drawn strictly in order
  static const size_t kAhead = 3;
  unsigned int seeds_[2];
  std::thread producers_[2];
  std::deque<MiniBatch> ready_[2];
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false;
  std::exception_ptr error_;
  uint32_t stepCount_ = 0, ppxCalls_ = 0;
  int phase_ = 0;
  uint64_t edgesProcessed_ = 0;
  uint64_t time_ = 0, samplingTime_ = 0;
};

}  // namespace mcmc

#endif  // MCMC_B200_SHARDED_LEARNER_H_
