// mcmc/sample.h -- mini-batch strategies (host) and the neighbor sampler (device).
// Reference: mcmc/sample.h:16-125, sample.cc.  The six strategies are host code by
// contract: their output (edge order included) is a function of glibc rand_r and
// std::unordered_set iteration order, which is what "same seed, same mini-batch" means
// for the reference.
#ifndef MCMC_B200_SAMPLE_H_
#define MCMC_B200_SAMPLE_H_

#include <atomic>
#include <condition_variable>
#include <exception>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "mcmc/random.h"
#include "mcmc/types.h"

namespace mcmc {

struct Config;

enum SampleStrategy { Node, NodeLink, NodeNonLink, BFLink, BFNonLink, BF };
std::string to_string(const SampleStrategy& s);
std::istream& operator>>(std::istream& in, SampleStrategy& strategy);

// each returns the mini-batch scale ("weight") and fills `edges`
Float sampleBreadthFirstLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
Float sampleBreadthFirstNonLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
Float sampleBreadthFirst(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
Float sampleNodeLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
Float sampleNodeNonLink(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
Float sampleNode(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);

// endpoints of a mini-batch in std::unordered_set<Vertex> order (Learner::ExtractNodesFromMiniBatch)
void ExtractNodesFromMiniBatch(const std::vector<Edge>& edges, std::vector<Vertex>* nodes);

// capacities the reference allocates with (sample.cc:86-97,129-131)
uint64_t MaxMiniBatchNodes(const Config& cfg);
uint64_t MaxMiniBatchEdges(const Config& cfg);

class NeighborSampler {
 public:
  NeighborSampler(const Config& cfg, clcuda::Queue queue);

  // draws cfg.num_node_sample distinct neighbors != node for the first num_samples nodes
  void operator()(uint32_t num_samples, clcuda::Buffer<Vertex>* nodes);
  // same, and a copy of the result into `out` (a ring slot); GetData() keeps the latest result
  void operator()(uint32_t num_samples, clcuda::Buffer<Vertex>* nodes, clcuda::Buffer<Vertex>* out);

  clcuda::Buffer<Vertex>& GetHash() { return hash_; }
  clcuda::Buffer<Vertex>& GetData() { return data_; }
  uint32_t HashCapacityPerSample() { return capacity_; }
  uint32_t DataSizePerSample();
  random::OpenClRandom* GetRandom() { return rand_.get(); }
  // when set, the open-addressing tables are also written to GetHash() (test hook)
  void ExportHash(bool on) { export_hash_ = on; }

  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);

 private:
  const Config& cfg_;
  uint32_t capacity_;
  uint32_t local_;
  clcuda::Queue queue_;
  clcuda::Buffer<Vertex> hash_;
  clcuda::Buffer<Vertex> data_;
  std::shared_ptr<random::OpenClRandomFactory> randFactory_;
  std::unique_ptr<random::OpenClRandom> rand_;
  bool export_hash_ = false;
};

// host time spent by the sampler threads, by stage (ns), and bytes copied to the device
struct SamplerStats {
  std::atomic<uint64_t> strategy{0}, extract{0}, copy{0}, neighbors{0}, h2d_bytes{0};
};

// one drawn mini-batch: host vectors and their device copies
struct SampleSlot {
  std::vector<Edge> edges;
  std::vector<Vertex> nodes_vec;
  clcuda::Buffer<Edge> dev_edges;
  clcuda::Buffer<Vertex> dev_nodes;
  clcuda::Buffer<Vertex> neighbors;  // [nodes, num_node_sample]
  Float weight = 0;
  SampleSlot(const Config& cfg, const clcuda::Context& ctx);
};

// The graph side of the device mini-batch strategies (csrc/graph.cu + csrc/orderset.cu): the
// training adjacency in mcmc::Graph's order (data.cc:12-25) and the vertex degrees, uploaded once
// and shared by the sampler streams; the cuckoo sets are the Learner's device sets.
struct DeviceStrategyData {
  DeviceStrategyData(const Config& cfg, clcuda::Queue queue, ammsb_set* training, ammsb_set* heldout);
  clcuda::Buffer<uint64_t> offsets;  // [N + 1]
  clcuda::Buffer<Vertex> adjacency;  // [2 |training|]
  std::vector<uint32_t> degree;
  ammsb_set* training;
  ammsb_set* heldout;
};

// One sampler stream (the reference's struct Sample, sample.h:51-92: its own seed, queue and
// NeighborSampler with its own RNG pool).  The Learner alternates between two of them.
//
// The reference draws mini-batch t+1 on one std::async thread while t is processed
// (learner.cc:216-232).  Here each stream is a three-stage pipeline over a small ring of slots,
//   stage A (thread): host strategy with the stream's seed -> slot.edges, weight
//   stage B (thread): node extraction                      -> slot.nodes_vec
//   stage C (thread): H2D copies, neighbor sampling        -> device buffers
// so several mini-batches of a stream are in flight.  (Stage C mostly waits: the neighbor
// kernel shares the GPU with an update_phi launch that holds every SM.)  Nothing in either stage reads the model,
// and each stage handles the stream's mini-batches strictly in order (seed and RNG pool advance
// exactly as in the reference), so the mini-batches are the same; only how far ahead they are
// drawn differs.  `Allow()` bounds that: a Run(n) call lets the streams draw only the n
// mini-batches it consumes plus the one the reference leaves in flight, so the state seen by
// Serialize() on return is the reference's.
struct Sample {
  typedef Float (*Strategy)(const Config& cfg, std::vector<Edge>* edges, unsigned int* seed);
  static const uint64_t kRing = 6;

  clcuda::Queue queue;
  std::vector<std::unique_ptr<SampleSlot>> ring;
  unsigned int seed;
  NeighborSampler neighbor_sampler;

  Sample(const Config& cfg, clcuda::Queue queue);
  ~Sample();
  void Start(Strategy strategy, SamplerStats* stats);
  // Config::device_sampler: stage A draws the mini-batch ON THE DEVICE -- the host draws the coin
  // and the vertex with rand_r exactly as sampleNode / sampleNodeLink / sampleNodeNonLink do
  // (sample.cc:253-302), the device examines the candidate stream, puts edges and nodes in the
  // reference's std::unordered_set order and leaves the seed where the host strategy leaves it:
  // the mini-batch is the host strategy's, element for element.  Stages B and the copies of stage
  // C fall away; the host vectors are read back (Serialize() stores them).
  void StartOnDevice(SampleStrategy which, std::shared_ptr<DeviceStrategyData> data, SamplerStats* stats);
  // let the stream draw until `more` mini-batches beyond the consumed ones exist
  void Allow(uint64_t more);
  // the oldest drawn-but-unconsumed mini-batch, fully on the device (blocks; rethrows a failure);
  // skip = mini-batches handed out earlier and not yet Release()d (still read by the GPU)
  SampleSlot& WaitReady(uint64_t skip = 0);
  void Release();  // the mini-batch returned by WaitReady() has been consumed
  // stop drawing and wait for the stages to drain; returns the number of pending mini-batches
  uint64_t Quiesce();
  SampleSlot& Latest();  // most recently drawn mini-batch (the reference's Sample contents)
  bool Serialize(std::ostream* out);
  // pending: the parsed mini-batch is the next one to consume (else it was already consumed)
  bool Parse(std::istream* in, bool pending);

 private:
  void StageA();
  void StageB();
  void StageC();
  Float DrawOnDevice(SampleSlot* slot);
  const Config& cfg_;
  std::shared_ptr<DeviceStrategyData> dev_;
  SampleStrategy dev_strategy_ = Node;
  ammsb_sampler* dev_sampler_ = nullptr;
  ammsb_orderset* dev_order_ = nullptr;
  Strategy strategy_ = nullptr;
  SamplerStats* stats_ = nullptr;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t drawn_ = 0, extracted_ = 0, ready_ = 0, consumed_ = 0, allowed_ = 0;  // mini-batches of this stream
  bool stop_ = false, a_busy_ = false;
  std::exception_ptr error_;
  std::thread a_, b_, c_;
};

}  // namespace mcmc

#endif  // MCMC_B200_SAMPLE_H_
