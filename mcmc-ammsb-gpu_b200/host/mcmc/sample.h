// mcmc/sample.h -- mini-batch strategies (host) and the neighbor sampler (device).
// Reference: mcmc/sample.h:16-125, sample.cc.  The six strategies are host code by
// contract: their output (edge order included) is a function of glibc rand_r and
// std::unordered_set iteration order, which is what "same seed, same mini-batch" means
// for the reference.
#ifndef MCMC_B200_SAMPLE_H_
#define MCMC_B200_SAMPLE_H_

#include <iostream>
#include <memory>
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

// one in-flight mini-batch: host vectors, device copies, its own queue and sampler
struct Sample {
  clcuda::Queue queue;
  std::vector<Edge> edges;
  clcuda::Buffer<Edge> dev_edges;
  std::vector<Vertex> nodes_vec;
  clcuda::Buffer<Vertex> dev_nodes;
  unsigned int seed;
  NeighborSampler neighbor_sampler;

  Sample(const Config& cfg, clcuda::Queue queue);
  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);
};

}  // namespace mcmc

#endif  // MCMC_B200_SAMPLE_H_
