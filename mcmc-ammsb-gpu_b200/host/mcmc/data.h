// mcmc/data.h -- graph container, SNAP edge-list loader and the training / held-out
// split.  Same entry points and semantics as the reference's mcmc/data.h:16-52
// (data.cc:12-146): results depend on libc rand()/std::unordered_set order and are
// reproduced by using the same library calls in the same order.
#ifndef MCMC_B200_DATA_H_
#define MCMC_B200_DATA_H_

#include <memory>
#include <string>
#include <vector>

#include "mcmc/cuckoo.h"

namespace mcmc {

using namespace mcmc::cuckoo;  // mcmc::Set is the cuckoo set (reference data.h:13-14)

class Graph {
 public:
  Graph(uint64_t num_nodes, const std::vector<Edge>& unique_edges);

  Edge GetRandomEdge() const;
  const std::vector<Vertex>& NeighborsOf(Vertex u) const { return adjacency_[u]; }
  const std::vector<Edge>& UniqueEdges() const { return unique_edges_; }
  uint64_t MaxFanOut() const { return max_fan_out_; }

 private:
  uint64_t num_nodes_;
  std::vector<Edge> unique_edges_;
  std::vector<std::vector<Vertex>> adjacency_;
  uint64_t max_fan_out_;
};

bool GetUniqueEdgesFromFile(const std::string& filename, uint64_t* count_vertices, std::vector<Edge>* vals);

bool GenerateSetsFromEdges(uint64_t N, const std::vector<Edge>& vals, double heldout_ratio,
                           std::vector<Edge>* training_edges, std::vector<Edge>* heldout_edges,
                           std::unique_ptr<Set>* training, std::unique_ptr<Set>* heldout);

bool GenerateSetsFromFile(const std::string& filename, double heldout_ratio, uint64_t* count_vertices,
                          std::vector<Edge>* training_edges, std::vector<Edge>* heldout_edges,
                          std::unique_ptr<Set>* training, std::unique_ptr<Set>* heldout);

// gzip dataset dump of main.cc:109-143: u64 N, f32 heldout_ratio, u64 num_edges, u64 edges[]
bool DumpDataset(const std::string& path, uint64_t N, Float heldout_ratio, const std::vector<Edge>& edges);
bool LoadDataset(const std::string& path, uint64_t* N, Float* heldout_ratio, std::vector<Edge>* edges);

}  // namespace mcmc

#endif  // MCMC_B200_DATA_H_
