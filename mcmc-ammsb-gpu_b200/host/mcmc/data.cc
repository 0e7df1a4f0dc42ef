#include "mcmc/data.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <unordered_map>
#include <unordered_set>

namespace mcmc {

Graph::Graph(uint64_t num_nodes, const std::vector<Edge>& unique_edges)
    : num_nodes_(num_nodes), unique_edges_(unique_edges), adjacency_(num_nodes), max_fan_out_(0) {
  for (Edge e : unique_edges_) {
    const Vertex u = static_cast<Vertex>(e >> 32), v = static_cast<Vertex>(e);
    adjacency_[u].push_back(v);
    adjacency_[v].push_back(u);
    max_fan_out_ = std::max<uint64_t>(max_fan_out_, std::max(adjacency_[u].size(), adjacency_[v].size()));
  }
}

Edge Graph::GetRandomEdge() const {
  Vertex u;
  do {
    u = rand() % num_nodes_;
  } while (adjacency_[u].empty());
  return MakeEdge(u, adjacency_[u][rand() % adjacency_[u].size()]);
}

// SNAP text format: four header lines, then "a b" pairs.  Vertices are renumbered in
// std::unordered_set iteration order, edges sorted, de-duplicated and shuffled with
// std::random_shuffle (libc rand()) -- reference data.cc:36-78.
bool GetUniqueEdgesFromFile(const std::string& filename, uint64_t* count_vertices, std::vector<Edge>* vals) {
  std::ifstream in(filename);
  std::string header;
  for (int i = 0; i < 4; ++i) std::getline(in, header);
  std::unordered_set<Vertex> seen;
  std::vector<Edge> raw;
  do {
    uint64_t a, b;
    in >> a >> b;
    if (!in.eof()) {
      const uint64_t lo = std::min(a, b), hi = std::max(a, b);
      raw.push_back(MakeEdge(lo, hi));
      seen.insert(lo);
      seen.insert(hi);
    }
  } while (in.good());
  if (in.bad()) {
    std::cerr << "Error reading file " << filename << std::endl;
    return false;
  }
  std::unordered_map<Vertex, Vertex> renumber;
  Vertex next = 0;
  for (Vertex v : seen) renumber[v] = next++;
  *count_vertices = renumber.size();
  for (Edge e : raw) vals->push_back(MakeEdge(renumber[static_cast<Vertex>(e >> 32)], renumber[static_cast<Vertex>(e)]));
  std::sort(vals->begin(), vals->end());
  vals->erase(std::unique(vals->begin(), vals->end()), vals->end());
  std::random_shuffle(vals->begin(), vals->end());
  return true;
}

// Held-out links = the first E - ceil((1 - r/2) E) entries of the shuffled list, training =
// the rest; as many fake non-links (libc rand(), u != v, in neither set, unique) are appended
// to the held-out list -- reference data.cc:80-128.
bool GenerateSetsFromEdges(uint64_t N, const std::vector<Edge>& vals, double heldout_ratio,
                           std::vector<Edge>* training_edges, std::vector<Edge>* heldout_edges,
                           std::unique_ptr<Set>* training, std::unique_ptr<Set>* heldout) {
  const size_t training_len = static_cast<size_t>(std::ceil((1 - heldout_ratio / 2) * vals.size()));
  const size_t heldout_len = vals.size() - training_len;
  const auto split = vals.begin() + heldout_len;
  if (heldout_len > 0) {
    heldout->reset(new Set(heldout_len));
    if (!(*heldout)->SetContents(vals.begin(), split)) {
      std::cerr << "Failed to insert into heldout set" << std::endl;
      heldout->reset();
      return false;
    }
    heldout_edges->insert(heldout_edges->end(), vals.begin(), split);
  }
  training->reset(new Set(training_len));
  if (!(*training)->SetContents(split, vals.end())) {
    std::cerr << "Failed to insert into training set" << std::endl;
    training->reset();
    if (heldout_len > 0) heldout->reset();
    return false;
  }
  training_edges->insert(training_edges->end(), split, vals.end());
  std::unordered_set<Edge> fakes;
  for (size_t i = 0; i < heldout_len; ++i) {
    Edge e;
    do {
      const Vertex u = rand() % N;
      Vertex v;
      do {
        v = rand() % N;
      } while (u == v);
      e = MakeEdge(std::min(u, v), std::max(u, v));
    } while (fakes.count(e) || (*heldout)->Has(e) || (*training)->Has(e));
    fakes.insert(e);
    heldout_edges->push_back(e);
  }
  return true;
}

bool GenerateSetsFromFile(const std::string& filename, double heldout_ratio, uint64_t* count_vertices,
                          std::vector<Edge>* training_edges, std::vector<Edge>* heldout_edges,
                          std::unique_ptr<Set>* training, std::unique_ptr<Set>* heldout) {
  std::vector<Edge> vals;
  return GetUniqueEdgesFromFile(filename, count_vertices, &vals) &&
         GenerateSetsFromEdges(*count_vertices, vals, heldout_ratio, training_edges, heldout_edges, training,
                               heldout);
}

bool DumpDataset(const std::string& path, uint64_t N, Float heldout_ratio, const std::vector<Edge>& edges) {
  gzFile f = gzopen(path.c_str(), "wb");
  if (!f) return false;
  const uint64_t count = edges.size();
  bool ok = gzwrite(f, &N, sizeof N) == (int)sizeof N &&
            gzwrite(f, &heldout_ratio, sizeof heldout_ratio) == (int)sizeof heldout_ratio &&
            gzwrite(f, &count, sizeof count) == (int)sizeof count;
  const char* p = reinterpret_cast<const char*>(edges.data());
  size_t left = count * sizeof(Edge);
  while (ok && left) {
    const unsigned chunk = static_cast<unsigned>(std::min<size_t>(left, 1u << 30));
    ok = gzwrite(f, p, chunk) == (int)chunk;
    p += chunk;
    left -= chunk;
  }
  return gzclose(f) == Z_OK && ok;
}

bool LoadDataset(const std::string& path, uint64_t* N, Float* heldout_ratio, std::vector<Edge>* edges) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) return false;
  uint64_t count = 0;
  bool ok = gzread(f, N, sizeof *N) == (int)sizeof *N &&
            gzread(f, heldout_ratio, sizeof *heldout_ratio) == (int)sizeof *heldout_ratio &&
            gzread(f, &count, sizeof count) == (int)sizeof count;
  if (ok) {
    edges->resize(count);
    char* p = reinterpret_cast<char*>(edges->data());
    size_t left = count * sizeof(Edge);
    while (ok && left) {
      const unsigned chunk = static_cast<unsigned>(std::min<size_t>(left, 1u << 30));
      ok = gzread(f, p, chunk) == (int)chunk;
      p += chunk;
      left -= chunk;
    }
  }
  gzclose(f);
  return ok;
}

}  // namespace mcmc
