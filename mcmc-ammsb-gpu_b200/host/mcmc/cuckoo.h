// mcmc/cuckoo.h -- static cuckoo hash set of 64-bit edge keys.
// Host build + membership with the reference's exact placement policy
// (mcmc/cuckoo.h:16-67, cuckoo.cc:98-220) so that tables, prime choice and therefore
// membership answers are identical; the device copy is an ammsb_set.
#ifndef MCMC_B200_CUCKOO_H_
#define MCMC_B200_CUCKOO_H_

#include <array>
#include <limits>
#include <memory>
#include <mutex>
#include <vector>

#include "mcmc/fastmod.h"
#include "mcmc/types.h"

namespace mcmc {
namespace cuckoo {

class Set {
 public:
  static const size_t NUM_BUCKETS = 2;
  static const size_t NUM_SLOTS = 4;
  static const Edge KEY_INVALID;

  explicit Set(size_t n);

  bool SetContents(std::vector<Edge>::const_iterator start, std::vector<Edge>::const_iterator end);
  bool Has(Edge k) const;
  // Has() split in two so that a caller can hash a batch of keys, prefetch their bins and
  // only then compare (the host mini-batch sampler does; the answers are those of Has)
  void Locate(Edge k, size_t bins[NUM_BUCKETS]) const {
    bins[0] = Bin(k, 0);
    bins[1] = Bin(k, 1);
    __builtin_prefetch(Cell(0, bins[0]));
    __builtin_prefetch(Cell(1, bins[1]));
  }
  bool HasAt(Edge k, const size_t bins[NUM_BUCKETS]) const {
    for (size_t b = 0; b < NUM_BUCKETS; ++b) {
      const Edge* cell = Cell(b, bins[b]);
      for (size_t s = 0; s < NUM_SLOTS; ++s)
        if (cell[s] == k) return true;
    }
    return false;
  }

  size_t BinsPerBucket() const { return bins_; }
  size_t Capacity() const { return bins_ * NUM_SLOTS * NUM_BUCKETS; }
  // number of successful Insert() calls over all attempts (the reference never resets it)
  size_t Size() const { return inserted_; }
  uint32_t PrimeIdx() const { return prime_idx_; }

  // flat [bucket][bin][slot] image, the layout the device lookup indexes
  std::vector<Edge> Serialize() const { return cells_; }

  // The stored keys seen from one endpoint: Partners(u) lists every v for which the canonical
  // key (min(u,v) << 32 | max(u,v)) is stored, so Has(canonical(u, v)) == "v is in Partners(u)".
  // Built from the cells on first use (hence consistent with Has() by construction) and only
  // for sets of at most kMaxIndexedKeys keys; PartnerIndex() returns nullptr otherwise.  The
  // non-link mini-batch strategy tests thousands of pairs that share one endpoint: with the
  // partner list of that endpoint it needs no table lookups (each of which is two cache misses).
  struct Partners {
    std::vector<uint64_t> offsets;  // [max vertex + 2]
    std::vector<Vertex> partners;
    const Vertex* begin(Vertex u) const { return u + 1 < offsets.size() ? partners.data() + offsets[u] : nullptr; }
    const Vertex* end(Vertex u) const { return u + 1 < offsets.size() ? partners.data() + offsets[u + 1] : nullptr; }
  };
  static const size_t kMaxIndexedKeys = size_t(1) << 27;
  const Partners* PartnerIndex() const;

 private:
  size_t Bin(Edge k, size_t bucket) const;
  Edge* Cell(size_t bucket, size_t bin) { return &cells_[(bucket * bins_ + bin) * NUM_SLOTS]; }
  const Edge* Cell(size_t bucket, size_t bin) const { return &cells_[(bucket * bins_ + bin) * NUM_SLOTS]; }
  bool Insert(Edge k);

  size_t inserted_;
  const size_t bins_;
  const FastMod64 mod_bins_;  // k % bins_ without a divide (same value)
  std::vector<Edge> cells_;
  unsigned int rand_state_;
  const size_t max_displacements_;
  uint32_t prime_idx_;
  mutable std::mutex index_mu_;
  mutable bool index_built_ = false;
  mutable std::unique_ptr<Partners> index_;
};

class OpenClSetFactory;

// device-resident copy of a Set (reference: OpenClSet, cuckoo.h:71-86)
class OpenClSet {
 public:
  ~OpenClSet();
  ammsb_set* Get() const { return handle_; }

 private:
  OpenClSet(std::shared_ptr<OpenClSetFactory> factory, clcuda::Queue queue, const Set& set);
  std::shared_ptr<OpenClSetFactory> factory_;
  clcuda::Queue queue_;
  ammsb_set* handle_ = nullptr;
  friend class OpenClSetFactory;
};

class OpenClSetFactory : public std::enable_shared_from_this<OpenClSetFactory> {
 public:
  static std::shared_ptr<OpenClSetFactory> New(clcuda::Queue queue) {
    return std::shared_ptr<OpenClSetFactory>(new OpenClSetFactory(queue));
  }
  OpenClSet* CreateSet(const Set& set) { return new OpenClSet(shared_from_this(), queue_, set); }

 private:
  explicit OpenClSetFactory(clcuda::Queue queue) : queue_(queue) {}
  clcuda::Queue queue_;
};

}  // namespace cuckoo
}  // namespace mcmc

#endif  // MCMC_B200_CUCKOO_H_
