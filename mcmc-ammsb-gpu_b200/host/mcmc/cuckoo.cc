#include "mcmc/cuckoo.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace mcmc {
namespace cuckoo {

const Edge Set::KEY_INVALID = std::numeric_limits<Edge>::max();

namespace {
// hash multipliers / xor masks, four fallback pairs (reference cuckoo.cc:92-96)
const uint64_t kPrimePairs[4][2] = {{15485807ull, 920429591ull},
                                    {379906717ull, 740320571ull},
                                    {256204747ull, 379927517ull},
                                    {13ull, 17ull}};
}  // namespace

// bins per bucket: 1 + ceil(1.15 n / 8)  ->  load factor ~0.87 (reference cuckoo.cc:98-104)
Set::Set(size_t n)
    : inserted_(0),
      bins_(static_cast<size_t>(1 + std::ceil((1.15 * n) / (NUM_BUCKETS * NUM_SLOTS)))),
      mod_bins_(bins_),
      rand_state_(42),
      max_displacements_(n / 2 + 1),
      prime_idx_(0) {}

size_t Set::Bin(Edge k, size_t bucket) const {
  // (P1 * k) % bins (64-bit wrap of the product) and (k ^ P2) % bins, reference cuckoo.cc:199-209
  return bucket == 0 ? mod_bins_.Mod(kPrimePairs[prime_idx_][0] * k) : mod_bins_.Mod(k ^ kPrimePairs[prime_idx_][1]);
}

bool Set::SetContents(std::vector<Edge>::const_iterator start, std::vector<Edge>::const_iterator end) {
  {
    std::lock_guard<std::mutex> lock(index_mu_);
    index_built_ = false;
    index_.reset();
  }
  for (prime_idx_ = 0; prime_idx_ < 4; ++prime_idx_) {
    cells_.assign(Capacity(), KEY_INVALID);
    bool ok = true;
    for (auto it = start; ok && it != end; ++it) ok = Insert(*it);
    if (ok) return true;
    std::cerr << "cuckoo::Set: attempt " << prime_idx_ << " failed" << std::endl;
  }
  return false;
}

// Random-walk insertion.  The order of the rand_r draws (bucket choice, then victim
// slot only when the chosen bin is full) decides the final layout, so it follows the
// reference exactly (cuckoo.cc:140-161,187-197).
bool Set::Insert(Edge k) {
  size_t displaced = 0;
  do {
    for (size_t b = 0; b < NUM_BUCKETS; ++b) {
      Edge* cell = Cell(b, Bin(k, b));
      bool has_room = false, present = false;
      for (size_t s = 0; s < NUM_SLOTS; ++s) {
        if (cell[s] == KEY_INVALID) has_room = true;
        if (cell[s] == k) {
          present = true;
          break;
        }
      }
      if (has_room && !present) {
        for (size_t s = 0; s < NUM_SLOTS; ++s) {
          if (cell[s] == KEY_INVALID) {
            cell[s] = k;
            break;
          }
        }
        ++inserted_;
        return true;
      }
    }
    const size_t b = rand_r(&rand_state_) % NUM_BUCKETS;
    Edge* cell = Cell(b, Bin(k, b));
    size_t free_slot = NUM_SLOTS;
    for (size_t s = 0; s < NUM_SLOTS; ++s) {
      if (cell[s] == KEY_INVALID) {
        free_slot = s;
        break;
      }
    }
    if (free_slot < NUM_SLOTS) {
      cell[free_slot] = k;
      k = KEY_INVALID;
    } else {
      const size_t victim = rand_r(&rand_state_) % NUM_SLOTS;
      std::swap(k, cell[victim]);
    }
  } while (++displaced < max_displacements_);
  return false;
}

bool Set::Has(Edge k) const {
  for (size_t b = 0; b < NUM_BUCKETS; ++b) {
    const Edge* cell = Cell(b, Bin(k, b));
    for (size_t s = 0; s < NUM_SLOTS; ++s)
      if (cell[s] == k) return true;
  }
  return false;
}

const Set::Partners* Set::PartnerIndex() const {
  std::lock_guard<std::mutex> lock(index_mu_);
  if (index_built_) return index_.get();
  index_built_ = true;
  [this] {
    size_t keys = 0;
    Vertex top = 0;
    for (Edge e : cells_) {
      if (e == KEY_INVALID) continue;
      ++keys;
      top = std::max(top, std::max(static_cast<Vertex>(e >> 32), static_cast<Vertex>(e)));
    }
    if (keys == 0 || keys > kMaxIndexedKeys) return;
    std::unique_ptr<Partners> idx(new Partners);
    idx->offsets.assign(static_cast<size_t>(top) + 2, 0);
    // a key stored with its endpoints out of order can never equal a canonical query: skipped
    for (Edge e : cells_) {
      if (e == KEY_INVALID) continue;
      const Vertex a = static_cast<Vertex>(e >> 32), b = static_cast<Vertex>(e);
      if (a > b) continue;
      ++idx->offsets[a + 1];
      if (a != b) ++idx->offsets[b + 1];
    }
    for (size_t i = 1; i < idx->offsets.size(); ++i) idx->offsets[i] += idx->offsets[i - 1];
    idx->partners.resize(idx->offsets.back());
    std::vector<uint64_t> fill(idx->offsets.begin(), idx->offsets.end() - 1);
    for (Edge e : cells_) {
      if (e == KEY_INVALID) continue;
      const Vertex a = static_cast<Vertex>(e >> 32), b = static_cast<Vertex>(e);
      if (a > b) continue;
      idx->partners[fill[a]++] = b;
      if (a != b) idx->partners[fill[b]++] = a;
    }
    index_ = std::move(idx);
  }();
  return index_.get();
}

OpenClSet::OpenClSet(std::shared_ptr<OpenClSetFactory> factory, clcuda::Queue queue, const Set& set)
    : factory_(factory), queue_(queue) {
  std::vector<Edge> image = set.Serialize();
  AmmsbCheck(ammsb_set_create(queue_(), image.data(), set.BinsPerBucket(), set.PrimeIdx(), &handle_));
}

OpenClSet::~OpenClSet() { ammsb_set_destroy(handle_); }

}  // namespace cuckoo
}  // namespace mcmc
