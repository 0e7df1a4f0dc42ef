// mcmc/std_order_set.h -- an insert-only hash set whose ITERATION ORDER is that of libstdc++'s
// std::unordered_set<Key> (std::hash = identity for integers) fed the same insert sequence.
//
// Why: the reference emits every mini-batch in std::unordered_set iteration order
// (sample.cc:267,290; learner.cc:164-172), so "same seed -> same mini-batch" includes that
// order.  std::unordered_set pays a heap node per element, pointer chasing and an integer
// division per chain step.  This container does not keep libstdc++'s node list at all; it
// computes the order the list would have, from two facts about libstdc++'s hashtable:
//
//  (1) A new node whose bucket is empty goes to the very front of the list; a new node whose
//      bucket is not empty goes to the front of its bucket's run.  A rehash visits the nodes in
//      list order and places each one by the same rule into the new bucket array.  So if a
//      table (empty, B buckets) is fed the sequence S, the list ends up as
//          R(S, B) = reverse( S stably grouped by bucket, groups in order of first appearance )
//      and a rehash of list L to B' buckets followed by the inserts T gives R(L ++ T, B').
//  (2) When a rehash happens, and to how many buckets, depends only on the element count: it is
//      decided by std::__detail::_Prime_rehash_policy, the very object used here.
//
// Insert() therefore only de-duplicates (open addressing, no order) and appends the key to the
// insert sequence; EmitTo() replays the ~log2(n) growth phases as counting passes over flat
// arrays (every pass: a remainder without a divide, one counter table, sequential reads) --
// no dependent loads, about a fifth of the time of the linked structure.  tests/test_host.py
// compares it with std::unordered_set on random insert sequences.
//
// The same two facts are what a data-parallel (GPU) version would use: each phase is a stable
// sort by (first appearance of the bucket, position).
#ifndef MCMC_B200_STD_ORDER_SET_H_
#define MCMC_B200_STD_ORDER_SET_H_

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <unordered_set>
#include <utility>
#include <vector>

#include "mcmc/fastmod.h"

namespace mcmc {

template <class Key>
class StdOrderSet {
 public:
  StdOrderSet() { Clear(); }

  void Clear() {
    keys_.clear();
    blocked_.clear();
    has_empty_key_ = blocked_empty_key_ = false;
    // the table restarts small (a link mini-batch has a handful of keys) but keeps its memory
    if (table_.size() < kMinTable) table_.resize(kMinTable);
    mask_ = kMinTable - 1;
    shift_ = 64 - kMinTableLog2;
    std::fill(table_.begin(), table_.begin() + kMinTable, kEmptyKey);
  }

  // Call right after Clear() when the number of keys to come is known (a non-link mini-batch: m
  // edges; its endpoints: m + 1 vertices): the de-duplication table takes its final size at once
  // instead of growing x4 from 64 entries with a refill of everything inserted so far each time
  // (at m = 131072 that is 1.3 extra cache-missing inserts per key).  Order and results do not
  // depend on the table size.
  void Reserve(size_t n) {
    unsigned log2 = kMinTableLog2;
    while ((size_t(1) << log2) < 2 * (n + 1)) ++log2;
    const size_t size = size_t(1) << log2;
    if (size <= mask_ + 1) return;
    if (table_.size() < size) table_.resize(size);
    mask_ = size - 1;
    shift_ = 64 - log2;
    std::fill(table_.begin(), table_.begin() + size, kEmptyKey);
    keys_.reserve(n);
  }

  size_t size() const { return keys_.size(); }

  // true if k was not present
  bool Insert(Key k) {
    if (k == kEmptyKey) {  // the one key the open-addressing table cannot hold
      if (has_empty_key_ || blocked_empty_key_) return false;
      has_empty_key_ = true;
      keys_.push_back(k);
      return true;
    }
    if ((keys_.size() + blocked_.size() + 1) * 2 > mask_ + 1) Grow();
    const size_t mask = mask_;
    size_t i = Slot(k);
    while (table_[i] != kEmptyKey) {
      if (table_[i] == k) return false;
      i = (i + 1) & mask;
    }
    table_[i] = k;
    keys_.push_back(k);
    return true;
  }

  // Append a key the CALLER knows to be new (it de-duplicates by other means): the key joins the
  // insert sequence without touching the table.  Not to be mixed with Insert() of keys that may
  // repeat an appended one.
  void AppendUnique(Key k) { keys_.push_back(k); }

  // From now on Insert(k) is refused (returns false) as if k were present, but k is not an
  // element: it is neither counted by size() nor emitted.  Lets a caller fold a small list of
  // forbidden keys into the de-duplication probe it pays for anyway.
  void Block(Key k) {
    if (k == kEmptyKey) {
      if (!has_empty_key_) blocked_empty_key_ = true;
      return;
    }
    if ((keys_.size() + blocked_.size() + 1) * 2 > mask_ + 1) Grow();
    size_t i = Slot(k);
    while (table_[i] != kEmptyKey) {
      if (table_[i] == k) return;
      i = (i + 1) & mask_;
    }
    table_[i] = k;
    blocked_.push_back(k);
  }

  // keys in insertion order
  const std::vector<Key>& InsertionOrder() const { return keys_; }

  // elements in iteration order, inserted at the front of *out (as vector::insert(begin(), ...))
  template <class T>
  void EmitTo(std::vector<T>* out) {
    const size_t old = out->size(), n = keys_.size();
    out->resize(old + n);
    if (old) std::move_backward(out->begin(), out->begin() + old, out->end());
    const Key* order = ComputeOrder();
    for (size_t i = 0; i < n; ++i) (*out)[i] = static_cast<T>(order[i]);
  }

 private:
  static constexpr Key kEmptyKey = static_cast<Key>(~static_cast<Key>(0));
  static constexpr unsigned kMinTableLog2 = 6;
  static constexpr size_t kMinTable = size_t(1) << kMinTableLog2;

  size_t Slot(Key k) const { return static_cast<size_t>((static_cast<uint64_t>(k) * 0x9E3779B97F4A7C15ull) >> shift_); }

  void Grow() {  // x4: the table is refilled from the insertion sequence
    const size_t size = (mask_ + 1) * 4;
    if (table_.size() < size) table_.resize(size);
    mask_ = size - 1;
    shift_ -= 2;
    std::fill(table_.begin(), table_.begin() + size, kEmptyKey);
    for (const std::vector<Key>* src : {&keys_, &blocked_}) {
      for (Key k : *src) {
        if (k == kEmptyKey) continue;
        size_t i = Slot(k);
        while (table_[i] != kEmptyKey) i = (i + 1) & mask_;
        table_[i] = k;
      }
    }
  }

  // One growth phase: `seq` (n keys) fed to an empty table of `buckets` buckets; writes the
  // resulting list order R(seq, buckets) to `dst`.
  void Phase(const Key* seq, size_t n, size_t buckets, Key* dst) {
    first_.assign(buckets, kNoGroup);
    if (gid_.size() < n) gid_.resize(n);
    gsize_.assign(std::min(n, buckets) + 1, 0);
    uint32_t groups = 0;
    const FastMod64 mod(buckets);
    for (size_t i = 0; i < n; ++i) {
      const size_t b = static_cast<size_t>(mod.Mod(static_cast<uint64_t>(seq[i])));  // std::hash = identity
      // branch-free: whether a key opens a new group is a coin flip the predictor loses
      uint32_t g = first_[b];
      const uint32_t opens = g == kNoGroup;
      g = opens ? groups : g;
      first_[b] = g;
      groups += opens;
      gid_[i] = g;
      ++gsize_[g];
    }
    // groups in order of first appearance, the whole thing reversed: group g ends where the
    // groups after it begin, and inside a group later keys come first
    size_t end = n;
    for (uint32_t g = 0; g < groups; ++g) {
      const size_t len = gsize_[g];
      gsize_[g] = static_cast<uint32_t>(end);  // one past the slot of the group's first key
      end -= len;
    }
    for (size_t i = 0; i < n; ++i) dst[--gsize_[gid_[i]]] = seq[i];
  }

  const Key* ComputeOrder() {
    const size_t n = keys_.size();
    if (a_.size() < n) {
      a_.resize(n);
      b_.resize(n);
    }
    Key* cur = a_.data();  // the list after the phases so far (its first `have` entries)
    Key* nxt = b_.data();
    size_t have = 0;       // == number of keys already placed
    // replay libstdc++'s growth decisions: before the insert of element i (0-based) the table
    // asks the policy whether i + 1 elements still fit (unordered_set::insert ->
    // _M_insert_unique_node -> _M_need_rehash(bucket_count, element_count, 1))
    std::__detail::_Prime_rehash_policy policy;
    size_t buckets = 1, phase_buckets = 0;
    size_t i = 0;
    auto run_phase = [&](size_t upto) {  // list = R(list ++ keys_[have, upto), phase_buckets)
      if (phase_buckets == 0 || upto == have) return;
      std::memcpy(cur + have, keys_.data() + have, (upto - have) * sizeof(Key));
      Phase(cur, upto, phase_buckets, nxt);
      std::swap(cur, nxt);
      have = upto;
    };
    while (i < n) {
      if (i + 1 > policy._M_next_resize) {
        const std::pair<bool, size_t> grow = policy._M_need_rehash(buckets, i, 1);
        if (grow.first) {
          run_phase(i);  // the inserts made with the old bucket count
          buckets = phase_buckets = grow.second;
        }
      }
      // nothing can change before element number _M_next_resize
      i = std::max<size_t>(i + 1, policy._M_next_resize);
    }
    run_phase(n);
    return cur;
  }

  static constexpr uint32_t kNoGroup = 0xffffffffu;

  std::vector<Key> keys_;   // insertion order
  std::vector<Key> table_;  // open addressing, de-duplication only; logical size mask_ + 1
  size_t mask_ = kMinTable - 1;
  unsigned shift_ = 64 - kMinTableLog2;
  std::vector<Key> blocked_;  // in the table, not elements
  bool has_empty_key_ = false, blocked_empty_key_ = false;
  std::vector<Key> a_, b_;  // ping-pong list buffers
  std::vector<uint32_t> first_, gid_, gsize_;
};

}  // namespace mcmc

#endif  // MCMC_B200_STD_ORDER_SET_H_
