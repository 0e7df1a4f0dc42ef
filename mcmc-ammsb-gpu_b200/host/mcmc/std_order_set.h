// mcmc/std_order_set.h -- an insert-only hash set whose ITERATION ORDER is that of libstdc++'s
// std::unordered_set<Key> (std::hash = identity for integers) fed the same insert sequence.
//
// Why: the reference emits every mini-batch in std::unordered_set iteration order
// (sample.cc:267,290; learner.cc:164-172), so "same seed -> same mini-batch" includes that
// order.  std::unordered_set pays a heap node per element, pointer chasing, and one integer
// division per chain step (integer hashes are not cached in the node).  This container keeps
// the same singly-linked node list and "bucket -> node before its first node" table, but in
// flat arrays (32-bit links, the bucket of every node cached), and takes the bucket-count
// sequence from the very policy object libstdc++ uses (std::__detail::_Prime_rehash_policy),
// so growth happens at the same sizes to the same prime counts.  tests/test_host.py compares it
// with std::unordered_set on random insert sequences.
#ifndef MCMC_B200_STD_ORDER_SET_H_
#define MCMC_B200_STD_ORDER_SET_H_

#include <cstddef>
#include <cstdint>
#include <unordered_set>
#include <vector>

namespace mcmc {

template <class Key>
class StdOrderSet {
 public:
  StdOrderSet() { Clear(); }

  void Clear() {
    keys_.clear();
    next_.clear();
    node_bucket_.clear();
    buckets_.assign(1, kEmpty);  // a default-constructed unordered_set has one (empty) bucket
    head_ = kNull;
    policy_ = std::__detail::_Prime_rehash_policy();
  }

  size_t size() const { return keys_.size(); }

  // true if k was not present
  bool Insert(Key k) {
    size_t b = Bucket(k, buckets_.size());
    if (!keys_.empty() && FindInBucket(b, k)) return false;
    // _M_need_rehash() starts with this very comparison and does nothing when it is false
    if (keys_.size() + 1 > policy_._M_next_resize) {
      const std::pair<bool, size_t> grow = policy_._M_need_rehash(buckets_.size(), keys_.size(), 1);
      if (grow.first) {
        Rehash(grow.second);
        b = Bucket(k, buckets_.size());
      }
    }
    const int32_t node = static_cast<int32_t>(keys_.size());
    keys_.push_back(k);
    node_bucket_.push_back(static_cast<uint32_t>(b));
    if (buckets_[b] != kEmpty) {  // goes to the front of its bucket's run
      const int32_t prev = buckets_[b];
      next_.push_back(Next(prev));
      SetNext(prev, node);
    } else {  // an empty bucket starts at the very front of the list
      next_.push_back(head_);
      head_ = node;
      if (next_[node] != kNull) buckets_[node_bucket_[next_[node]]] = node;
      buckets_[b] = kBeforeBegin;
    }
    return true;
  }

  // elements in iteration order, inserted at the front of *out (as vector::insert(begin(), ...))
  template <class T>
  void EmitTo(std::vector<T>* out) const {
    const size_t old = out->size();
    out->resize(old + keys_.size());
    if (old) std::move_backward(out->begin(), out->begin() + old, out->end());
    size_t i = 0;
    for (int32_t p = head_; p != kNull; p = next_[p]) (*out)[i++] = static_cast<T>(keys_[p]);
  }

 private:
  static constexpr int32_t kNull = -1, kEmpty = -1, kBeforeBegin = -2;

  // std::hash of an integer is the value itself; bucket = hash % count.  For 32-bit keys and
  // counts the 32-bit remainder is the same number and a cheaper divide.
  static size_t Bucket(Key k, size_t count) {
    if (sizeof(Key) <= 4 && count <= 0xffffffffu)
      return static_cast<uint32_t>(k) % static_cast<uint32_t>(count);
    return static_cast<size_t>(k) % count;
  }

  int32_t Next(int32_t prev) const { return prev == kBeforeBegin ? head_ : next_[prev]; }
  void SetNext(int32_t prev, int32_t node) {
    if (prev == kBeforeBegin) head_ = node; else next_[prev] = node;
  }

  bool FindInBucket(size_t b, Key k) const {
    if (buckets_[b] == kEmpty) return false;
    for (int32_t p = Next(buckets_[b]);; p = next_[p]) {
      if (keys_[p] == k) return true;
      const int32_t nx = next_[p];
      if (nx == kNull || node_bucket_[nx] != b) return false;
    }
  }

  void Rehash(size_t count) {
    scratch_.assign(count, kEmpty);
    int32_t p = head_;
    head_ = kNull;
    size_t front_bucket = 0;
    while (p != kNull) {
      const int32_t nx = next_[p];
      const size_t b = Bucket(keys_[p], count);
      node_bucket_[p] = static_cast<uint32_t>(b);
      if (scratch_[b] == kEmpty) {
        next_[p] = head_;
        head_ = p;
        scratch_[b] = kBeforeBegin;
        if (next_[p] != kNull) scratch_[front_bucket] = p;
        front_bucket = b;
      } else {
        const int32_t prev = scratch_[b];
        if (prev == kBeforeBegin) {
          next_[p] = head_;
          head_ = p;
        } else {
          next_[p] = next_[prev];
          next_[prev] = p;
        }
      }
      p = nx;
    }
    buckets_.swap(scratch_);
  }

  std::vector<Key> keys_;
  std::vector<int32_t> next_;
  std::vector<uint32_t> node_bucket_;
  std::vector<int32_t> buckets_, scratch_;
  int32_t head_;
  std::__detail::_Prime_rehash_policy policy_;
};

}  // namespace mcmc

#endif  // MCMC_B200_STD_ORDER_SET_H_
