// mcmc/partitioned-alloc.h -- the pi matrix [rows, cols] (+ the phi row sums).
// The reference splits the matrix into <= 32 separately allocated row blocks to dodge the
// per-allocation limit (partitioned-alloc.h:14-29,73-141).  On a B200 a shard is one
// contiguous allocation in 180 GB of HBM; the same row -> (block, offset) rule is used one
// level up, to partition rows across the GPUs of a box (ammsb_store).  RowsPerBlock() keeps
// reporting the reference's block size because the checkpoint format records it.
#ifndef MCMC_B200_PARTITIONED_ALLOC_H_
#define MCMC_B200_PARTITIONED_ALLOC_H_

#include <memory>
#include <vector>

#include "mcmc/types.h"

namespace mcmc {

template <class T>
class RowPartitionedMatrixFactory;

template <class T>
class RowPartitionedMatrix {
 public:
  ~RowPartitionedMatrix() { ammsb_store_destroy(store_); }
  uint32_t Rows() const { return rows_; }
  uint32_t Cols() const { return cols_; }
  uint32_t RowsPerBlock() const { return rows_per_block_; }
  uint32_t NumBlocks() const { return (rows_ + rows_per_block_ - 1) / rows_per_block_; }
  ammsb_store* Get() const { return store_; }

  void ReadRows(uint64_t row0, uint64_t nrows, T* host) const { AmmsbCheck(ammsb_store_read_pi(store_, row0, nrows, host)); }
  void WriteRows(uint64_t row0, uint64_t nrows, const T* host) { AmmsbCheck(ammsb_store_write_pi(store_, row0, nrows, host)); }
  void ReadSums(uint64_t row0, uint64_t nrows, T* host) const { AmmsbCheck(ammsb_store_read_phi(store_, row0, nrows, host)); }
  void WriteSums(uint64_t row0, uint64_t nrows, const T* host) { AmmsbCheck(ammsb_store_write_phi(store_, row0, nrows, host)); }

 private:
  RowPartitionedMatrix(clcuda::Queue queue, uint32_t rows, uint32_t cols, uint32_t rows_in_block)
      : queue_(queue), rows_(rows), cols_(cols) {
    static_assert(sizeof(T) == sizeof(float), "the device store holds fp32");
    // reference CUDA build: 512 MiB blocks (partitioned-alloc.h:122-131)
    rows_per_block_ = rows_in_block ? rows_in_block
                                    : static_cast<uint32_t>((512ull << 20) / (static_cast<uint64_t>(cols) * sizeof(T)));
    AmmsbCheck(ammsb_store_create(queue_(), rows, cols, 1, 0, &store_));
  }
  clcuda::Queue queue_;
  uint32_t rows_, cols_, rows_per_block_;
  ammsb_store* store_ = nullptr;
  friend class RowPartitionedMatrixFactory<T>;
};

template <class T>
class RowPartitionedMatrixFactory : public std::enable_shared_from_this<RowPartitionedMatrixFactory<T>> {
 public:
  static std::shared_ptr<RowPartitionedMatrixFactory> New(clcuda::Queue queue) {
    return std::shared_ptr<RowPartitionedMatrixFactory>(new RowPartitionedMatrixFactory(queue));
  }
  RowPartitionedMatrix<T>* CreateMatrix(uint32_t rows, uint32_t cols, uint32_t rowsInBlock = 0) {
    return new RowPartitionedMatrix<T>(queue_, rows, cols, rowsInBlock);
  }

 private:
  explicit RowPartitionedMatrixFactory(clcuda::Queue queue) : queue_(queue) {}
  clcuda::Queue queue_;
};

}  // namespace mcmc

#endif  // MCMC_B200_PARTITIONED_ALLOC_H_
