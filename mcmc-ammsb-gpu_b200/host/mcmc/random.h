// mcmc/random.h -- per-work-item RNG state pools on the device.
// Reference: mcmc/random.h:13-79.  A pool of n states is seeded state[i] = (sx+i, sy+i);
// the xorshift128+/Ziggurat/gamma transforms live in csrc/common.cuh.
#ifndef MCMC_B200_RANDOM_H_
#define MCMC_B200_RANDOM_H_

#include <algorithm>
#include <memory>
#include <vector>

#include "mcmc/types.h"

namespace mcmc {

template <class T>
class RowPartitionedMatrix;

namespace random {

typedef ::mcmc::ulong2 gsl_rng;
typedef gsl_rng random_seed_t;

class OpenClRandomFactory;

class OpenClRandom {
 public:
  ~OpenClRandom();
  ammsb_rng* Get() const { return handle_; }
  uint64_t NumSeeds() const { return size_; }
  std::vector<random_seed_t> GetSeeds();            // device -> host copy of the pool
  void SetSeeds(const std::vector<random_seed_t>&);  // host -> device
  bool Serialize(std::ostream* out);
  bool Parse(std::istream* in);

 private:
  OpenClRandom(std::shared_ptr<OpenClRandomFactory> factory, clcuda::Queue queue, uint64_t size,
               random_seed_t seed);
  std::shared_ptr<OpenClRandomFactory> factory_;
  clcuda::Queue queue_;
  uint64_t size_;
  ammsb_rng* handle_ = nullptr;
  friend class OpenClRandomFactory;
};

class OpenClRandomFactory : public std::enable_shared_from_this<OpenClRandomFactory> {
 public:
  static std::shared_ptr<OpenClRandomFactory> New(clcuda::Queue queue) {
    return std::shared_ptr<OpenClRandomFactory>(new OpenClRandomFactory(queue));
  }
  OpenClRandom* CreateRandom(uint64_t size, random_seed_t seed) {
    return new OpenClRandom(shared_from_this(), queue_, size, seed);
  }

 private:
  explicit OpenClRandomFactory(clcuda::Queue queue) : queue_(queue) {}
  clcuda::Queue queue_;
};

// pi ~ Gamma(eta0, eta1) per element, rows normalised, row sums -> sum (reference
// random.cc:159-167: stream = pool rows*32 seeded {11,113}, one group of 32 per row)
void RandomGammaAndNormalize(clcuda::Queue* queue, Float eta0, Float eta1, RowPartitionedMatrix<Float>* norm,
                             clcuda::Buffer<Float>* sum);

// host generator -> base and its row-normalised copy norm (reference random.h:70-79)
void NormalizeRowsOnDevice(clcuda::Queue* queue, clcuda::Buffer<Float>* norm, uint32_t cols);
template <class Generator>
void RandomAndNormalize(clcuda::Queue* queue, Generator* gen, clcuda::Buffer<Float>* base,
                        clcuda::Buffer<Float>* norm, uint32_t cols) {
  std::vector<Float> host(base->GetSize() / sizeof(Float));
  std::generate(host.begin(), host.end(), *gen);
  base->Write(*queue, host.size(), host.data());
  norm->Write(*queue, host.size(), host.data());
  NormalizeRowsOnDevice(queue, norm, cols);
}

}  // namespace random
}  // namespace mcmc

#endif  // MCMC_B200_RANDOM_H_
