#include "mcmc/random.h"

#include "mcmc/partitioned-alloc.h"
#include "mcmc/serialize.h"

namespace mcmc {
namespace random {

OpenClRandom::OpenClRandom(std::shared_ptr<OpenClRandomFactory> factory, clcuda::Queue queue, uint64_t size,
                           random_seed_t seed)
    : factory_(factory), queue_(queue), size_(size) {
  AmmsbCheck(ammsb_rng_create(queue_(), size, seed[0], seed[1], &handle_));
}

OpenClRandom::~OpenClRandom() { ammsb_rng_destroy(handle_); }

std::vector<random_seed_t> OpenClRandom::GetSeeds() {
  std::vector<random_seed_t> host(size_);
  AmmsbCheck(ammsb_rng_get_state(handle_, reinterpret_cast<uint64_t*>(host.data())));
  return host;
}

void OpenClRandom::SetSeeds(const std::vector<random_seed_t>& host) {
  if (host.size() != size_) throw BackendError("RNG pool size mismatch");
  AmmsbCheck(ammsb_rng_set_state(handle_, reinterpret_cast<const uint64_t*>(host.data())));
}

bool OpenClRandom::Serialize(std::ostream* out) {
  std::vector<random_seed_t> host = GetSeeds();
  return SerializeBytes(out, host.data(), host.size() * sizeof(random_seed_t));
}

bool OpenClRandom::Parse(std::istream* in) {
  std::vector<random_seed_t> host(size_);
  if (!ParseBytes(in, host.data(), host.size() * sizeof(random_seed_t))) return false;
  SetSeeds(host);
  return true;
}

void RandomGammaAndNormalize(clcuda::Queue* queue, Float eta0, Float eta1, RowPartitionedMatrix<Float>* norm,
                             clcuda::Buffer<Float>* sum) {
  AmmsbCheck(ammsb_store_init_pi(norm->Get(), eta0, eta1));
  if (sum != nullptr) {
    float* d_phi = nullptr;
    AmmsbCheck(ammsb_store_local_ptrs(norm->Get(), nullptr, &d_phi));
    AmmsbCheck(ammsb_d2d((*queue)(), sum->data(), d_phi, sizeof(Float) * norm->Rows()));
  }
  queue->Finish();
}

void NormalizeRowsOnDevice(clcuda::Queue* queue, clcuda::Buffer<Float>* norm, uint32_t cols) {
  const uint32_t rows = static_cast<uint32_t>(norm->GetSize() / sizeof(Float) / cols);
  AmmsbCheck(ammsb_row_normalize((*queue)(), norm->data(), rows, cols, nullptr));
  queue->Finish();
}

}  // namespace random
}  // namespace mcmc
