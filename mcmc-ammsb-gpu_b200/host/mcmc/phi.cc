#include "mcmc/phi.h"

#include "mcmc/serialize.h"

namespace mcmc {

PhiUpdater::PhiUpdater(const Config& cfg, clcuda::Queue queue, clcuda::Buffer<Float>& beta,
                       RowPartitionedMatrix<Float>* pi, clcuda::Buffer<Float>& phi, OpenClSet* trainingSet,
                       const std::vector<std::string>&, const std::string&)
    : cfg_(cfg),
      queue_(queue),
      beta_(beta),
      pi_(pi),
      phi_(phi),
      phi_vec_(queue_.GetContext(), MaxMiniBatchNodes(cfg) * cfg.K),
      phi_sum_(queue_.GetContext(), MaxMiniBatchNodes(cfg)),
      trainingSet_(trainingSet),
      randFactory_(random::OpenClRandomFactory::New(queue_)),
      // one state per work-item of the reference launch (phi.cc:625-629)
      rand_(randFactory_->CreateRandom(
          MaxMiniBatchNodes(cfg) * (cfg.phi_mode == PHI_NODE_PER_THREAD ? 1 : cfg.phi_wg_size),
          random::random_seed_t{cfg.phi_seed[0], cfg.phi_seed[1]})),
      params_(MakeParams(cfg)),
      opts_(MakePhiOpts(cfg)),
      count_calls_(0),
      t_update_phi_(0),
      t_update_pi_(0) {}

void PhiUpdater::operator()(clcuda::Buffer<Vertex>& mini_batch_nodes, clcuda::Buffer<Vertex>& neighbors,
                            uint32_t num_mini_batch_nodes) {
  if (num_mini_batch_nodes == 0) throw BackendError("mini-batch nodes size = 0!");
  ++count_calls_;  // the operator's own 1-based step counter (phi.cc:739,754)
  ammsb_ctx* c = queue_();
  float ms = 0;
  if (cfg_.stage_timers) AmmsbCheck(ammsb_timer_start(c));
  AmmsbCheck(ammsb_update_phi(c, &params_, &opts_, beta_.data(), pi_->Get(), trainingSet_->Get(),
                              mini_batch_nodes.data(), neighbors.data(), num_mini_batch_nodes, count_calls_,
                              rand_->Get(), phi_vec_.data(), phi_sum_.data()));
  if (cfg_.stage_timers) {
    AmmsbCheck(ammsb_timer_stop_ms(c, &ms));
    t_update_phi_ += ms;
    AmmsbCheck(ammsb_timer_start(c));
  }
  AmmsbCheck(ammsb_update_pi(c, params_.K, pi_->Get(), phi_vec_.data(), phi_sum_.data(), mini_batch_nodes.data(),
                             num_mini_batch_nodes));
  if (cfg_.stage_timers) {
    AmmsbCheck(ammsb_timer_stop_ms(c, &ms));
    t_update_pi_ += ms;
  }
}

bool PhiUpdater::Serialize(std::ostream* out) {
  PhiProperties props;
  props.count_calls = count_calls_;
  props.update_phi_time = t_update_phi_;
  props.update_pi_time = t_update_pi_;
  return rand_->Serialize(out) && SerializeMessage(out, props);
}

bool PhiUpdater::Parse(std::istream* in) {
  PhiProperties props;
  if (!(rand_->Parse(in) && ParseMessage(in, &props))) return false;
  count_calls_ = props.count_calls;
  t_update_phi_ = props.update_phi_time;
  t_update_pi_ = props.update_pi_time;
  return true;
}

}  // namespace mcmc
