#include "mcmc/beta.h"

#include "mcmc/serialize.h"

namespace mcmc {

namespace {
size_t BetaWorkspace(clcuda::Queue& q, uint32_t K) {
  size_t bytes = 0;
  AmmsbCheck(ammsb_beta_workspace_bytes(q(), K, &bytes));
  return bytes;
}
}  // namespace

BetaUpdater::BetaUpdater(Mode, const Config& cfg, clcuda::Queue queue, clcuda::Buffer<Float>& theta,
                         clcuda::Buffer<Float>& beta, RowPartitionedMatrix<Float>* pi, OpenClSet* trainingSet,
                         const std::vector<std::string>&, const std::string&)
    : cfg_(cfg),
      queue_(queue),
      theta_(theta),
      beta_(beta),
      pi_(pi),
      trainingSet_(trainingSet),
      randFactory_(random::OpenClRandomFactory::New(queue_)),
      rand_(randFactory_->CreateRandom(cfg.K, random::random_seed_t{cfg.beta_seed[0], cfg.beta_seed[1]})),
      params_(MakeParams(cfg)),
      count_calls_(0),
      theta_sum_(queue_.GetContext(), cfg.K),
      grads_(queue_.GetContext(), 2 * cfg.K),
      workspace_(queue_.GetContext(), BetaWorkspace(queue_, static_cast<uint32_t>(cfg.K))),
      t_grads_(0),
      t_update_theta_(0) {}

void BetaUpdater::operator()(clcuda::Buffer<Edge>* edges, uint32_t num_edges, Float scale) {
  ++count_calls_;
  ammsb_ctx* c = queue_();
  float ms = 0;
  if (!cfg_.stage_timers) {  // one call: the theta step rides on the gradient reduction
    AmmsbCheck(ammsb_update_beta(c, &params_, theta_.data(), beta_.data(), pi_->Get(), trainingSet_->Get(),
                                 edges->data(), num_edges, scale, count_calls_, rand_->Get(), theta_sum_.data(),
                                 grads_.data(), workspace_.data(), workspace_.GetSize()));
    return;
  }
  if (cfg_.stage_timers) AmmsbCheck(ammsb_timer_start(c));
  AmmsbCheck(ammsb_beta_grads(c, &params_, theta_.data(), beta_.data(), pi_->Get(), trainingSet_->Get(),
                              edges->data(), num_edges, theta_sum_.data(), grads_.data(), workspace_.data(),
                              workspace_.GetSize()));
  if (cfg_.stage_timers) {
    AmmsbCheck(ammsb_timer_stop_ms(c, &ms));
    t_grads_ += ms;
    AmmsbCheck(ammsb_timer_start(c));
  }
  AmmsbCheck(ammsb_update_theta(c, &params_, theta_.data(), beta_.data(), grads_.data(), scale, count_calls_,
                                rand_->Get()));
  if (cfg_.stage_timers) {
    AmmsbCheck(ammsb_timer_stop_ms(c, &ms));
    t_update_theta_ += ms;
  }
}

bool BetaUpdater::Serialize(std::ostream* out) {
  BetaProperties props;
  props.count_calls = count_calls_;
  props.grads_partial_time = t_grads_;
  props.update_theta_time = t_update_theta_;
  return rand_->Serialize(out) && ::mcmc::Serialize(out, &theta_sum_, &queue_) && SerializeMessage(out, props);
}

bool BetaUpdater::Parse(std::istream* in) {
  BetaProperties props;
  if (!(rand_->Parse(in) && ::mcmc::Parse(in, &theta_sum_, &queue_) && ParseMessage(in, &props))) return false;
  count_calls_ = props.count_calls;
  t_grads_ = props.grads_partial_time;
  t_update_theta_ = props.update_theta_time;
  return true;
}

}  // namespace mcmc
