#include "mcmc/perplexity.h"

#include "mcmc/serialize.h"

namespace mcmc {

namespace {
size_t PpxWorkspace(clcuda::Queue& q) {
  size_t bytes = 0;
  AmmsbCheck(ammsb_perplexity_workspace_bytes(q(), &bytes));
  return bytes;
}
}  // namespace

PerplexityCalculator::PerplexityCalculator(Mode, const Config& cfg, clcuda::Queue queue,
                                           clcuda::Buffer<Float>& beta, RowPartitionedMatrix<Float>* pi,
                                           clcuda::Buffer<Edge>& edges, OpenClSet* edgeSet,
                                           const std::vector<std::string>&, const std::string&)
    : queue_(queue),
      beta_(beta),
      pi_(pi),
      edges_(edges),
      edgeSet_(edgeSet),
      ppx_per_edge_(queue_.GetContext(), edges_.GetSize() / sizeof(Edge)),
      workspace_(queue_.GetContext(), PpxWorkspace(queue_)),
      params_(MakeParams(cfg)),
      count_calls_(0),
      sums_{0, 0, 0, 0},
      t_ppx_(0) {
  AmmsbCheck(ammsb_memset(queue_(), ppx_per_edge_.data(), 0, ppx_per_edge_.GetSize()));
  queue_.Finish();
}

Float PerplexityCalculator::operator()() {
  ++count_calls_;
  double avg = 0;
  AmmsbCheck(ammsb_timer_start(queue_()));
  AmmsbCheck(ammsb_perplexity(queue_(), &params_, pi_->Get(), beta_.data(), edgeSet_->Get(), edges_.data(),
                              static_cast<uint32_t>(edges_.GetSize() / sizeof(Edge)), ppx_per_edge_.data(),
                              count_calls_, sums_, &avg, workspace_.data(), workspace_.GetSize()));
  float ms = 0;
  AmmsbCheck(ammsb_timer_stop_ms(queue_(), &ms));
  t_ppx_ += ms;
  return static_cast<Float>(avg);
}

bool PerplexityCalculator::Serialize(std::ostream* out) {
  PerplexityProperties props;
  props.count_calls = count_calls_;
  props.ppx_time = t_ppx_;
  return SerializeMessage(out, props) && ::mcmc::Serialize(out, &ppx_per_edge_, &queue_);
}

bool PerplexityCalculator::Parse(std::istream* in) {
  PerplexityProperties props;
  if (!ParseMessage(in, &props)) return false;
  count_calls_ = props.count_calls;
  t_ppx_ = props.ppx_time;
  return ::mcmc::Parse(in, &ppx_per_edge_, &queue_);
}

}  // namespace mcmc
