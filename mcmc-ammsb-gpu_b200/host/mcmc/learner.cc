#include "mcmc/learner.h"

#include <chrono>
#include <cmath>
#include <functional>
#include <random>

#include "mcmc/serialize.h"

using namespace std::chrono;

namespace mcmc {

namespace {
// reference learner.cc:47-75: the first ratio*|training| training links, then
// links * (N(N-1)/2) / E random pairs that are in neither set (libc rand(), as the reference)
std::vector<Edge> MakeEdgesForTrainingPerplexity(const Config& cfg) {
  const uint64_t total = (cfg.N * (cfg.N - 1)) / 2;
  const uint64_t num_links = static_cast<uint64_t>(cfg.training_ppx_ratio * cfg.training_edges.size());
  const uint64_t num_non_links = static_cast<uint64_t>(num_links * total / static_cast<double>(cfg.E));
  std::vector<Edge> ret(num_links + num_non_links);
  std::copy(cfg.training_edges.begin(), cfg.training_edges.begin() + num_links, ret.begin());
  for (uint64_t i = num_links; i < ret.size(); ++i) {
    Edge e;
    do {
      const Vertex u = rand() % cfg.N;
      Vertex v;
      do {
        v = rand() % cfg.N;
      } while (u == v);
      e = MakeEdge(u, v);  // as the reference: not canonicalised (u may exceed v)
    } while (cfg.training->Has(e) || cfg.heldout->Has(e));
    ret[i] = e;
  }
  return ret;
}
}  // namespace

namespace {
typedef Float (*SamplerFn)(const Config&, std::vector<Edge>*, unsigned int*);
SamplerFn PickSampler(SampleStrategy s) {
  switch (s) {
    case NodeLink: return sampleNodeLink;
    case NodeNonLink: return sampleNodeNonLink;
    case Node: return sampleNode;
    case BFLink: return sampleBreadthFirstLink;
    case BFNonLink: return sampleBreadthFirstNonLink;
    case BF: return sampleBreadthFirst;
  }
  throw std::invalid_argument("Unknown sample strategy");
}
}  // namespace

Learner::Learner(const Config& cfg, clcuda::Queue queue)
    : cfg_(cfg),
      queue_(queue),
      beta_(queue_.GetContext(), 2 * cfg_.K),
      theta_(queue_.GetContext(), 2 * cfg_.K),
      allocFactory_(RowPartitionedMatrixFactory<Float>::New(queue_)),
      pi_(allocFactory_->CreateMatrix(cfg_.N, cfg_.K)),
      phi_(queue_.GetContext(), cfg_.N),
      setFactory_(OpenClSetFactory::New(queue_)),
      trainingSet_(setFactory_->CreateSet(*cfg_.training)),
      heldoutSet_(setFactory_->CreateSet(*cfg_.heldout)),
      trainingEdges_(queue_.GetContext(), queue_, cfg_.training_edges.begin(), cfg_.training_edges.end()),
      heldoutEdges_(queue_.GetContext(), queue_, cfg_.heldout_edges.begin(), cfg_.heldout_edges.end()),
      compileFlags_(MakeCompileFlags(cfg_)),
      heldoutPerplexity_(PerplexityCalculator::EDGE_PER_WORKGROUP, cfg_, queue_, beta_, pi_.get(), heldoutEdges_,
                         heldoutSet_.get(), compileFlags_),
      phiUpdater_(cfg_, queue_, beta_, pi_.get(), phi_, trainingSet_.get(), compileFlags_),
      betaUpdater_(BetaUpdater::EDGE_PER_WORKGROUP, cfg_, queue_, theta_, beta_, pi_.get(), trainingSet_.get(),
                   compileFlags_),
      sampler_(PickSampler(cfg_.strategy)),
      stepCount_(1),
      time_(0),
      samplingTime_(0),
      edgesProcessed_(0),
      phase_(0),
      betaMirror_(nullptr) {
  if (cfg_.calc_train_ppx) {  // before the Samples draw their seeds, as in the reference's member order
    trainingPerplexityEdges_ = MakeEdgesForTrainingPerplexity(cfg_);
    devTrainingPerplexityEdges_.reset(new clcuda::Buffer<Edge>(queue_.GetContext(), queue_,
                                                               trainingPerplexityEdges_.begin(),
                                                               trainingPerplexityEdges_.end()));
    trainingPerplexity_.reset(new PerplexityCalculator(PerplexityCalculator::EDGE_PER_WORKGROUP, cfg_, queue_, beta_,
                                                       pi_.get(), *devTrainingPerplexityEdges_, trainingSet_.get(),
                                                       compileFlags_));
  }
  for (auto& ev : iterDone_) AmmsbCheck(ammsb_event_create(queue_(), &ev));
  std::shared_ptr<DeviceStrategyData> on_device;
  if (cfg_.device_sampler)
    on_device.reset(new DeviceStrategyData(cfg_, queue_, trainingSet_->Get(), heldoutSet_->Get()));
  for (auto& sample : samples_) {
    sample.reset(new Sample(cfg_, queue_));  // seed = rand(), as in the reference (sample.cc:132)
    if (on_device) sample->StartOnDevice(cfg_.strategy, on_device, &stats_);
    else sample->Start(sampler_, &stats_);
  }
  // phi lives in a Buffer of its own in the reference; the store adopts that memory
  AmmsbCheck(ammsb_store_bind_phi(pi_->Get(), phi_.data()));
  // theta ~ Gamma(eta0, eta1) on the host, fixed seed; beta = theta row-normalised
  // (reference learner.cc:149-153 -- same libstdc++ engine and distribution)
  std::mt19937 engine(6342455113);
  std::gamma_distribution<Float> gamma_distribution(cfg_.eta0, cfg_.eta1);
  auto gamma = std::bind(gamma_distribution, engine);
  random::RandomAndNormalize(&queue_, &gamma, &theta_, &beta_, 2);
  // pi ~ Gamma on the device, row-normalised; phi = row sums (learner.cc:154-155)
  random::RandomGammaAndNormalize(&queue_, cfg_.eta0, cfg_.eta1, pi_.get(), nullptr);
}

Learner::~Learner() {
  for (auto& ev : iterDone_) ammsb_event_destroy(ev);
}

Float Learner::SampleMiniBatch(std::vector<Edge>* edges, unsigned int* seed) { return sampler_(cfg_, edges, seed); }

const SampleSlot& Learner::PeekNextSample() { return samples_[phase_]->WaitReady(); }

namespace {
// mini-batches of `stream` that are launched but not yet retired
uint64_t CountInFlight(Sample* const* owner, uint64_t launched, uint64_t retired, const Sample* stream) {
  uint64_t n = 0;
  for (uint64_t k = retired; k < launched; ++k)
    if (owner[k % 3] == stream) ++n;
  return n;
}
}  // namespace

Float Learner::HeldoutPerplexity() {
  const auto t1 = high_resolution_clock::now();
  const Float avg = heldoutPerplexity_();
  time_ += duration_cast<nanoseconds>(high_resolution_clock::now() - t1).count();
  return std::exp(avg);
}

Float Learner::TrainingPerplexity() {
  if (!trainingPerplexity_) throw BackendError("TrainingPerplexity() needs Config::calc_train_ppx");
  const auto t1 = high_resolution_clock::now();
  const Float avg = (*trainingPerplexity_)();
  time_ += duration_cast<nanoseconds>(high_resolution_clock::now() - t1).count();
  return std::exp(avg);
}

void Learner::Run(uint32_t max_iters, sig_atomic_t* signaled) {
  const auto t1 = high_resolution_clock::now();
  // This call consumes max_iters mini-batches, alternating between the two streams starting
  // with `phase_`, and -- like the reference, which always has the next one in flight
  // (learner.cc:228) -- leaves exactly one more drawn.  That is all the streams may draw.
  const uint64_t total = static_cast<uint64_t>(max_iters) + 1;
  samples_[phase_]->Allow((total + 1) / 2);
  samples_[1 - phase_]->Allow(total / 2);
  // The stream stays fed: up to kInFlight iterations are enqueued before the host waits for the
  // oldest one (the reference drains the queue after every kernel).  A slot returns to its
  // sampler stream when the iteration that read it has completed.
  Sample* owner[kInFlight] = {nullptr};
  uint64_t launched = 0, retired = 0;
  auto retire_oldest = [&]() {
    const auto td = high_resolution_clock::now();
    AmmsbCheck(ammsb_event_sync(iterDone_[retired % kInFlight]));
    tDrain_ += duration_cast<nanoseconds>(high_resolution_clock::now() - td).count();
    owner[retired % kInFlight]->Release();
    ++retired;
  };
  for (uint64_t i = 0; i < max_iters && (signaled == nullptr || !*signaled); ++i, ++stepCount_) {
    const auto ts = high_resolution_clock::now();
    Sample& stream = *samples_[phase_];
    // WaitReady() returns the oldest unreleased mini-batch of the stream: the ones still in
    // flight on the GPU must be retired first if they belong to this stream's ring position
    while (launched - retired >= static_cast<uint64_t>(kInFlight - 1)) retire_oldest();
    SampleSlot& s = stream.WaitReady(/*skip=*/CountInFlight(owner, launched, retired, &stream));
    const auto tk = high_resolution_clock::now();
    samplingTime_ += duration_cast<nanoseconds>(tk - ts).count();

    phiUpdater_(s.dev_nodes, s.neighbors, static_cast<uint32_t>(s.nodes_vec.size()));
    betaUpdater_(&s.dev_edges, static_cast<uint32_t>(s.edges.size()), s.weight);
    edgesProcessed_ += s.edges.size();
    if (betaMirror_ != nullptr) beta_.ReadAsync(queue_, 2 * cfg_.K, betaMirror_);
    AmmsbCheck(ammsb_event_record(queue_(), iterDone_[launched % kInFlight]));
    owner[launched % kInFlight] = &stream;
    ++launched;
    tKernelsHost_ += duration_cast<nanoseconds>(high_resolution_clock::now() - tk).count();
    phase_ = 1 - phase_;
  }
  while (retired < launched) retire_oldest();
  time_ += duration_cast<nanoseconds>(high_resolution_clock::now() - t1).count();
}

void Learner::PrintStats() {
  const double total_s = time_ / 1.0e9;
  auto line = [&](const char* name, double seconds) {
    std::cerr << name << seconds << " (%" << (total_s > 0 ? 100 * seconds / total_s : 0) << ")" << std::endl;
  };
  std::cerr << "TOTAL    : " << total_s << std::endl;
  line("PPX CALC : ", heldoutPerplexity_.PerplexityTime() / 1.0e3);
  line("PPX ACCUM: ", heldoutPerplexity_.AccumulateTime() / 1.0e3);
  line("SAMPLING : ", samplingTime_ / 1.0e9);
  line("PHI      : ", phiUpdater_.UpdatePhiTime() / 1.0e3);
  line("PI       : ", phiUpdater_.UpdatePiTime() / 1.0e3);
  line("THETA SUM   : ", betaUpdater_.ThetaSumTime() / 1.0e3);
  line("GRADS PAR   : ", betaUpdater_.GradsPartialTime() / 1.0e3);
  line("GRADS SUM   : ", betaUpdater_.GradsSumTime() / 1.0e3);
  line("UPDATE THETA: ", betaUpdater_.UpdateThetaTime() / 1.0e3);
  line("NORM THETA  : ", betaUpdater_.NormalizeTime() / 1.0e3);
  line("  sampler threads: strategy ", stats_.strategy / 1.0e9);
  line("  sampler threads: extract  ", stats_.extract / 1.0e9);
  line("  sampler threads: H2D      ", stats_.copy / 1.0e9);
  line("  sampler threads: neighbors", stats_.neighbors / 1.0e9);
  line("  main thread: launches     ", tKernelsHost_ / 1.0e9);
  line("  main thread: drain        ", tDrain_ / 1.0e9);
  std::cerr << "ITERATIONS  : " << (stepCount_ - 1) << ", MINI-BATCH EDGES: " << edgesProcessed_ << std::endl;
}

// Record order of the reference (learner.cc:301-361): beta, theta, pi, phi, phi updater,
// beta updater, perplexity, LearnerProperties, sample 0, sample 1.
bool Learner::Serialize(std::ostream* out) {
  // Drain the sampler streams.  After a Run() that ran to completion the state is the
  // reference's: the next mini-batch (stream `phase_`) drawn and pending, the other stream's
  // latest mini-batch consumed.  (A Run() cut short by `signaled` may leave more mini-batches
  // drawn than the reference's format can carry; those are redrawn differently after Parse.)
  samples_[phase_]->Allow(1);
  samples_[phase_]->WaitReady();
  samples_[0]->Quiesce();
  samples_[1]->Quiesce();
  LearnerProperties props;
  props.stepCount = stepCount_;
  props.time = time_;
  props.samplingTime = samplingTime_;
  props.phase = phase_;
  props.weight = samples_[phase_]->WaitReady().weight;
  return ::mcmc::Serialize(out, &beta_, &queue_) && ::mcmc::Serialize(out, &theta_, &queue_) &&
         SerializeRpm(out, pi_.get()) && ::mcmc::Serialize(out, &phi_, &queue_) && phiUpdater_.Serialize(out) &&
         betaUpdater_.Serialize(out) && (!trainingPerplexity_ || trainingPerplexity_->Serialize(out)) &&
         heldoutPerplexity_.Serialize(out) && SerializeMessage(out, props) &&
         samples_[0]->Serialize(out) && samples_[1]->Serialize(out);
}

bool Learner::Parse(std::istream* in) {
  LearnerProperties props;
  if (!(::mcmc::Parse(in, &beta_, &queue_) && ::mcmc::Parse(in, &theta_, &queue_) && ParseRpm(in, pi_.get()) &&
        ::mcmc::Parse(in, &phi_, &queue_) && phiUpdater_.Parse(in) && betaUpdater_.Parse(in) &&
        (!trainingPerplexity_ || trainingPerplexity_->Parse(in)) && heldoutPerplexity_.Parse(in) &&
        ParseMessage(in, &props)))
    return false;
  if (props.phase != 0 && props.phase != 1) return false;  // indexes samples_[]
  stepCount_ = props.stepCount;
  time_ = props.time;
  samplingTime_ = props.samplingTime;
  phase_ = props.phase;
  // the stream of `phase_` holds the pending next mini-batch, the other one a consumed one
  if (!(samples_[0]->Parse(in, phase_ == 0) && samples_[1]->Parse(in, phase_ == 1))) return false;
  samples_[phase_]->WaitReady().weight = static_cast<Float>(props.weight);
  return true;
}

}  // namespace mcmc
