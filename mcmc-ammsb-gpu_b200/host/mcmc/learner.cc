#include "mcmc/learner.h"

#include <chrono>
#include <cmath>
#include <functional>
#include <random>

#include "mcmc/serialize.h"

using namespace std::chrono;

namespace mcmc {

SamplerThread::SamplerThread() : thread_(&SamplerThread::Loop, this) {}

SamplerThread::~SamplerThread() {
  {
    std::unique_lock<std::mutex> lock(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  thread_.join();
}

void SamplerThread::Loop() {
  std::unique_lock<std::mutex> lock(mu_);
  for (;;) {
    cv_.wait(lock, [this] { return has_task_ || stop_; });
    if (has_task_) {
      std::function<Float()> task = std::move(task_);
      has_task_ = false;
      lock.unlock();
      Float r = 0;
      std::exception_ptr err;
      try {
        r = task();
      } catch (...) {
        err = std::current_exception();
      }
      lock.lock();
      result_ = r;
      error_ = err;
      done_ = true;
      cv_.notify_all();
    } else if (stop_) {
      return;
    }
  }
}

void SamplerThread::Launch(std::function<Float()> task) {
  wait();
  {
    std::unique_lock<std::mutex> lock(mu_);
    task_ = std::move(task);
    has_task_ = true;
    done_ = false;
    valid_ = true;
  }
  cv_.notify_all();
}

void SamplerThread::wait() {
  if (!valid_) return;
  std::unique_lock<std::mutex> lock(mu_);
  cv_.wait(lock, [this] { return done_; });
}

Float SamplerThread::get() {
  if (!valid_) throw BackendError("SamplerThread::get() without a launched task");
  wait();
  valid_ = false;
  if (error_) {
    std::exception_ptr e = error_;
    error_ = nullptr;
    std::rethrow_exception(e);
  }
  return result_;
}

void SamplerThread::Reset() {
  wait();
  valid_ = false;
  error_ = nullptr;
}

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
      h2dBytes_(0),
      samples_{Sample(cfg_, queue_), Sample(cfg_, queue_)},
      pendingWeight_{0, 0},
      pendingValid_{false, false},
      phase_(0),
      betaMirror_(nullptr) {
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
  for (auto& f : futures_)
    if (f.valid()) f.wait();
}

Float Learner::SampleMiniBatch(std::vector<Edge>* edges, unsigned int* seed) { return sampler_(cfg_, edges, seed); }

// host mini-batch -> device copies -> neighbor sampling, all on the Sample's own queue so
// that it overlaps the previous iteration's kernels (reference learner.cc:175-194)
Float Learner::DoSample(Sample* sample) {
  const auto t0 = high_resolution_clock::now();
  sample->edges.clear();
  const Float weight = SampleMiniBatch(&sample->edges, &sample->seed);
  const auto t1 = high_resolution_clock::now();
  ExtractNodesFromMiniBatch(sample->edges, &sample->nodes_vec);
  const auto t2 = high_resolution_clock::now();
  if (sample->nodes_vec.empty()) throw BackendError("mini-batch size = 0!");
  if (sample->edges.size() > sample->dev_edges.GetSize() / sizeof(Edge) ||
      sample->nodes_vec.size() > sample->dev_nodes.GetSize() / sizeof(Vertex))
    throw BackendError("mini-batch exceeds the device buffers");
  sample->dev_edges.Write(sample->queue, sample->edges.size(), sample->edges.data());
  sample->dev_nodes.Write(sample->queue, sample->nodes_vec.size(), sample->nodes_vec.data());
  h2dBytes_ += sample->edges.size() * sizeof(Edge) + sample->nodes_vec.size() * sizeof(Vertex);
  const auto t3 = high_resolution_clock::now();
  sample->neighbor_sampler(static_cast<uint32_t>(sample->nodes_vec.size()), &sample->dev_nodes);
  const auto t4 = high_resolution_clock::now();
  tStrategy_ += duration_cast<nanoseconds>(t1 - t0).count();
  tExtract_ += duration_cast<nanoseconds>(t2 - t1).count();
  tCopy_ += duration_cast<nanoseconds>(t3 - t2).count();
  tNeighbor_ += duration_cast<nanoseconds>(t4 - t3).count();
  return weight;
}

const Sample& Learner::PeekNextSample() {
  LaunchSampler(phase_);
  if (!pendingValid_[phase_]) {
    pendingWeight_[phase_] = futures_[phase_].get();
    pendingValid_[phase_] = true;
  }
  return samples_[phase_];
}

Float Learner::HeldoutPerplexity() {
  const auto t1 = high_resolution_clock::now();
  const Float avg = heldoutPerplexity_();
  time_ += duration_cast<nanoseconds>(high_resolution_clock::now() - t1).count();
  return std::exp(avg);
}

void Learner::LaunchSampler(int buffer) {
  if (futures_[buffer].valid() || pendingValid_[buffer]) return;  // already drawn / being drawn
  Sample* sample = &samples_[buffer];
  futures_[buffer].Launch([this, sample] { return DoSample(sample); });
}

void Learner::Run(uint32_t max_iters, sig_atomic_t* signaled) {
  const auto t1 = high_resolution_clock::now();
  LaunchSampler(phase_);
  for (uint64_t i = 0; i < max_iters && (signaled == nullptr || !*signaled); ++i, ++stepCount_) {
    const auto ts = high_resolution_clock::now();
    Float weight;
    if (pendingValid_[phase_]) {
      weight = pendingWeight_[phase_];
      pendingValid_[phase_] = false;
    } else {
      weight = futures_[phase_].get();
    }
    // kernels of iteration t still read samples_[phase_]; the other buffer is free
    // (reference learner.cc:228: mini-batch t+1 is drawn while t is processed)
    LaunchSampler(1 - phase_);
    samplingTime_ += duration_cast<nanoseconds>(high_resolution_clock::now() - ts).count();

    const auto tk = high_resolution_clock::now();
    Sample& s = samples_[phase_];
    phiUpdater_(s.dev_nodes, s.neighbor_sampler.GetData(), static_cast<uint32_t>(s.nodes_vec.size()));
    betaUpdater_(&s.dev_edges, static_cast<uint32_t>(s.edges.size()), weight);
    edgesProcessed_ += s.edges.size();
    if (betaMirror_ != nullptr) beta_.ReadAsync(queue_, 2 * cfg_.K, betaMirror_);
    // the sampler thread reuses this buffer two iterations from now; drain before flipping
    const auto td = high_resolution_clock::now();
    queue_.Finish();
    tKernelsHost_ += duration_cast<nanoseconds>(td - tk).count();
    tDrain_ += duration_cast<nanoseconds>(high_resolution_clock::now() - td).count();
    const int consumed = phase_;
    phase_ = 1 - phase_;
    // The two Samples draw from independent seeds and RNG pools and sampling never reads
    // the model, so mini-batch t+2 can be drawn into the buffer that has just been consumed
    // while t+1 is still being drawn: same mini-batches, two sampler threads in flight.
    // Only when iteration t+2 belongs to this Run() call, so that a caller (and Serialize)
    // always finds the reference's state on return: exactly one mini-batch in flight.
    if (i + 2 < max_iters && (signaled == nullptr || !*signaled)) LaunchSampler(consumed);
  }
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
  line("  sampler threads: strategy ", tStrategy_ / 1.0e9);
  line("  sampler threads: extract  ", tExtract_ / 1.0e9);
  line("  sampler threads: H2D      ", tCopy_ / 1.0e9);
  line("  sampler threads: neighbors", tNeighbor_ / 1.0e9);
  line("  main thread: launches     ", tKernelsHost_ / 1.0e9);
  line("  main thread: drain        ", tDrain_ / 1.0e9);
  std::cerr << "ITERATIONS  : " << (stepCount_ - 1) << ", MINI-BATCH EDGES: " << edgesProcessed_ << std::endl;
}

// Record order of the reference (learner.cc:301-361): beta, theta, pi, phi, phi updater,
// beta updater, perplexity, LearnerProperties, sample 0, sample 1.
bool Learner::Serialize(std::ostream* out) {
  PeekNextSample();  // drain the in-flight sampler so its state is final
  if (futures_[1 - phase_].valid()) {  // only after a Run() cut short by `signaled`
    pendingWeight_[1 - phase_] = futures_[1 - phase_].get();
    pendingValid_[1 - phase_] = true;
  }
  LearnerProperties props;
  props.stepCount = stepCount_;
  props.time = time_;
  props.samplingTime = samplingTime_;
  props.phase = phase_;
  props.weight = pendingWeight_[phase_];
  return ::mcmc::Serialize(out, &beta_, &queue_) && ::mcmc::Serialize(out, &theta_, &queue_) &&
         SerializeRpm(out, pi_.get()) && ::mcmc::Serialize(out, &phi_, &queue_) && phiUpdater_.Serialize(out) &&
         betaUpdater_.Serialize(out) && heldoutPerplexity_.Serialize(out) && SerializeMessage(out, props) &&
         samples_[0].Serialize(out) && samples_[1].Serialize(out);
}

bool Learner::Parse(std::istream* in) {
  for (auto& f : futures_)
    if (f.valid()) f.wait();
  LearnerProperties props;
  if (!(::mcmc::Parse(in, &beta_, &queue_) && ::mcmc::Parse(in, &theta_, &queue_) && ParseRpm(in, pi_.get()) &&
        ::mcmc::Parse(in, &phi_, &queue_) && phiUpdater_.Parse(in) && betaUpdater_.Parse(in) &&
        heldoutPerplexity_.Parse(in) && ParseMessage(in, &props)))
    return false;
  stepCount_ = props.stepCount;
  time_ = props.time;
  samplingTime_ = props.samplingTime;
  phase_ = props.phase;
  if (!(samples_[0].Parse(in) && samples_[1].Parse(in))) return false;
  futures_[0].Reset();
  futures_[1].Reset();
  pendingValid_[0] = pendingValid_[1] = false;
  pendingWeight_[phase_] = static_cast<Float>(props.weight);
  pendingValid_[phase_] = true;
  return true;
}

}  // namespace mcmc
