#include "mcmc/learner.h"

#include <chrono>
#include <cmath>
#include <functional>
#include <random>

#include "mcmc/serialize.h"

using namespace std::chrono;

namespace mcmc {

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
      phase_(0) {
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
  sample->edges.clear();
  const Float weight = SampleMiniBatch(&sample->edges, &sample->seed);
  ExtractNodesFromMiniBatch(sample->edges, &sample->nodes_vec);
  if (sample->nodes_vec.empty()) throw BackendError("mini-batch size = 0!");
  if (sample->edges.size() > sample->dev_edges.GetSize() / sizeof(Edge) ||
      sample->nodes_vec.size() > sample->dev_nodes.GetSize() / sizeof(Vertex))
    throw BackendError("mini-batch exceeds the device buffers");
  sample->dev_edges.Write(sample->queue, sample->edges.size(), sample->edges.data());
  sample->dev_nodes.Write(sample->queue, sample->nodes_vec.size(), sample->nodes_vec.data());
  h2dBytes_ += sample->edges.size() * sizeof(Edge) + sample->nodes_vec.size() * sizeof(Vertex);
  sample->neighbor_sampler(static_cast<uint32_t>(sample->nodes_vec.size()), &sample->dev_nodes);
  return weight;
}

const Sample& Learner::PeekNextSample() {
  if (!futures_[phase_].valid() && !pendingValid_[phase_])
    futures_[phase_] = std::async(std::launch::async, &Learner::DoSample, this, &samples_[phase_]);
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

void Learner::Run(uint32_t max_iters, sig_atomic_t* signaled) {
  const auto t1 = high_resolution_clock::now();
  if (!futures_[phase_].valid() && !pendingValid_[phase_])
    futures_[phase_] = std::async(std::launch::async, &Learner::DoSample, this, &samples_[phase_]);
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
    futures_[1 - phase_] = std::async(std::launch::async, &Learner::DoSample, this, &samples_[1 - phase_]);
    samplingTime_ += duration_cast<nanoseconds>(high_resolution_clock::now() - ts).count();

    Sample& s = samples_[phase_];
    phiUpdater_(s.dev_nodes, s.neighbor_sampler.GetData(), static_cast<uint32_t>(s.nodes_vec.size()));
    betaUpdater_(&s.dev_edges, static_cast<uint32_t>(s.edges.size()), weight);
    edgesProcessed_ += s.edges.size();
    // the sampler thread reuses this buffer two iterations from now; drain before flipping
    queue_.Finish();
    phase_ = 1 - phase_;
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
  std::cerr << "ITERATIONS  : " << (stepCount_ - 1) << ", MINI-BATCH EDGES: " << edgesProcessed_ << std::endl;
}

// Record order of the reference (learner.cc:301-361): beta, theta, pi, phi, phi updater,
// beta updater, perplexity, LearnerProperties, sample 0, sample 1.
bool Learner::Serialize(std::ostream* out) {
  PeekNextSample();  // drain the in-flight sampler so its state is final
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
  futures_[0] = std::future<Float>();
  futures_[1] = std::future<Float>();
  pendingValid_[0] = pendingValid_[1] = false;
  pendingWeight_[phase_] = static_cast<Float>(props.weight);
  pendingValid_[phase_] = true;
  return true;
}

}  // namespace mcmc
