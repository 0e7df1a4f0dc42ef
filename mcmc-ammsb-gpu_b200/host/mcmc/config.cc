#include "mcmc/config.h"

#include <algorithm>
#include <cctype>
#include <sstream>

namespace mcmc {

namespace {
bool EqualsIgnoreCase(const std::string& a, const std::string& b) {
  return a.size() == b.size() &&
         std::equal(a.begin(), a.end(), b.begin(),
                    [](char x, char y) { return std::tolower((unsigned char)x) == std::tolower((unsigned char)y); });
}
// Float -> scientific text with 6 decimals + 'f' (what the reference hands to its JIT)
std::string FloatFlag(Float f) {
  std::ostringstream out;
  out << std::scientific << f << "f";
  return out.str();
}
}  // namespace

const std::string& GetSourceGuard() {
  static const std::string kNone;  // no JIT source to guard
  return kNone;
}

std::vector<std::string> MakeCompileFlags(const Config& cfg) {
  std::vector<std::string> flags = {"-DFLOAT_TYPE=float", "-DVERTEX_TYPE=uint", "-DEDGE_TYPE=ulong"};
  flags.push_back("-DK=" + std::to_string(cfg.K));
  flags.push_back("-DN=" + std::to_string(cfg.N));
  flags.push_back("-DE=" + std::to_string(cfg.E));
  flags.push_back("-DALPHA=" + FloatFlag(cfg.alpha));
  flags.push_back("-DEPS_A=" + FloatFlag(cfg.a));
  flags.push_back("-DEPS_B=" + FloatFlag(cfg.b));
  flags.push_back("-DEPS_C=" + FloatFlag(cfg.c));
  flags.push_back("-DEPSILON=" + FloatFlag(cfg.epsilon));
  flags.push_back("-DETA0=" + FloatFlag(cfg.eta0));
  flags.push_back("-DETA1=" + FloatFlag(cfg.eta1));
  flags.push_back("-DNUM_NEIGHBORS=" + std::to_string(cfg.num_node_sample));
  return flags;
}

ammsb_params MakeParams(const Config& cfg) {
  ammsb_params p;
  p.N = cfg.N;
  p.E = cfg.E;
  p.K = static_cast<uint32_t>(cfg.K);
  p.num_neighbors = static_cast<uint32_t>(cfg.num_node_sample);
  p.alpha = ammsb_round_param(cfg.alpha);
  p.a = ammsb_round_param(cfg.a);
  p.b = ammsb_round_param(cfg.b);
  p.c = ammsb_round_param(cfg.c);
  p.epsilon = ammsb_round_param(cfg.epsilon);
  p.eta0 = ammsb_round_param(cfg.eta0);
  p.eta1 = ammsb_round_param(cfg.eta1);
  return p;
}

ammsb_phi_opts MakePhiOpts(const Config& cfg) {
  ammsb_phi_opts o;
  o.mode = cfg.phi_mode == PHI_NODE_PER_THREAD ? AMMSB_MODE_THREAD : AMMSB_MODE_WG;
  o.wg = cfg.phi_wg_size;
  o.disable_noise = cfg.phi_disable_noise ? 1 : 0;
  o.strict = cfg.phi_strict ? 1 : 0;
  o.part_index = 0;  // single-GPU Learner: every slot
  o.part_count = 1;
  return o;
}

std::ostream& operator<<(std::ostream& out, const Config& cfg) {
  out << "Config:\n"
      << "heldout ratio: " << cfg.heldout_ratio << "\n"
      << "alpha: " << cfg.alpha << "\n"
      << "a: " << cfg.a << ", b: " << cfg.b << ", c: " << cfg.c << "\n"
      << "epsilon: " << cfg.epsilon << "\n"
      << "eta: (" << cfg.eta0 << ", " << cfg.eta1 << ")\n"
      << "K: " << cfg.K << "\n"
      << "m: " << cfg.mini_batch_size << "\n"
      << "n: " << cfg.num_node_sample << "\n"
      << "strategy: " << to_string(cfg.strategy) << "\n"
      << "ppx-wg: " << cfg.ppx_wg_size << "\n"
      << "phi-wg: " << cfg.phi_wg_size << "\n"
      << "beta-wg: " << cfg.beta_wg_size << "\n"
      << "phi-seed: " << cfg.phi_seed << "\n"
      << "beta-seed: " << cfg.beta_seed << "\n"
      << "neighbor-seed: " << cfg.neighbor_seed << "\n"
      << "|N|: " << cfg.N << "\n"
      << "|E|: " << cfg.E << "\n"
      << "phi_mode: " << to_string(cfg.phi_mode) << "\n"
      << "phi_vwidth: " << cfg.phi_vector_width << "\n";
  if (cfg.phi_mode == PHI_NODE_PER_WORKGROUP_CODE_GEN) {
    out << "phi_probs_shared: " << cfg.phi_probs_shared << "\n"
        << "phi_grads_shared: " << cfg.phi_grads_shared << "\n"
        << "phi_pi_shared: " << cfg.phi_pi_shared << "\n";
  }
  if (cfg.training) out << "|Training edges|: " << cfg.training->Size() << "\n";
  if (cfg.heldout) out << "|Heldout edges|: " << cfg.heldout->Size() << "\n";
  return out;
}

std::istream& operator>>(std::istream& in, PhiUpdaterMode& mode) {
  std::string token;
  in >> token;
  if (EqualsIgnoreCase(token, "THREAD")) mode = PHI_NODE_PER_THREAD;
  else if (EqualsIgnoreCase(token, "WG-NAIVE")) mode = PHI_NODE_PER_WORKGROUP_NAIVE;
  else if (EqualsIgnoreCase(token, "WG-SHARED")) mode = PHI_NODE_PER_WORKGROUP_SHARED;
  else if (EqualsIgnoreCase(token, "WG-GEN")) mode = PHI_NODE_PER_WORKGROUP_CODE_GEN;
  else throw std::invalid_argument("Invalid phi mode");
  return in;
}

std::string to_string(const PhiUpdaterMode& mode) {
  switch (mode) {
    case PHI_NODE_PER_THREAD: return "THREAD";
    case PHI_NODE_PER_WORKGROUP_NAIVE: return "WG-NAIVE";
    case PHI_NODE_PER_WORKGROUP_SHARED: return "WG-SHARED";
    case PHI_NODE_PER_WORKGROUP_CODE_GEN: return "WG-GEN";
  }
  throw std::invalid_argument("Invalid phi mode");
}

}  // namespace mcmc
