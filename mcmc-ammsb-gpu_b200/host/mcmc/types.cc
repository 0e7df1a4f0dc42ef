#include "mcmc/types.h"

namespace mcmc {

uint32_t GetMaxGroups() { return 65535; }

std::ostream& operator<<(std::ostream& out, const ulong2& v) { return out << v[0] << "," << v[1]; }

// "a,b".  (The reference stores the second number into v[2] -- out of bounds,
// types.cc:551 -- which makes its --phi-seed/--beta-seed/--neighbor-seed flags unusable;
// this parser accepts the documented syntax.)
std::istream& operator>>(std::istream& in, ulong2& v) {
  in >> v[0];
  if (in.get() != ',') {
    in.setstate(std::ios::failbit);
    throw std::invalid_argument("Invalid ulong2");
  }
  in >> v[1];
  return in;
}

namespace clcuda {

static std::shared_ptr<ammsb_ctx> NewCtx(int ordinal) {
  ammsb_ctx* c = nullptr;
  AmmsbCheck(ammsb_ctx_create(ordinal, &c));
  return std::shared_ptr<ammsb_ctx>(c, [](ammsb_ctx* p) { ammsb_ctx_destroy(p); });
}

std::string Device::Name() const {
  Context tmp(*this);
  char buf[256];
  AmmsbCheck(ammsb_ctx_device_name(tmp.get(), buf, sizeof buf));
  return buf;
}

Context::Context(const Device& dev) : impl_(NewCtx(dev.Ordinal())), ordinal_(dev.Ordinal()) {}

Queue::Queue(const Context& ctx, const Device&) : ctx_(ctx), stream_(NewCtx(ctx.DeviceOrdinal())) {}

}  // namespace clcuda
}  // namespace mcmc
