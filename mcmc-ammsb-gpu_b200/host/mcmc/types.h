// mcmc/types.h -- scalar types, edge encoding and the backend handle of the drop-in
// host API.  Mirrors the public surface of the reference's mcmc/types.h:31-74; the
// CLCudaAPI backend (namespace alias mcmc::clcuda, types.h:29) is replaced by thin
// RAII stand-ins over the C ABI of include/ammsb.h, exposing the subset of
// Platform/Device/Context/Queue/Buffer/Event the reference's callers touch.
#ifndef MCMC_B200_TYPES_H_
#define MCMC_B200_TYPES_H_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "ammsb.h"

namespace mcmc {

typedef uint64_t Edge;    // (min(u,v) << 32) | max(u,v)
typedef uint32_t Vertex;
typedef float Float;

// 128-bit RNG seed / state pair (reference: struct ulong2, types.h:35-54)
struct alignas(16) ulong2 {
  uint64_t values[2];
  ulong2() : values{0, 0} {}
  ulong2(std::initializer_list<uint64_t> l) { *this = l; }
  ulong2& operator=(std::initializer_list<uint64_t> l) {
    auto it = l.begin();
    values[0] = *it++;
    values[1] = *it;
    return *this;
  }
  uint64_t& operator[](size_t i) { return values[i]; }
  uint64_t operator[](size_t i) const { return values[i]; }
};
std::ostream& operator<<(std::ostream& out, const ulong2& v);
std::istream& operator>>(std::istream& in, ulong2& v);

inline std::tuple<Vertex, Vertex> Vertices(Edge e) {
  return std::tuple<Vertex, Vertex>(static_cast<Vertex>(e >> 32), static_cast<Vertex>(e & 0xffffffffu));
}
inline Edge MakeEdge(Vertex u, Vertex v) { return (static_cast<Edge>(u) << 32) | static_cast<Edge>(v); }

uint32_t GetMaxGroups();  // 65535: the reference's grid cap, which fixes its RNG-state mapping

// Every failure of the backend is fatal in the reference (LOG(FATAL)); here it throws.
struct BackendError : std::runtime_error {
  explicit BackendError(const std::string& what) : std::runtime_error(what) {}
};
inline void AmmsbCheck(int rc) {
  if (rc != 0) throw BackendError(ammsb_last_error());
}

namespace clcuda {

class Platform {
 public:
  explicit Platform(size_t id = 0) : id_(id) {}
  size_t id_;
};

class Device {
 public:
  Device() : ordinal_(0) {}
  Device(const Platform&, size_t ordinal) : ordinal_(static_cast<int>(ordinal)) {}
  explicit Device(int ordinal) : ordinal_(ordinal) {}
  std::string Type() const { return "GPU"; }
  std::string Vendor() const { return "NVIDIA"; }
  std::string Name() const;
  std::string Version() const { return ammsb_version(); }
  int Ordinal() const { return ordinal_; }

 private:
  int ordinal_;
};

// Context owns one ammsb_ctx (device + stream); Queue is a shared handle to a context,
// so copies of a Queue enqueue on the same stream (CLCudaAPI semantics).
// Context = a device plus its primary stream (allocations and synchronous copies go
// through it); Queue = one more stream on that device.  Copies of either share the
// underlying ammsb_ctx, as CLCudaAPI handles do.
class Context {
 public:
  Context() {}
  explicit Context(const Device& dev);
  ammsb_ctx* get() const { return impl_.get(); }
  const std::shared_ptr<ammsb_ctx>& shared() const { return impl_; }
  int DeviceOrdinal() const { return ordinal_; }

 private:
  std::shared_ptr<ammsb_ctx> impl_;
  int ordinal_ = 0;
};

class Queue {
 public:
  Queue() {}
  Queue(const Context& ctx, const Device& dev);  // a stream of its own on ctx's device
  Context GetContext() const { return ctx_; }
  Device GetDevice() const { return Device(ctx_.DeviceOrdinal()); }
  void Finish() const { AmmsbCheck(ammsb_ctx_sync(stream_.get())); }
  ammsb_ctx* operator()() const { return stream_.get(); }

 private:
  Context ctx_;
  std::shared_ptr<ammsb_ctx> stream_;
};

class Event {
 public:
  float GetElapsedTime() const { return ms_; }
  float ms_ = 0;
};

template <class T>
class Buffer {
 public:
  Buffer(const Context& ctx, size_t count) : ctx_(ctx), count_(count) { Alloc(); }
  template <class It>
  Buffer(const Context& ctx, const Queue&, It begin, It end) : ctx_(ctx) {
    std::vector<T> host(begin, end);
    count_ = host.size();
    Alloc();
    if (count_) AmmsbCheck(ammsb_h2d(ctx_.get(), ptr_.get(), host.data(), count_ * sizeof(T)));
  }
  size_t GetSize() const { return count_ * sizeof(T); }
  void Read(const Queue& q, size_t n, T* host, size_t offset = 0) const {
    AmmsbCheck(ammsb_d2h(q(), host, data() + offset, n * sizeof(T)));
  }
  void Read(const Queue& q, size_t n, std::vector<T>& host, size_t offset = 0) const {
    Read(q, n, host.data(), offset);
  }
  void ReadAsync(const Queue& q, size_t n, T* host, size_t offset = 0) const {
    AmmsbCheck(ammsb_d2h_async(q(), host, data() + offset, n * sizeof(T)));
  }
  void Write(const Queue& q, size_t n, const T* host, size_t offset = 0) {
    AmmsbCheck(ammsb_h2d(q(), data() + offset, host, n * sizeof(T)));
  }
  void Write(const Queue& q, size_t n, const std::vector<T>& host, size_t offset = 0) {
    Write(q, n, host.data(), offset);
  }
  void WriteAsync(const Queue& q, size_t n, const T* host, size_t offset = 0) {
    AmmsbCheck(ammsb_h2d_async(q(), data() + offset, host, n * sizeof(T)));
  }
  void CopyTo(const Queue& q, size_t n, Buffer<T>& dst) const {
    AmmsbCheck(ammsb_d2d(q(), dst.data(), data(), n * sizeof(T)));
  }
  T* data() const { return static_cast<T*>(ptr_.get()); }
  T* operator()() const { return data(); }

 private:
  void Alloc() {
    void* p = nullptr;
    AmmsbCheck(ammsb_malloc(ctx_.get(), count_ * sizeof(T), &p));
    std::shared_ptr<ammsb_ctx> keep = ctx_.shared();
    ptr_ = std::shared_ptr<void>(p, [keep](void* q) { ammsb_free(keep.get(), q); });
  }
  Context ctx_;
  size_t count_ = 0;
  std::shared_ptr<void> ptr_;
};

}  // namespace clcuda

}  // namespace mcmc

#endif  // MCMC_B200_TYPES_H_
