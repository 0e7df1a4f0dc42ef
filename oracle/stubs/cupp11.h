// Shape-only stand-in for CLCudaAPI's cupp11.h: lets the reference's host headers
// compile; nothing here ever runs (oracle/_ref only).
#ifndef ORACLE_STUB_CUPP11_H_
#define ORACLE_STUB_CUPP11_H_
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
namespace CLCudaAPI {
enum class BuildStatus { kSuccess, kError, kInvalid };
struct Platform { explicit Platform(size_t) {} };
struct Device {
  Device() {}
  Device(const Platform&, size_t) {}
  std::string Type() const { return "CPU"; }
  std::string Name() const { return "stub"; }
  std::string Vendor() const { return "stub"; }
  std::string Version() const { return "0"; }
  uint64_t MaxAllocSize() const { return 1ull << 40; }
};
struct Context { Context() {} explicit Context(const Device&) {} void* operator()() const { return nullptr; } };
struct Event { float GetElapsedTime() const { return 0; } };
struct Queue {
  Queue() {}
  Queue(const Context&, const Device&) {}
  Context GetContext() const { return Context(); }
  Device GetDevice() const { return Device(); }
  void Finish() const {}
  void* operator()() const { return nullptr; }
};
struct Program {
  Program(const Context&, const std::string&) {}
  BuildStatus Build(const Device&, std::vector<std::string>&) { return BuildStatus::kSuccess; }
  std::string GetBuildInfo(const Device&) const { return ""; }
};
template <class T>
struct Buffer {
  std::vector<T> host_;
  Buffer(const Context&, size_t n) : host_(n) {}
  template <class It> Buffer(const Context&, const Queue&, It b, It e) : host_(b, e) {}
  size_t GetSize() const { return host_.size() * sizeof(T); }
  void Read(const Queue&, size_t n, T* p, size_t off = 0) const { for (size_t i = 0; i < n; ++i) p[i] = host_[off + i]; }
  void Read(const Queue&, size_t n, std::vector<T>& v, size_t off = 0) const { Read(Queue(), n, v.data(), off); }
  void Write(const Queue&, size_t n, const T* p, size_t off = 0) { for (size_t i = 0; i < n; ++i) host_[off + i] = p[i]; }
  void Write(const Queue&, size_t n, const std::vector<T>& v, size_t off = 0) { Write(Queue(), n, v.data(), off); }
  void CopyTo(const Queue&, size_t n, Buffer<T>& d) const { for (size_t i = 0; i < n; ++i) d.host_[i] = host_[i]; }
  T* operator()() { return host_.data(); }
};
struct Kernel {
  Kernel(const Program&, const std::string&) {}
  template <class T> void SetArgument(size_t, const T&) {}
  void Launch(const Queue&, const std::vector<size_t>&, const std::vector<size_t>&, Event&) {}
};
}  // namespace CLCudaAPI
inline int cuCtxSetCurrent(void*) { return 0; }
#endif
