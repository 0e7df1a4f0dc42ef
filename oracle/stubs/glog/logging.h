// Stub of glog for compiling the reference's host sources unmodified (oracle/_ref only).
#ifndef ORACLE_STUB_GLOG_H_
#define ORACLE_STUB_GLOG_H_
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <vector>
#include <string>
#include <random>
#include <sstream>
namespace stubglog {
enum Severity { INFO, WARNING, ERROR, FATAL, DFATAL };
struct Msg {
  Severity s; bool on; std::ostringstream o;
  Msg(Severity sv, bool enabled) : s(sv), on(enabled) {}
  ~Msg() {
    if (!on) return;
    if (s >= ERROR) std::cerr << "[ref] " << o.str() << std::endl;
    if (s == FATAL) abort();
  }
  template <class T> Msg& operator<<(const T& v) { if (on) o << v; return *this; }
  Msg& operator<<(std::ostream& (*f)(std::ostream&)) { if (on) o << f; return *this; }
};
}  // namespace stubglog
#define LOG(sev) ::stubglog::Msg(::stubglog::sev, true)
#define LOG_IF(sev, cond) ::stubglog::Msg(::stubglog::sev, (cond))
#define CHECK(cond) ::stubglog::Msg(::stubglog::FATAL, !(cond))
#endif
