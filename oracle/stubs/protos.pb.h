// Shape-only protobuf messages (oracle/_ref only; checkpoint code is never executed here).
#ifndef ORACLE_STUB_PROTOS_H_
#define ORACLE_STUB_PROTOS_H_
#include <cstdint>
#include <string>
namespace mcmc {
struct StubMessage {
  int ByteSize() const { return 0; }
  bool SerializeToArray(void*, int) const { return true; }
  bool ParseFromArray(const void*, int) { return true; }
};
struct VectorStorage : StubMessage { std::string s_; std::string* mutable_storage() { return &s_; } const std::string& storage() const { return s_; } };
struct RpmProperties : StubMessage {
  uint32_t r_ = 0, c_ = 0, b_ = 0;
  void set_rows(uint32_t v) { r_ = v; } void set_cols(uint32_t v) { c_ = v; } void set_rows_in_block(uint32_t v) { b_ = v; }
  uint32_t rows() const { return r_; } uint32_t cols() const { return c_; } uint32_t rows_in_block() const { return b_; }
};
struct SampleStorage : StubMessage {
  std::string e_, n_; uint32_t seed_ = 0;
  std::string* mutable_edges() { return &e_; } std::string* mutable_nodes_vec() { return &n_; }
  const std::string& edges() const { return e_; } const std::string& nodes_vec() const { return n_; }
  void set_seed(uint32_t s) { seed_ = s; } uint32_t seed() const { return seed_; }
};
}  // namespace mcmc
#endif
