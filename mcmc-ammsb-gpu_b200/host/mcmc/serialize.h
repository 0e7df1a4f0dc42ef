// mcmc/serialize.h -- checkpoint stream: records of `uint64 byte_size` + a proto2
// message (reference serialize.h:13-38, protos.proto).  The eight tiny messages are
// encoded/decoded by hand (no protobuf dependency); the wire bytes are those protoc's
// generated code would write for the same field values.
#ifndef MCMC_B200_SERIALIZE_H_
#define MCMC_B200_SERIALIZE_H_

#include <cstring>
#include <istream>
#include <ostream>
#include <string>
#include <vector>

#include "mcmc/types.h"

namespace mcmc {

// ---- proto2 wire primitives ----
namespace wire {
void PutVarint(std::string* s, uint64_t v);
void PutTag(std::string* s, uint32_t field, uint32_t type);
void PutUInt(std::string* s, uint32_t field, uint64_t v);        // varint (uint32/uint64/int32)
void PutDouble(std::string* s, uint32_t field, double v);        // fixed64
void PutBytes(std::string* s, uint32_t field, const void* p, size_t n);  // length-delimited
struct Reader {
  const char* p;
  const char* end;
  bool ok = true;
  bool Next(uint32_t* field, uint32_t* type);
  uint64_t Varint();
  double Double();
  bool Bytes(const char** data, size_t* n);
  void Skip(uint32_t type);
};
}  // namespace wire

bool WriteRecord(std::ostream* out, const std::string& payload);
bool ReadRecord(std::istream* in, std::string* payload);

// message VectorStorage { required bytes storage = 1; }
bool SerializeBytes(std::ostream* out, const void* data, size_t n);
bool ParseBytes(std::istream* in, void* data, size_t n);  // fails on size mismatch (serialize.h:62-69)

template <class T>
bool Serialize(std::ostream* out, clcuda::Buffer<T>* buf, clcuda::Queue* queue) {
  std::vector<T> host(buf->GetSize() / sizeof(T));
  buf->Read(*queue, host.size(), host.data());
  return SerializeBytes(out, host.data(), buf->GetSize());
}
template <class T>
bool Parse(std::istream* in, clcuda::Buffer<T>* buf, clcuda::Queue* queue) {
  std::vector<T> host(buf->GetSize() / sizeof(T));
  if (!ParseBytes(in, host.data(), buf->GetSize())) return false;
  buf->Write(*queue, host.size(), host.data());
  return true;
}

// message RpmProperties { uint32 rows = 1; uint32 cols = 2; uint32 rows_in_block = 3; }
// followed by one VectorStorage per block of rows_in_block rows (serialize.h:72-113)
template <class T>
class RowPartitionedMatrix;
bool WriteRpmProperties(std::ostream* out, uint32_t rows, uint32_t cols, uint32_t rows_in_block);
bool ReadRpmProperties(std::istream* in, uint32_t* rows, uint32_t* cols, uint32_t* rows_in_block);
bool SerializeRpm(std::ostream* out, RowPartitionedMatrix<Float>* rpm);
bool ParseRpm(std::istream* in, RowPartitionedMatrix<Float>* rpm);

struct BetaProperties {
  uint32_t count_calls = 0;
  double theta_sum_time = 0, grads_partial_time = 0, grads_sum_time = 0, update_theta_time = 0, normalize_time = 0;
};
struct PhiProperties {
  uint32_t count_calls = 0;
  double update_phi_time = 0, update_pi_time = 0;
};
struct PerplexityProperties {
  uint32_t count_calls = 0;
  double ppx_time = 0, accumulate_time = 0;
};
struct SampleStorage {
  std::string edges, nodes_vec;
  uint32_t seed = 0;
};
struct LearnerProperties {
  uint32_t stepCount = 0;
  uint64_t time = 0, samplingTime = 0;
  int32_t phase = 0;
  double weight = 0;
};

bool SerializeMessage(std::ostream* out, const BetaProperties& m);
bool SerializeMessage(std::ostream* out, const PhiProperties& m);
bool SerializeMessage(std::ostream* out, const PerplexityProperties& m);
bool SerializeMessage(std::ostream* out, const SampleStorage& m);
bool SerializeMessage(std::ostream* out, const LearnerProperties& m);
bool ParseMessage(std::istream* in, BetaProperties* m);
bool ParseMessage(std::istream* in, PhiProperties* m);
bool ParseMessage(std::istream* in, PerplexityProperties* m);
bool ParseMessage(std::istream* in, SampleStorage* m);
bool ParseMessage(std::istream* in, LearnerProperties* m);

}  // namespace mcmc

#endif  // MCMC_B200_SERIALIZE_H_
