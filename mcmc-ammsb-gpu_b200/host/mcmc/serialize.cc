#include "mcmc/serialize.h"

#include <algorithm>

#include "mcmc/partitioned-alloc.h"

namespace mcmc {

namespace wire {

void PutVarint(std::string* s, uint64_t v) {
  while (v >= 0x80) {
    s->push_back(static_cast<char>((v & 0x7f) | 0x80));
    v >>= 7;
  }
  s->push_back(static_cast<char>(v));
}
void PutTag(std::string* s, uint32_t field, uint32_t type) { PutVarint(s, (field << 3) | type); }
void PutUInt(std::string* s, uint32_t field, uint64_t v) {
  PutTag(s, field, 0);
  PutVarint(s, v);
}
void PutDouble(std::string* s, uint32_t field, double v) {
  PutTag(s, field, 1);
  char b[8];
  std::memcpy(b, &v, 8);
  s->append(b, 8);
}
void PutBytes(std::string* s, uint32_t field, const void* p, size_t n) {
  PutTag(s, field, 2);
  PutVarint(s, n);
  s->append(static_cast<const char*>(p), n);
}

uint64_t Reader::Varint() {
  uint64_t v = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = static_cast<uint8_t>(*p++);
    v |= static_cast<uint64_t>(b & 0x7f) << shift;
    if (!(b & 0x80)) return v;
  }
  ok = false;
  return 0;
}
bool Reader::Next(uint32_t* field, uint32_t* type) {
  if (!ok || p >= end) return false;
  const uint64_t t = Varint();
  *field = static_cast<uint32_t>(t >> 3);
  *type = static_cast<uint32_t>(t & 7);
  return ok;
}
double Reader::Double() {
  double v = 0;
  if (end - p < 8) {
    ok = false;
    return 0;
  }
  std::memcpy(&v, p, 8);
  p += 8;
  return v;
}
bool Reader::Bytes(const char** data, size_t* n) {
  const uint64_t len = Varint();
  if (!ok || static_cast<uint64_t>(end - p) < len) {
    ok = false;
    return false;
  }
  *data = p;
  *n = len;
  p += len;
  return true;
}
void Reader::Skip(uint32_t type) {
  const char* d;
  size_t n;
  switch (type) {
    case 0: Varint(); break;
    case 1: Double(); break;
    case 2: Bytes(&d, &n); break;
    case 5: if (end - p >= 4) p += 4; else ok = false; break;
    default: ok = false;
  }
}

}  // namespace wire

bool WriteRecord(std::ostream* out, const std::string& payload) {
  const uint64_t n = payload.size();
  out->write(reinterpret_cast<const char*>(&n), sizeof n);
  out->write(payload.data(), n);
  return out->good();
}

bool ReadRecord(std::istream* in, std::string* payload) {
  uint64_t n = 0;
  in->read(reinterpret_cast<char*>(&n), sizeof n);
  if (!in->good()) return false;
  // the length prefix is file content: never allocate more than the stream can still deliver
  const std::streampos cur = in->tellg();
  if (cur != std::streampos(-1)) {
    in->seekg(0, std::ios::end);
    const std::streampos last = in->tellg();
    in->seekg(cur);
    if (last == std::streampos(-1) || static_cast<uint64_t>(last - cur) < n) return false;
    payload->resize(n);
    in->read(&(*payload)[0], n);
    return static_cast<uint64_t>(in->gcount()) == n;
  }
  payload->clear();  // not seekable: grow with what actually arrives
  const uint64_t kChunk = 64ull << 20;
  while (payload->size() < n) {
    const uint64_t want = std::min<uint64_t>(kChunk, n - payload->size());
    const size_t at = payload->size();
    payload->resize(at + want);
    in->read(&(*payload)[at], want);
    if (static_cast<uint64_t>(in->gcount()) != want) return false;
  }
  return true;
}

bool SerializeBytes(std::ostream* out, const void* data, size_t n) {
  std::string msg;
  msg.reserve(n + 16);
  wire::PutBytes(&msg, 1, data, n);
  return WriteRecord(out, msg);
}

bool ParseBytes(std::istream* in, void* data, size_t n) {
  std::string msg;
  if (!ReadRecord(in, &msg)) return false;
  wire::Reader r{msg.data(), msg.data() + msg.size()};
  uint32_t f, t;
  bool found = false;
  while (r.Next(&f, &t)) {
    if (f == 1 && t == 2) {
      const char* d;
      size_t len;
      if (!r.Bytes(&d, &len) || len != n) return false;
      std::memcpy(data, d, n);
      found = true;
    } else {
      r.Skip(t);
    }
  }
  return r.ok && found;
}

bool WriteRpmProperties(std::ostream* out, uint32_t rows, uint32_t cols, uint32_t rows_in_block) {
  std::string props;
  wire::PutUInt(&props, 1, rows);
  wire::PutUInt(&props, 2, cols);
  wire::PutUInt(&props, 3, rows_in_block);
  return WriteRecord(out, props);
}

bool ReadRpmProperties(std::istream* in, uint32_t* rows, uint32_t* cols, uint32_t* rows_in_block) {
  std::string props;
  if (!ReadRecord(in, &props)) return false;
  wire::Reader r{props.data(), props.data() + props.size()};
  uint32_t f, t, seen = 0;
  while (r.Next(&f, &t)) {
    if (t != 0) { r.Skip(t); continue; }
    const uint64_t v = r.Varint();
    if (f == 1) *rows = static_cast<uint32_t>(v);
    if (f == 2) *cols = static_cast<uint32_t>(v);
    if (f == 3) *rows_in_block = static_cast<uint32_t>(v);
    if (f >= 1 && f <= 3) seen |= 1u << f;
  }
  return r.ok && seen == 0xeu;  // all three are `required` (protos.proto)
}

bool SerializeRpm(std::ostream* out, RowPartitionedMatrix<Float>* rpm) {
  if (!WriteRpmProperties(out, static_cast<uint32_t>(rpm->Rows()), static_cast<uint32_t>(rpm->Cols()),
                          static_cast<uint32_t>(rpm->RowsPerBlock())))
    return false;
  std::vector<Float> host;
  for (uint64_t row = 0; row < rpm->Rows(); row += rpm->RowsPerBlock()) {
    const uint64_t n = std::min<uint64_t>(rpm->RowsPerBlock(), rpm->Rows() - row);
    host.resize(n * rpm->Cols());
    rpm->ReadRows(row, n, host.data());
    if (!SerializeBytes(out, host.data(), host.size() * sizeof(Float))) return false;
  }
  return true;
}

bool ParseRpm(std::istream* in, RowPartitionedMatrix<Float>* rpm) {
  uint32_t rows = 0, cols = 0, rib = 0;
  if (!ReadRpmProperties(in, &rows, &cols, &rib)) return false;
  if (rows != rpm->Rows() || cols != rpm->Cols() || rib != rpm->RowsPerBlock()) return false;
  std::vector<Float> host;
  for (uint64_t row = 0; row < rows; row += rib) {
    const uint64_t n = std::min<uint64_t>(rib, rows - row);
    host.resize(n * cols);
    if (!ParseBytes(in, host.data(), host.size() * sizeof(Float))) return false;
    rpm->WriteRows(row, n, host.data());
  }
  return true;
}

// ---- fixed-shape property messages: field 1 = count, doubles from field 2 on ----
namespace {
bool WriteCountAndDoubles(std::ostream* out, uint32_t count, const double* d, int n) {
  std::string msg;
  wire::PutUInt(&msg, 1, count);
  for (int i = 0; i < n; ++i) wire::PutDouble(&msg, 2 + i, d[i]);
  return WriteRecord(out, msg);
}
bool ReadCountAndDoubles(std::istream* in, uint32_t* count, double* d, int n) {
  std::string msg;
  if (!ReadRecord(in, &msg)) return false;
  wire::Reader r{msg.data(), msg.data() + msg.size()};
  uint32_t f, t, seen = 0;
  while (r.Next(&f, &t)) {
    if (f == 1 && t == 0) { *count = static_cast<uint32_t>(r.Varint()); seen |= 2u; }
    else if (t == 1 && f >= 2 && f < 2u + n) { d[f - 2] = r.Double(); seen |= 1u << f; }
    else r.Skip(t);
  }
  return r.ok && seen == ((1u << (2 + n)) - 2u);  // every field is `required` (protos.proto)
}
}  // namespace

bool SerializeMessage(std::ostream* out, const BetaProperties& m) {
  const double d[5] = {m.theta_sum_time, m.grads_partial_time, m.grads_sum_time, m.update_theta_time, m.normalize_time};
  return WriteCountAndDoubles(out, m.count_calls, d, 5);
}
bool ParseMessage(std::istream* in, BetaProperties* m) {
  double d[5] = {0, 0, 0, 0, 0};
  if (!ReadCountAndDoubles(in, &m->count_calls, d, 5)) return false;
  m->theta_sum_time = d[0]; m->grads_partial_time = d[1]; m->grads_sum_time = d[2];
  m->update_theta_time = d[3]; m->normalize_time = d[4];
  return true;
}
bool SerializeMessage(std::ostream* out, const PhiProperties& m) {
  const double d[2] = {m.update_phi_time, m.update_pi_time};
  return WriteCountAndDoubles(out, m.count_calls, d, 2);
}
bool ParseMessage(std::istream* in, PhiProperties* m) {
  double d[2] = {0, 0};
  if (!ReadCountAndDoubles(in, &m->count_calls, d, 2)) return false;
  m->update_phi_time = d[0]; m->update_pi_time = d[1];
  return true;
}
bool SerializeMessage(std::ostream* out, const PerplexityProperties& m) {
  const double d[2] = {m.ppx_time, m.accumulate_time};
  return WriteCountAndDoubles(out, m.count_calls, d, 2);
}
bool ParseMessage(std::istream* in, PerplexityProperties* m) {
  double d[2] = {0, 0};
  if (!ReadCountAndDoubles(in, &m->count_calls, d, 2)) return false;
  m->ppx_time = d[0]; m->accumulate_time = d[1];
  return true;
}

bool SerializeMessage(std::ostream* out, const SampleStorage& m) {
  std::string msg;
  wire::PutBytes(&msg, 1, m.edges.data(), m.edges.size());
  wire::PutBytes(&msg, 2, m.nodes_vec.data(), m.nodes_vec.size());
  wire::PutUInt(&msg, 3, m.seed);
  return WriteRecord(out, msg);
}
bool ParseMessage(std::istream* in, SampleStorage* m) {
  std::string msg;
  if (!ReadRecord(in, &msg)) return false;
  wire::Reader r{msg.data(), msg.data() + msg.size()};
  uint32_t f, t, seen = 0;
  while (r.Next(&f, &t)) {
    const char* d;
    size_t n;
    if (f == 1 && t == 2 && r.Bytes(&d, &n)) { m->edges.assign(d, n); seen |= 2u; }
    else if (f == 2 && t == 2 && r.Bytes(&d, &n)) { m->nodes_vec.assign(d, n); seen |= 4u; }
    else if (f == 3 && t == 0) { m->seed = static_cast<uint32_t>(r.Varint()); seen |= 8u; }
    else r.Skip(t);
  }
  return r.ok && seen == 0xeu;
}

bool SerializeMessage(std::ostream* out, const LearnerProperties& m) {
  std::string msg;
  wire::PutUInt(&msg, 1, m.stepCount);
  wire::PutUInt(&msg, 2, m.time);
  wire::PutUInt(&msg, 3, m.samplingTime);
  wire::PutUInt(&msg, 4, static_cast<uint64_t>(static_cast<int64_t>(m.phase)));  // int32: sign-extended varint
  wire::PutDouble(&msg, 5, m.weight);
  return WriteRecord(out, msg);
}
bool ParseMessage(std::istream* in, LearnerProperties* m) {
  std::string msg;
  if (!ReadRecord(in, &msg)) return false;
  wire::Reader r{msg.data(), msg.data() + msg.size()};
  uint32_t f, t, seen = 0;
  while (r.Next(&f, &t)) {
    if (t == 0) {
      const uint64_t v = r.Varint();
      if (f == 1) m->stepCount = static_cast<uint32_t>(v);
      if (f == 2) m->time = v;
      if (f == 3) m->samplingTime = v;
      if (f == 4) m->phase = static_cast<int32_t>(v);
      if (f >= 1 && f <= 4) seen |= 1u << f;
    } else if (t == 1 && f == 5) {
      m->weight = r.Double();
      seen |= 1u << 5;
    } else {
      r.Skip(t);
    }
  }
  return r.ok && seen == 0x3eu;
}

}  // namespace mcmc
