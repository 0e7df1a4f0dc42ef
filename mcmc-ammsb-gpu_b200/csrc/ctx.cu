// ctx.cu -- context, buffers, RNG pools, cuckoo sets and the node-partitioned
// pi/phi store behind the C ABI (include/ammsb.h).  Replaces the reference's
// OpenCL/CLCudaAPI context + program plumbing (learner.cc:77-156, types.cc).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launch_count{0};

void ammsb_set_error(const std::string& msg) { t_last_error = msg; }

extern "C" const char* ammsb_last_error(void) { return t_last_error.c_str(); }
extern "C" const char* ammsb_version(void) { return "ammsb-b200 0.1 (sm_100a)"; }

extern "C" int ammsb_launch_count(uint64_t* count) {
  *count = g_launch_count.load();
  return 0;
}

// config.cc:57-64: `out << std::scientific << f` then the literal is re-parsed.
extern "C" float ammsb_round_param(float f) {
  char buf[64];
  snprintf(buf, sizeof buf, "%e", (double)f);
  return strtof(buf, nullptr);
}

// learner.cc:41-43.  Evaluated on the host in fp32 and passed to the kernels as an
// argument (the reference evaluates the same expression per work-item).
extern "C" float ammsb_eps_t(const ammsb_params* p, uint32_t step_count) {
  return p->a * powf(1 + step_count / p->b, -p->c);
}

extern "C" int ammsb_device_count(int* count) {
  AMMSB_CHECK_CUDA(cudaGetDeviceCount(count));
  return 0;
}

extern "C" int ammsb_ctx_create(int device, ammsb_ctx** out) {
  int n = 0;
  AMMSB_CHECK_CUDA(cudaGetDeviceCount(&n));
  AMMSB_REQUIRE(n > 0, "no CUDA device: this library has no CPU fallback");
  AMMSB_REQUIRE(device >= 0 && device < n, "device ordinal out of range");
  AMMSB_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AMMSB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  AMMSB_REQUIRE(prop.major >= 10, "built for sm_100a (B200); device is older");
  ammsb_ctx* c = new ammsb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  strncpy(c->name, prop.name, sizeof(c->name) - 1);
  AMMSB_CHECK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  AMMSB_CHECK_CUDA(cudaEventCreate(&c->ev_start));
  AMMSB_CHECK_CUDA(cudaEventCreate(&c->ev_stop));
  *out = c;
  return 0;
}

extern "C" int ammsb_ctx_destroy(ammsb_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_stop) cudaEventDestroy(c->ev_stop);
  delete c;
  return 0;
}

extern "C" int ammsb_ctx_sync(ammsb_ctx* c) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int ammsb_ctx_set_stream(ammsb_ctx* c, void* s) {
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return 0;
}

extern "C" int ammsb_ctx_device(const ammsb_ctx* c, int* device) {
  *device = c->device;
  return 0;
}
extern "C" int ammsb_ctx_device_name(const ammsb_ctx* c, char* buf, size_t len) {
  if (len == 0) return 0;
  strncpy(buf, c->name, len - 1);
  buf[len - 1] = 0;
  return 0;
}
extern "C" int ammsb_ctx_sm_count(const ammsb_ctx* c, int* sms) {
  *sms = c->sm_count;
  return 0;
}

// ---- buffers ----
extern "C" int ammsb_malloc(ammsb_ctx* c, size_t bytes, void** d_ptr) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
  return 0;
}
extern "C" int ammsb_free(ammsb_ctx* c, void* d_ptr) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaFree(d_ptr));
  return 0;
}
extern "C" int ammsb_memset(ammsb_ctx* c, void* d_ptr, int value, size_t bytes) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMemsetAsync(d_ptr, value, bytes, c->stream));
  return 0;
}
extern "C" int ammsb_h2d(ammsb_ctx* c, void* d, const void* h, size_t bytes) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int ammsb_d2h(ammsb_ctx* c, void* h, const void* d, size_t bytes) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int ammsb_h2d_async(ammsb_ctx* c, void* d, const void* h, size_t bytes) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
extern "C" int ammsb_d2h_async(ammsb_ctx* c, void* h, const void* d, size_t bytes) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
extern "C" int ammsb_d2d(ammsb_ctx* c, void* dst, const void* src, size_t bytes) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
extern "C" int ammsb_host_alloc(size_t bytes, void** h_ptr) {
  AMMSB_CHECK_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 16));
  return 0;
}
extern "C" int ammsb_host_free(void* h_ptr) {
  AMMSB_CHECK_CUDA(cudaFreeHost(h_ptr));
  return 0;
}

extern "C" int ammsb_timer_start(ammsb_ctx* c) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaEventRecord(c->ev_start, c->stream));
  return 0;
}
extern "C" int ammsb_timer_stop_ms(ammsb_ctx* c, float* ms) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaEventRecord(c->ev_stop, c->stream));
  AMMSB_CHECK_CUDA(cudaEventSynchronize(c->ev_stop));
  AMMSB_CHECK_CUDA(cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
  return 0;
}

struct ammsb_event {
  int device;
  cudaEvent_t ev;
};
extern "C" int ammsb_event_create(ammsb_ctx* c, ammsb_event** out) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_event* e = new ammsb_event();
  e->device = c->device;
  AMMSB_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming));
  *out = e;
  return 0;
}
extern "C" int ammsb_event_record(ammsb_ctx* c, ammsb_event* e) {
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  AMMSB_CHECK_CUDA(cudaEventRecord(e->ev, c->stream));
  return 0;
}
extern "C" int ammsb_event_sync(ammsb_event* e) {
  AMMSB_CHECK_CUDA(cudaEventSynchronize(e->ev));
  return 0;
}
extern "C" int ammsb_event_destroy(ammsb_event* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaEventDestroy(e->ev);
  delete e;
  return 0;
}

// ---- cuckoo set ----
static const uint64_t kSetPrimes[4][2] = {  // cuckoo.cc:30-35
    {15485807ull, 920429591ull},
    {379906717ull, 740320571ull},
    {256204747ull, 379927517ull},
    {13ull, 17ull}};

SetView ammsb_set::view() const {
  SetView v;
  v.base = d_table;
  v.num_bins = num_bins;
  v.p1 = kSetPrimes[prime_idx][0];
  v.p2 = kSetPrimes[prime_idx][1];
  // ceil(2^128 / num_bins) for the divide-free exact remainder (set_mod, common.cuh); wraps to 0
  // for num_bins == 1, for which the remainder formula then yields 0 as it should
  const unsigned __int128 m = ~static_cast<unsigned __int128>(0) / num_bins + 1;
  v.m_hi = static_cast<uint64_t>(m >> 64);
  v.m_lo = static_cast<uint64_t>(m);
  return v;
}

extern "C" int ammsb_set_create(ammsb_ctx* c, const uint64_t* h_table, uint64_t num_bins,
                                uint32_t prime_idx, ammsb_set** out) {
  AMMSB_REQUIRE(prime_idx < 4, "prime_idx out of range");
  AMMSB_REQUIRE(num_bins > 0, "empty set table");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_set* s = new ammsb_set();
  s->ctx = c;
  s->num_bins = num_bins;
  s->prime_idx = prime_idx;
  size_t bytes = sizeof(uint64_t) * 2 * 4 * num_bins;
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&s->d_table, bytes));
  AMMSB_CHECK_CUDA(cudaMemcpyAsync(s->d_table, h_table, bytes, cudaMemcpyHostToDevice, c->stream));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  *out = s;
  return 0;
}

extern "C" int ammsb_set_destroy(ammsb_set* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  cudaFree(s->d_table);
  delete s;
  return 0;
}

// ---- store ----
StoreView ammsb_store::view() const {
  StoreView v;
  for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) {
    v.pi[i] = peer_pi[i];
    v.phi[i] = peer_phi[i];
  }
  v.rows_per_shard = (uint32_t)rows_per_shard;
  v.num_shards = num_shards;
  v.K = K;
  v.N = (uint32_t)N;
  for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) {
    v.mirror_pi[i] = mirror_pi[i];
    v.mirror_phi[i] = mirror_phi[i];
  }
  v.num_mirrors = num_mirrors;
  return v;
}

static int store_create_impl(ammsb_ctx* c, uint64_t N, uint32_t K, uint32_t num_shards, uint32_t shard_id,
                             bool shareable, ammsb_store** out) {
  AMMSB_REQUIRE(N > 0 && K > 0, "empty store");
  AMMSB_REQUIRE(N < 0xffffffffull, "N must fit a 32-bit Vertex (types.h:32)");
  AMMSB_REQUIRE(num_shards >= 1 && num_shards <= AMMSB_MAX_SHARDS, "num_shards out of range");
  AMMSB_REQUIRE(shard_id < num_shards, "shard_id out of range");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_store* s = new ammsb_store();
  s->ctx = c;
  s->N = N;
  s->K = K;
  s->num_shards = num_shards;
  s->shard_id = shard_id;
  s->rows_per_shard = (N + num_shards - 1) / num_shards;
  s->first_row = s->rows_per_shard * shard_id;
  uint64_t end = s->first_row + s->rows_per_shard;
  if (end > N) end = N;
  s->local_rows = end > s->first_row ? end - s->first_row : 0;
  s->shareable = shareable;
  // every shard is allocated at the full rows_per_shard so that row addressing is uniform
  if (shareable) {
    if (vmm_alloc(c->device, sizeof(float) * s->rows_per_shard * K, &s->vmm_pi) ||
        vmm_alloc(c->device, sizeof(float) * s->rows_per_shard, &s->vmm_phi)) {
      vmm_free(&s->vmm_pi);
      delete s;
      return 1;
    }
    s->d_pi = reinterpret_cast<float*>(s->vmm_pi.ptr);
    s->d_phi = reinterpret_cast<float*>(s->vmm_phi.ptr);
  } else {
    cudaError_t e = cudaMalloc((void**)&s->d_pi, sizeof(float) * s->rows_per_shard * K);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_phi, sizeof(float) * s->rows_per_shard);
    if (e != cudaSuccess) {
      cudaFree(s->d_pi);
      delete s;
      AMMSB_CHECK_CUDA(e);
    }
  }
  s->peer_pi[shard_id] = s->d_pi;
  s->peer_phi[shard_id] = s->d_phi;
  *out = s;
  return 0;
}

extern "C" int ammsb_store_create(ammsb_ctx* c, uint64_t N, uint32_t K, uint32_t num_shards,
                                  uint32_t shard_id, ammsb_store** out) {
  return store_create_impl(c, N, K, num_shards, shard_id, false, out);
}

extern "C" int ammsb_store_create_shareable(ammsb_ctx* c, uint64_t N, uint32_t K, uint32_t num_shards,
                                            uint32_t shard_id, ammsb_store** out) {
  return store_create_impl(c, N, K, num_shards, shard_id, true, out);
}

extern "C" int ammsb_store_export_fd(ammsb_store* s, int* pi_fd, int* phi_fd) {
  AMMSB_REQUIRE(s->shareable, "store was not created with ammsb_store_create_shareable");
  if (vmm_export_fd(s->vmm_pi, pi_fd)) return 1;
  return vmm_export_fd(s->vmm_phi, phi_fd);
}

extern "C" int ammsb_store_attach_fd(ammsb_store* s, uint32_t shard, int pi_fd, int phi_fd) {
  AMMSB_REQUIRE(shard < s->num_shards && shard != s->shard_id, "bad peer shard");
  if (vmm_import_fd(s->ctx->device, pi_fd, sizeof(float) * s->rows_per_shard * s->K, &s->vmm_peer_pi[shard]) ||
      vmm_import_fd(s->ctx->device, phi_fd, sizeof(float) * s->rows_per_shard, &s->vmm_peer_phi[shard]))
    return 1;
  s->peer_pi[shard] = reinterpret_cast<float*>(s->vmm_peer_pi[shard].ptr);
  s->peer_phi[shard] = reinterpret_cast<float*>(s->vmm_peer_phi[shard].ptr);
  s->peer_is_ipc[shard] = false;
  return 0;
}

extern "C" int ammsb_store_add_mirror_fd(ammsb_store* s, int pi_fd, int phi_fd) {
  AMMSB_REQUIRE(s->num_shards == 1, "mirrors belong to a replicated (num_shards = 1) store");
  AMMSB_REQUIRE(s->num_mirrors < AMMSB_MAX_SHARDS - 1, "too many mirrors");
  const uint32_t i = s->num_mirrors;
  if (vmm_import_fd(s->ctx->device, pi_fd, sizeof(float) * s->rows_per_shard * s->K, &s->vmm_mirror_pi[i]) ||
      vmm_import_fd(s->ctx->device, phi_fd, sizeof(float) * s->rows_per_shard, &s->vmm_mirror_phi[i]))
    return 1;
  s->mirror_pi[i] = reinterpret_cast<float*>(s->vmm_mirror_pi[i].ptr);
  s->mirror_phi[i] = reinterpret_cast<float*>(s->vmm_mirror_phi[i].ptr);
  s->mirror_is_ipc[i] = false;
  ++s->num_mirrors;
  return 0;
}

extern "C" int ammsb_store_destroy(ammsb_store* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  for (uint32_t i = 0; i < s->num_shards; ++i) {
    if (i != s->shard_id && s->peer_is_ipc[i]) {
      if (s->peer_pi[i]) cudaIpcCloseMemHandle(s->peer_pi[i]);
      if (s->peer_phi[i]) cudaIpcCloseMemHandle(s->peer_phi[i]);
    }
  }
  for (uint32_t i = 0; i < s->num_mirrors; ++i) {
    if (s->mirror_is_ipc[i]) {
      cudaIpcCloseMemHandle(s->mirror_pi[i]);
      cudaIpcCloseMemHandle(s->mirror_phi[i]);
    }
  }
  for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) {
    vmm_free(&s->vmm_peer_pi[i]);
    vmm_free(&s->vmm_peer_phi[i]);
    vmm_free(&s->vmm_mirror_pi[i]);
    vmm_free(&s->vmm_mirror_phi[i]);
  }
  if (s->shareable) {
    vmm_free(&s->vmm_pi);
    vmm_free(&s->vmm_phi);
  } else {
    cudaFree(s->d_pi);
    if (s->owns_phi) cudaFree(s->d_phi);
  }
  delete s;
  return 0;
}

extern "C" int ammsb_store_export(ammsb_store* s, uint8_t* pi_handle, uint8_t* phi_handle) {
  static_assert(sizeof(cudaIpcMemHandle_t) == AMMSB_IPC_HANDLE_BYTES, "ipc handle size");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  cudaIpcMemHandle_t h;
  AMMSB_CHECK_CUDA(cudaIpcGetMemHandle(&h, s->d_pi));
  memcpy(pi_handle, &h, sizeof h);
  AMMSB_CHECK_CUDA(cudaIpcGetMemHandle(&h, s->d_phi));
  memcpy(phi_handle, &h, sizeof h);
  return 0;
}

extern "C" int ammsb_store_attach(ammsb_store* s, uint32_t shard, const uint8_t* pi_handle,
                                  const uint8_t* phi_handle) {
  AMMSB_REQUIRE(shard < s->num_shards && shard != s->shard_id, "bad peer shard");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  cudaIpcMemHandle_t h;
  void* p = nullptr;
  memcpy(&h, pi_handle, sizeof h);
  AMMSB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  s->peer_pi[shard] = (float*)p;
  memcpy(&h, phi_handle, sizeof h);
  AMMSB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  s->peer_phi[shard] = (float*)p;
  s->peer_is_ipc[shard] = true;
  return 0;
}

extern "C" int ammsb_store_attach_local(ammsb_store* s, uint32_t shard, ammsb_store* peer) {
  AMMSB_REQUIRE(shard < s->num_shards && shard != s->shard_id, "bad peer shard");
  AMMSB_REQUIRE(peer->shard_id == shard && peer->N == s->N && peer->K == s->K &&
                    peer->num_shards == s->num_shards,
                "peer store does not match");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  if (peer->ctx->device != s->ctx->device) {
    int can = 0;
    AMMSB_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, s->ctx->device, peer->ctx->device));
    AMMSB_REQUIRE(can, "devices are not peer-accessible");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->ctx->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
    } else {
      AMMSB_CHECK_CUDA(e);
    }
  }
  s->peer_pi[shard] = peer->d_pi;
  s->peer_phi[shard] = peer->d_phi;
  s->peer_is_ipc[shard] = false;
  return 0;
}

extern "C" int ammsb_store_add_mirror(ammsb_store* s, const uint8_t* pi_handle, const uint8_t* phi_handle) {
  AMMSB_REQUIRE(s->num_shards == 1, "mirrors belong to a replicated (num_shards = 1) store");
  AMMSB_REQUIRE(s->num_mirrors < AMMSB_MAX_SHARDS - 1, "too many mirrors");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  cudaIpcMemHandle_t h;
  void* p = nullptr;
  memcpy(&h, pi_handle, sizeof h);
  AMMSB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  s->mirror_pi[s->num_mirrors] = (float*)p;
  memcpy(&h, phi_handle, sizeof h);
  AMMSB_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  s->mirror_phi[s->num_mirrors] = (float*)p;
  s->mirror_is_ipc[s->num_mirrors] = true;
  ++s->num_mirrors;
  return 0;
}

extern "C" int ammsb_store_add_mirror_local(ammsb_store* s, ammsb_store* peer) {
  AMMSB_REQUIRE(s->num_shards == 1 && peer->num_shards == 1 && peer->N == s->N && peer->K == s->K && peer != s,
                "mirror must be another full copy of the same shape");
  AMMSB_REQUIRE(s->num_mirrors < AMMSB_MAX_SHARDS - 1, "too many mirrors");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  if (peer->ctx->device != s->ctx->device) {
    int can = 0;
    AMMSB_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, s->ctx->device, peer->ctx->device));
    AMMSB_REQUIRE(can, "devices are not peer-accessible");
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->ctx->device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
    } else {
      AMMSB_CHECK_CUDA(e);
    }
  }
  s->mirror_pi[s->num_mirrors] = peer->d_pi;
  s->mirror_phi[s->num_mirrors] = peer->d_phi;
  s->mirror_is_ipc[s->num_mirrors] = false;
  ++s->num_mirrors;
  return 0;
}

extern "C" int ammsb_store_rows(const ammsb_store* s, uint64_t* first_row, uint64_t* num_rows) {
  *first_row = s->first_row;
  *num_rows = s->local_rows;
  return 0;
}

extern "C" int ammsb_store_local_ptrs(ammsb_store* s, float** d_pi, float** d_phi) {
  if (d_pi) *d_pi = s->d_pi;
  if (d_phi) *d_phi = s->d_phi;
  return 0;
}

extern "C" int ammsb_store_bind_phi(ammsb_store* s, float* d_phi) {
  AMMSB_REQUIRE(d_phi != nullptr, "null phi");
  AMMSB_REQUIRE(s->num_shards == 1 && s->num_mirrors == 0 && !s->shareable,
                "bind_phi is for a single, unmirrored, non-shareable store");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  if (s->owns_phi) {
    AMMSB_CHECK_CUDA(cudaMemcpyAsync(d_phi, s->d_phi, sizeof(float) * s->local_rows, cudaMemcpyDeviceToDevice,
                                     s->ctx->stream));
    AMMSB_CHECK_CUDA(cudaStreamSynchronize(s->ctx->stream));
    AMMSB_CHECK_CUDA(cudaFree(s->d_phi));
  }
  s->owns_phi = false;
  s->d_phi = d_phi;
  s->peer_phi[s->shard_id] = d_phi;
  return 0;
}

static int store_check_rows(ammsb_store* s, uint64_t row0, uint64_t nrows) {
  AMMSB_REQUIRE(row0 >= s->first_row && row0 + nrows <= s->first_row + s->local_rows,
                "rows are not owned by this shard");
  return 0;
}

extern "C" int ammsb_store_write_pi(ammsb_store* s, uint64_t row0, uint64_t nrows, const float* h) {
  if (store_check_rows(s, row0, nrows)) return 1;
  return ammsb_h2d(s->ctx, s->d_pi + (row0 - s->first_row) * s->K, h, sizeof(float) * nrows * s->K);
}
extern "C" int ammsb_store_read_pi(ammsb_store* s, uint64_t row0, uint64_t nrows, float* h) {
  if (store_check_rows(s, row0, nrows)) return 1;
  return ammsb_d2h(s->ctx, h, s->d_pi + (row0 - s->first_row) * s->K, sizeof(float) * nrows * s->K);
}
extern "C" int ammsb_store_write_phi(ammsb_store* s, uint64_t row0, uint64_t nrows, const float* h) {
  if (store_check_rows(s, row0, nrows)) return 1;
  return ammsb_h2d(s->ctx, s->d_phi + (row0 - s->first_row), h, sizeof(float) * nrows);
}
extern "C" int ammsb_store_read_phi(ammsb_store* s, uint64_t row0, uint64_t nrows, float* h) {
  if (store_check_rows(s, row0, nrows)) return 1;
  return ammsb_d2h(s->ctx, h, s->d_phi + (row0 - s->first_row), sizeof(float) * nrows);
}
