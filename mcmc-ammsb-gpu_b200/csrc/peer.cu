// peer.cu -- the two exchange steps of the multi-GPU iteration as kernels over NVLink peer
// memory: a cross-GPU barrier and a rank-ordered all-reduce of a small vector.
//
// One process per GPU.  Every rank owns a "mailbox" (virtual-memory allocation shared with the
// peers as a file descriptor, like the pi shards): flags[world] + 2 x slots[world][slot_bytes].
//   barrier     thread r writes epoch into flags[my rank] of rank r's mailbox (st.release.sys over
//               NVLink) and waits until flags[r] of its own mailbox reaches epoch.  Because the
//               kernel is enqueued behind the rank's earlier kernels, "rank r has arrived" means
//               its earlier reads of pi are done and its peer stores have landed.
//   all-reduce  every rank stores its vector into slot[my rank] of every mailbox, barrier, then
//               sums the `world` slots of its own mailbox in rank order.  The sum is therefore
//               bit-identical on every rank and independent of any library's algorithm choice
//               (the reference's serialize-test.cc:132 determinism contract; beta replicas must
//               not drift apart).  Slots are double-buffered by epoch parity, so one barrier per
//               all-reduce suffices.
// The reference has no multi-device code; these replace what would otherwise be three NCCL
// all-reduces per iteration (two 1-element ones used as barriers, one of [2K] floats).
// Each GPU runs exactly one such kernel at a time, waiting only on kernels of OTHER GPUs.
#include "common.cuh"

struct PeerView {
  unsigned char* box[AMMSB_MAX_SHARDS];  // every rank's mailbox as mapped here
  uint32_t world, rank;
  uint32_t slot_bytes;
};

struct ammsb_peer {
  ammsb_ctx* ctx = nullptr;
  uint32_t world = 1, rank = 0;
  size_t slot_bytes = 0, bytes = 0;
  VmmAlloc local;
  VmmAlloc remote[AMMSB_MAX_SHARDS];
  unsigned char* box[AMMSB_MAX_SHARDS] = {nullptr};
  uint32_t epoch = 0;
  PeerView view() const {
    PeerView v;
    for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) v.box[i] = box[i];
    v.world = world;
    v.rank = rank;
    v.slot_bytes = (uint32_t)slot_bytes;
    return v;
  }
};

#define PEER_FLAG_STRIDE 32u                       // one flag per 128-byte line (in uint32 units)
#define PEER_HEADER_BYTES (AMMSB_MAX_SHARDS * 128u)
#define PEER_ERR_WORD 16u                          // second half of flag line 0: epoch of a timed-out wait
#define PEER_SPIN_LIMIT (1u << 26)                 // x >= 64 ns: more than 4 s

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// executed by threads 0..world-1 of one CTA; followed by a CTA barrier in the callers
__device__ __forceinline__ void peer_signal_and_wait(const PeerView& v, uint32_t epoch) {
  const uint32_t t = threadIdx.x;
  if (t < v.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(v.box[t]) + v.rank * PEER_FLAG_STRIDE, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(v.box[v.rank]) + t * PEER_FLAG_STRIDE;
    // bounded: a rank that died (or never launched) must not hang this GPU inside a kernel.  On
    // expiry the error word of the own mailbox is set (ammsb_peer_check reads it) and the kernel
    // carries on -- its results are then meaningless, but the process stays controllable.
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      __nanosleep(64);
      if (++spins > PEER_SPIN_LIMIT) {
        atomicExch(reinterpret_cast<uint32_t*>(v.box[v.rank]) + PEER_ERR_WORD, epoch);
        break;
      }
    }
  }
}

__global__ void __launch_bounds__(32) k_peer_barrier(const __grid_constant__ PeerView v, uint32_t epoch) {
  peer_signal_and_wait(v, epoch);
  __syncthreads();
  __threadfence_system();
}

template <class T>
__global__ void __launch_bounds__(256)
    k_peer_allreduce(const __grid_constant__ PeerView v, uint32_t epoch, T* inout, uint32_t count) {
  const uint32_t set = epoch & 1;
  const size_t slots_off = PEER_HEADER_BYTES + (size_t)set * v.world * v.slot_bytes;
  // 1. my vector into slot[my rank] of every mailbox (peer stores; my own is a local store)
  for (uint32_t r = 0; r < v.world; ++r) {
    T* dst = reinterpret_cast<T*>(v.box[r] + slots_off + (size_t)v.rank * v.slot_bytes);
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) dst[i] = inout[i];
  }
  __syncthreads();  // all of this CTA's stores are issued before the release below
  // 2. everyone has delivered
  peer_signal_and_wait(v, epoch);
  __syncthreads();
  // 3. rank-ordered sum of the slots of my own mailbox
  const unsigned char* base = v.box[v.rank] + slots_off;
  for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) {
    T s = reinterpret_cast<const volatile T*>(base)[i];
    for (uint32_t r = 1; r < v.world; ++r) s += reinterpret_cast<const volatile T*>(base + (size_t)r * v.slot_bytes)[i];
    inout[i] = s;
  }
}

extern "C" int ammsb_peer_create(ammsb_ctx* c, uint32_t world, uint32_t rank, size_t slot_bytes, ammsb_peer** out) {
  AMMSB_REQUIRE(world >= 1 && world <= AMMSB_MAX_SHARDS && rank < world, "bad world / rank");
  AMMSB_REQUIRE(slot_bytes > 0 && slot_bytes % 16 == 0, "slot_bytes must be a positive multiple of 16");
  int ndev = 0;
  AMMSB_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  // one rank per GPU: two ranks' waiting kernels on one device can starve each other
  AMMSB_REQUIRE((int)world <= ndev, "more ranks than CUDA devices");
  ammsb_peer* p = new ammsb_peer();
  p->ctx = c;
  p->world = world;
  p->rank = rank;
  p->slot_bytes = slot_bytes;
  p->bytes = PEER_HEADER_BYTES + 2 * (size_t)world * slot_bytes;
  if (vmm_alloc(c->device, p->bytes, &p->local)) {
    delete p;
    return 1;
  }
  p->box[rank] = reinterpret_cast<unsigned char*>(p->local.ptr);
  AMMSB_CHECK_CUDA(cudaMemsetAsync(p->box[rank], 0, p->bytes, c->stream));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  *out = p;
  return 0;
}

extern "C" int ammsb_peer_destroy(ammsb_peer* p) {
  if (!p) return 0;
  cudaSetDevice(p->ctx->device);
  for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) vmm_free(&p->remote[i]);
  vmm_free(&p->local);
  delete p;
  return 0;
}

extern "C" int ammsb_peer_export_fd(ammsb_peer* p, int* fd) { return vmm_export_fd(p->local, fd); }

extern "C" int ammsb_peer_attach_fd(ammsb_peer* p, uint32_t peer_rank, int fd) {
  AMMSB_REQUIRE(peer_rank < p->world && peer_rank != p->rank, "bad peer rank");
  if (vmm_import_fd(p->ctx->device, fd, p->bytes, &p->remote[peer_rank])) return 1;
  p->box[peer_rank] = reinterpret_cast<unsigned char*>(p->remote[peer_rank].ptr);
  return 0;
}

extern "C" int ammsb_peer_check(ammsb_peer* p, uint32_t* timed_out_epoch) {
  AMMSB_CHECK_CUDA(cudaSetDevice(p->ctx->device));
  AMMSB_CHECK_CUDA(cudaMemcpy(timed_out_epoch, p->box[p->rank] + 4 * PEER_ERR_WORD, 4, cudaMemcpyDeviceToHost));
  return 0;
}

static int peer_ready(const ammsb_peer* p) {
  for (uint32_t r = 0; r < p->world; ++r) AMMSB_REQUIRE(p->box[r] != nullptr, "a peer mailbox is not attached");
  return 0;
}

extern "C" int ammsb_peer_barrier(ammsb_ctx* c, ammsb_peer* p) {
  if (peer_ready(p)) return 1;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  k_peer_barrier<<<1, 32, 0, c->stream>>>(p->view(), ++p->epoch);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_peer_allreduce_f32(ammsb_ctx* c, ammsb_peer* p, float* d_inout, uint32_t count) {
  if (peer_ready(p)) return 1;
  AMMSB_REQUIRE((size_t)count * sizeof(float) <= p->slot_bytes, "vector larger than the mailbox slot");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  k_peer_allreduce<float><<<1, 256, 0, c->stream>>>(p->view(), ++p->epoch, d_inout, count);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_peer_allreduce_f64(ammsb_ctx* c, ammsb_peer* p, double* d_inout, uint32_t count) {
  if (peer_ready(p)) return 1;
  AMMSB_REQUIRE((size_t)count * sizeof(double) <= p->slot_bytes, "vector larger than the mailbox slot");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  k_peer_allreduce<double><<<1, 256, 0, c->stream>>>(p->view(), ++p->epoch, d_inout, count);
  AMMSB_LAUNCH_CHECK();
  return 0;
}
