// probe_cols.cu -- hardware probes behind the column-sharded update_phi design.
//   A. random gather of small row pieces (256 B .. 4 KB) from a multi-GB array with TMA bulk
//      copies into a shared-memory ring: achievable HBM GB/s per piece size.
//   B. (2 GPUs) peer store -> peer poll latency (ping-pong) and the throughput of small
//      self-validating messages (4-byte words, a warp-wide store per message line).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o probe_cols probe_cols.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// A. every warp: ring of STAGES stages, a stage = G pieces of `piece` bytes at random rows.
// lanes 0..G-1 each issue one bulk copy (G <= 32).  After a stage lands every lane reads one
// float4 of it (so the data is really consumed) and the stage is refilled.
__global__ void k_gather(const float* __restrict__ base, uint32_t rows, uint32_t row_floats, uint32_t piece_floats,
                         uint32_t G, uint32_t stages, uint32_t trips, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t stage_floats = G * piece_floats;
  float* ring = reinterpret_cast<float*>(smem) + (size_t)wib * stages * stage_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)warps * stages * stage_floats * 4) + wib * stages;
  if (lane == 0)
    for (uint32_t s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint32_t gw = blockIdx.x * warps + wib;
  uint32_t phase = 0;
  float acc = 0.f;
  auto issue = [&](uint32_t t) {
    const uint32_t s = t % stages;
    if (lane == 0) mbar_expect_tx(&bars[s], stage_floats * 4);
    __syncwarp();
    if (lane < G) {
      const uint32_t r = mix(gw * 0x9e3779b9u + t * 61u + lane) % rows;
      bulk_g2s(ring + (size_t)s * stage_floats + lane * piece_floats, base + (size_t)r * row_floats, piece_floats * 4, &bars[s]);
    }
  };
  for (uint32_t t = 0; t < stages && t < trips; ++t) issue(t);
  for (uint32_t t = 0; t < trips; ++t) {
    const uint32_t s = t % stages;
    mbar_wait(&bars[s], (phase >> s) & 1);
    phase ^= 1u << s;
    const float4* st = reinterpret_cast<const float4*>(ring + (size_t)s * stage_floats);
    for (uint32_t i = lane; i < stage_floats / 4; i += 32) {
      const float4 v = st[i];
      acc += v.x + v.y + v.z + v.w;
    }
    __syncwarp();
    if (t + stages < trips) issue(t + stages);
  }
  if (acc == 12345.678f) sink[0] = acc;
}

// B1. ping-pong: rank 0 writes seq to rank 1's flag, rank 1 echoes into rank 0's flag.
__global__ void k_pingpong(volatile uint32_t* mine, volatile uint32_t* theirs, int rank, int iters, long long* cycles) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (rank == 0) {
      *theirs = i;
      while (*mine != (uint32_t)i) {}
    } else {
      while (*mine != (uint32_t)i) {}
      *theirs = i;
    }
  }
  cycles[0] = clock64() - t0;
}

// B2. message stream: the sender's warps store `words` 4-byte values per message (a warp-wide
// store of 32-bit words, `words` lanes active) into the receiver's mailbox; the receiver's warps
// poll every word until it differs from the sentinel, then reset it.  One message per warp per
// step, `depth` messages in flight per warp (the receiver lags).
__global__ void k_send(uint32_t* box, uint32_t words, uint32_t msgs_per_warp, uint32_t stride_words) {
  const uint32_t lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t* b = box + (size_t)gw * msgs_per_warp * stride_words;
  for (uint32_t m = 0; m < msgs_per_warp; ++m)
    if (lane < words) b[(size_t)m * stride_words + lane] = m + 1;
}
__global__ void k_recv(volatile uint32_t* box, uint32_t words, uint32_t msgs_per_warp, uint32_t stride_words,
                       unsigned long long* bad) {
  const uint32_t lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  volatile uint32_t* b = box + (size_t)gw * msgs_per_warp * stride_words;
  unsigned long long spins = 0;
  for (uint32_t m = 0; m < msgs_per_warp; ++m) {
    if (lane < words) {
      uint32_t v;
      while ((v = b[(size_t)m * stride_words + lane]) == 0xffffffffu) {
        if (++spins > (1ull << 26)) break;
      }
      if (v != m + 1) atomicAdd(bad, 1ull);
    }
  }
}

static float time_gather(const float* d, uint32_t rows, uint32_t row_floats, uint32_t piece_floats, uint32_t G,
                         uint32_t stages, uint32_t warps, uint32_t ctas_per_sm, uint32_t trips, float* sink, int sms) {
  const size_t smem = (size_t)warps * stages * G * piece_floats * 4 + (size_t)warps * stages * 8;
  CK(cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const uint32_t grid = sms * ctas_per_sm;
  k_gather<<<grid, warps * 32, smem>>>(d, rows, row_floats, piece_floats, G, stages, trips / 4, sink);
  CK(cudaEventRecord(a));
  k_gather<<<grid, warps * 32, smem>>>(d, rows, row_floats, piece_floats, G, stages, trips, sink);
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double bytes = (double)grid * warps * trips * G * piece_floats * 4;
  printf("gather piece=%5u B  G=%2u stages=%2u warps=%2u ctas/sm=%u smem=%6zu  %8.1f GB/s  (%.3f ms)\n", piece_floats * 4, G,
         stages, warps, ctas_per_sm, smem, bytes / ms / 1e6, ms);
  fflush(stdout);
  return ms;
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, %d devices\n", prop.name, sms, ndev);

  // ---- A ----
  {
    const uint32_t row_floats = 1024;          // 4 KB rows (K = 1024 layout); pieces start at row starts
    const uint32_t rows = 4u << 20;            // 16 GiB
    float* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)rows * row_floats * 4));
    CK(cudaMemset(d, 0, (size_t)rows * row_floats * 4));
    float* sink;
    CK(cudaMalloc((void**)&sink, 4));
    // 4 KB whole rows, one per stage (the shape of k_update_phi_fast at K = 1024)
    time_gather(d, rows, row_floats, 1024, 1, 4, 4, 2, 4000, sink, sms);
    time_gather(d, rows, row_floats, 1024, 1, 5, 8, 1, 4000, sink, sms);
    // G pieces per stage, G * piece = 4 KB (the column layout: G slots x one neighbor piece)
    for (uint32_t piece : {512u, 256u, 128u, 64u}) {   // floats: 2 KB, 1 KB, 512 B, 256 B
      const uint32_t G = 1024 / piece;
      time_gather(d, rows, row_floats, piece, G, 4, 4, 2, 4000, sink, sms);
      time_gather(d, rows, row_floats, piece, G, 5, 8, 1, 4000, sink, sms);
      time_gather(d, rows, row_floats, piece, G, 10, 4, 1, 4000, sink, sms);
      time_gather(d, rows, row_floats, piece, G, 6, 8, 1, 4000, sink, sms);
    }
    // 256-B pieces, 8 per stage (K = 512 on 8 GPUs)
    time_gather(d, rows, row_floats, 64, 8, 8, 8, 1, 8000, sink, sms);
    time_gather(d, rows, row_floats, 64, 8, 12, 8, 1, 8000, sink, sms);
    time_gather(d, rows, row_floats, 64, 8, 16, 4, 2, 8000, sink, sms);
    CK(cudaFree(d));
    CK(cudaFree(sink));
  }
  if (ndev < 2 || (argc > 1 && atoi(argv[1]) == 1)) return 0;

  // ---- B ----
  int can01 = 0, can10 = 0;
  CK(cudaDeviceCanAccessPeer(&can01, 0, 1));
  CK(cudaDeviceCanAccessPeer(&can10, 1, 0));
  printf("peer access 0->1 %d, 1->0 %d\n", can01, can10);
  if (!can01 || !can10) return 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaSetDevice(1));
  CK(cudaDeviceEnablePeerAccess(0, 0));
  uint32_t* flag[2];
  long long* cyc[2];
  cudaStream_t st[2];
  for (int r = 0; r < 2; ++r) {
    CK(cudaSetDevice(r));
    CK(cudaMalloc((void**)&flag[r], 128));
    CK(cudaMemset(flag[r], 0, 128));
    CK(cudaMalloc((void**)&cyc[r], 8));
    CK(cudaStreamCreate(&st[r]));
  }
  for (int r = 0; r < 2; ++r) {
    CK(cudaSetDevice(r));
    CK(cudaDeviceSynchronize());
  }
  const int iters = 2000;
  for (int r = 1; r >= 0; --r) {
    CK(cudaSetDevice(r));
    k_pingpong<<<1, 32, 0, st[r]>>>(flag[r], flag[1 - r], r, iters, cyc[r]);
  }
  long long c0 = 0;
  int clk = 0;
  CK(cudaSetDevice(0));
  CK(cudaStreamSynchronize(st[0]));
  CK(cudaMemcpy(&c0, cyc[0], 8, cudaMemcpyDeviceToHost));
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  CK(cudaSetDevice(1));
  CK(cudaStreamSynchronize(st[1]));
  printf("ping-pong: %.0f cycles per round trip = %.2f us (clock %d kHz) -> one-way store-to-visible ~%.2f us\n",
         (double)c0 / iters, (double)c0 / iters / clk * 1e3, clk, (double)c0 / iters / clk * 1e3 / 2);

  // message stream 0 -> 1
  for (uint32_t words : {8u, 16u, 32u}) {
    for (uint32_t warps_total : {148u * 4, 148u * 16}) {
      const uint32_t msgs = 4096, stride = 32;
      const size_t n = (size_t)warps_total * msgs * stride;
      uint32_t* box;
      unsigned long long* bad;
      CK(cudaSetDevice(1));
      CK(cudaMalloc((void**)&box, n * 4));
      CK(cudaMemset(box, 0xff, n * 4));
      CK(cudaMalloc((void**)&bad, 8));
      CK(cudaMemset(bad, 0, 8));
      CK(cudaDeviceSynchronize());
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      CK(cudaEventRecord(a, st[1]));
      k_recv<<<warps_total / 4, 128, 0, st[1]>>>(box, words, msgs, stride, bad);
      CK(cudaEventRecord(b, st[1]));
      CK(cudaSetDevice(0));
      k_send<<<warps_total / 4, 128, 0, st[0]>>>(box, words, msgs, stride);
      CK(cudaStreamSynchronize(st[0]));
      CK(cudaSetDevice(1));
      CK(cudaEventSynchronize(b));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, a, b));
      unsigned long long hbad = 0;
      CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
      printf("messages %3u B x %u warps x %u: %.3f ms, %.1f GB/s payload, %llu bad\n", words * 4, warps_total, msgs, ms,
             (double)warps_total * msgs * words * 4 / ms / 1e6, hbad);
      CK(cudaFree(box));
      CK(cudaFree(bad));
    }
  }
  return 0;
}
