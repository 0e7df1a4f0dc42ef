// perplexity.cu -- held-out perplexity (reference: mcmc/perplexity.cc, perplexity.cu).
//
// The reference writes four [H] arrays (link/non-link log-likelihood and counts),
// reduces each with a library call, and in its GPU variant also round-trips a
// [H,K] scratch matrix.  Here one warp handles a held-out pair: both pi rows are
// read exactly once with 128-bit streaming loads, the running mean per pair is
// updated in place, and the four sums are reduced in a fixed order
// (warp -> CTA -> grid) in double precision.
#include "common.cuh"

// 128-thread CTAs: 91 registers/thread lets 5 of them share an SM (20 warps x 8 KB of row loads in
// flight); with 256 threads only 2 CTAs (16 warps) fit.
#define PPX_WARPS 4

struct PpxArgs {
  StoreView sv;
  SetView set;
  const float* beta;
  const uint64_t* edges;
  uint32_t H, K;
  float epsilon;
  float* ppx_per_edge;
  uint32_t call_count;
  double* partial;  // [gridDim.x][4]
};

__global__ void __launch_bounds__(PPX_WARPS * 32, 5) k_perplexity(const __grid_constant__ PpxArgs a) {
  extern __shared__ __align__(16) float s_beta[];  // [K]
  __shared__ double s_part[PPX_WARPS][4];
  const uint32_t K = a.K;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) s_beta[k] = __ldg(&a.beta[2 * k + 1]);
  __syncthreads();
  const bool vec = (K & 3) == 0;
  double link_lik = 0.0, non_lik = 0.0;
  uint32_t link_cnt = 0, non_cnt = 0;
  const uint32_t gwarp = blockIdx.x * PPX_WARPS + wib;
  const uint32_t nwarps = gridDim.x * PPX_WARPS;
  for (uint32_t i = gwarp; i < a.H; i += nwarps) {
    const uint64_t e = __ldg(&a.edges[i]);
    const uint32_t u = (uint32_t)(e >> 32), v = (uint32_t)(e & 0xffffffffu);
    const float* pa = store_row(a.sv, u);
    const float* pb = store_row(a.sv, v);
    // membership of the key as stored (perplexity.cc:45-47)
    const bool is_edge = set_has(a.set, e);
    float sb = 0.f, sq = 0.f;  // sum f*beta, sum f
    if (vec) {
      for (uint32_t k = lane * 4; k < K; k += 128) {
        const float4 x = ldg_stream4(pa + k);
        const float4 z = ldg_stream4(pb + k);
        const float4 b = *reinterpret_cast<const float4*>(s_beta + k);
        const float f0 = x.x * z.x, f1 = x.y * z.y, f2 = x.z * z.z, f3 = x.w * z.w;
        sq += (f0 + f1) + (f2 + f3);
        sb = fmaf(f0, b.x, sb);
        sb = fmaf(f1, b.y, sb);
        sb = fmaf(f2, b.z, sb);
        sb = fmaf(f3, b.w, sb);
      }
    } else {
      for (uint32_t k = lane; k < K; k += 32) {
        const float f = pa[k] * pb[k];
        sq += f;
        sb = fmaf(f, s_beta[k], sb);
      }
    }
    sb = warp_sum(sb);
    sq = warp_sum(sq);
    if (lane == 0) {
      // calculate_edge_likelihood, perplexity.cc:16-40:
      //   link:     s = sum f*beta
      //   non-link: s = sum f*(1-beta) + (1 - sum f)*(1 - epsilon)
      float s = is_edge ? sb : (sq - sb) + (1.0f - sq) * (1.0f - a.epsilon);
      if (s < 1.0e-30f) s = 1.0e-30f;
      // running mean over calls, perplexity.cc:51-52
      float ppx = a.ppx_per_edge[i];
      ppx = __fdiv_rn(__fadd_rn(__fmul_rn(ppx, (float)(a.call_count - 1)), s), (float)a.call_count);
      a.ppx_per_edge[i] = ppx;
      const float l = logf(ppx);
      if (is_edge) {
        link_lik += (double)l;
        ++link_cnt;
      } else {
        non_lik += (double)l;
        ++non_cnt;
      }
    }
  }
  if (lane == 0) {
    s_part[wib][0] = link_lik;
    s_part[wib][1] = non_lik;
    s_part[wib][2] = (double)link_cnt;
    s_part[wib][3] = (double)non_cnt;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < PPX_WARPS; ++w) s += s_part[w][threadIdx.x];
    a.partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

// one warp per quantity: lane-strided partial sums in a fixed order, then the shuffle tree
__global__ void __launch_bounds__(128) k_ppx_reduce(const double* __restrict__ partial, uint32_t P,
                                                    double* __restrict__ sums) {
  const uint32_t q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (uint32_t p = lane; p < P; p += 32) s += partial[(size_t)p * 4 + q];
  s = warp_sum_d(s);
  if (lane == 0) sums[q] = s;
}

static uint32_t ppx_max_ctas(const ammsb_ctx* c) { return (uint32_t)c->sm_count * 10; }

extern "C" int ammsb_perplexity_workspace_bytes(ammsb_ctx* c, size_t* bytes) {
  *bytes = sizeof(double) * 4 * (size_t)ppx_max_ctas(c) + sizeof(double) * 4;
  return 0;
}

extern "C" int ammsb_perplexity_partial(ammsb_ctx* c, const ammsb_params* p, ammsb_store* store,
                                        const float* d_beta, ammsb_set* heldout,
                                        const uint64_t* d_edges, uint32_t H, float* d_ppx_per_edge,
                                        uint32_t call_count, double* d_sums, void* d_ws,
                                        size_t ws_bytes) {
  AMMSB_REQUIRE(p->K == store->K, "params do not match the store");
  AMMSB_REQUIRE(call_count >= 1, "call_count is 1-based (perplexity.cc:252)");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint32_t ctas = (H + PPX_WARPS - 1) / PPX_WARPS;
  if (ctas > ppx_max_ctas(c)) ctas = ppx_max_ctas(c);
  AMMSB_REQUIRE(ws_bytes >= sizeof(double) * 4 * (size_t)(ctas ? ctas : 1), "perplexity workspace too small");
  if (ctas > 0) {
    PpxArgs a;
    a.sv = store->view();
    a.set = heldout->view();
    a.beta = d_beta;
    a.edges = d_edges;
    a.H = H;
    a.K = p->K;
    a.epsilon = p->epsilon;
    a.ppx_per_edge = d_ppx_per_edge;
    a.call_count = call_count;
    a.partial = (double*)d_ws;
    const size_t smem = sizeof(float) * p->K;
    if (smem > 48 * 1024)
      AMMSB_CHECK_CUDA(cudaFuncSetAttribute(k_perplexity, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    k_perplexity<<<ctas, PPX_WARPS * 32, smem, c->stream>>>(a);
    AMMSB_LAUNCH_CHECK();
  }
  k_ppx_reduce<<<1, 128, 0, c->stream>>>((const double*)d_ws, ctas, d_sums);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_perplexity(ammsb_ctx* c, const ammsb_params* p, ammsb_store* store,
                                const float* d_beta, ammsb_set* heldout, const uint64_t* d_edges,
                                uint32_t H, float* d_ppx_per_edge, uint32_t call_count,
                                double* h_sums, double* h_avg, void* d_ws, size_t ws_bytes) {
  AMMSB_REQUIRE(ws_bytes >= sizeof(double) * 4 * ((size_t)ppx_max_ctas(c) + 1), "perplexity workspace too small");
  double* d_sums = (double*)d_ws + (size_t)ppx_max_ctas(c) * 4;
  int rc = ammsb_perplexity_partial(c, p, store, d_beta, heldout, d_edges, H, d_ppx_per_edge,
                                    call_count, d_sums, d_ws, ws_bytes);
  if (rc) return rc;
  double sums[4];
  rc = ammsb_d2h(c, sums, d_sums, sizeof sums);
  if (rc) return rc;
  if (h_sums) for (int i = 0; i < 4; ++i) h_sums[i] = sums[i];
  // perplexity.cc:264-273
  double avg = 0.0;
  if (sums[2] + sums[3] != 0) avg = (sums[0] + sums[1]) / (sums[2] + sums[3]);
  if (h_avg) *h_avg = -avg;
  return 0;
}
