// beta.cu -- update_beta/theta (reference: mcmc/beta.cc).
//
// The reference runs sum_theta, calculate_grads_partial (one [2K] partial per
// mini-batch edge, 128 MB at m=16384,K=1024), a serial sum_grads, update_theta, a
// theta->beta copy and a normalise kernel, each followed by a queue.Finish().
// Here:  k_beta_partial  -- warp per edge, both pi rows read once with 128-bit loads,
//                           per-warp register accumulators of  A_k = sum_{y=0} p_k/S,
//                           B_k = sum_{y=1} p_k/S, combined per CTA in warp order
//        k_beta_reduce   -- fixed-order sum over CTAs, theta_sum and the [K,2] gradient
//        k_update_theta  -- Langevin step + beta = normalised theta
// The summation order is fixed by the launch geometry, never by atomics, so a run is
// reproducible (the reference's serialize-test.cc:132 contract).
#include "common.cuh"

// 4 warps per CTA: k_beta_partial<8> needs 156 registers, so 128-thread CTAs fit 3 per SM
// (12 warps x 8 KB of row loads in flight); 256-thread CTAs fit only one (measured 53 us -> see
// profiles/).
#define BETA_WARPS 4

struct BetaArgs {
  StoreView sv;
  SetView set;
  const float* beta;
  const uint64_t* edges;
  uint32_t E_mb, K;
  float epsilon;
  float* partial;  // [gridDim.x][2][K]
};

// KPL4 = number of float4 per lane per row = ceil(K / 128)
template <int KPL4>
__global__ void __launch_bounds__(BETA_WARPS * 32, 3)
    k_beta_partial(const __grid_constant__ BetaArgs a) {
  extern __shared__ __align__(16) float s_mem[];
  const uint32_t K = a.K;
  float* s_beta = s_mem;     // [K]  Beta(k) = beta[2k+1]
  float* s_acc = s_mem + K;  // [2][K]
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) s_beta[k] = __ldg(&a.beta[2 * k + 1]);
  __syncthreads();

  float4 accA[KPL4], accB[KPL4];
#pragma unroll
  for (int i = 0; i < KPL4; ++i) {
    accA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    accB[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const uint32_t gwarp = blockIdx.x * BETA_WARPS + wib;
  const uint32_t nwarps = gridDim.x * BETA_WARPS;
  for (uint32_t e = gwarp; e < a.E_mb; e += nwarps) {
    const uint64_t edge = __ldg(&a.edges[e]);
    const uint32_t u = (uint32_t)(edge >> 32), v = (uint32_t)(edge & 0xffffffffu);
    const float* pa = store_row(a.sv, u);
    const float* pb = store_row(a.sv, v);
    float4 q[KPL4];
#pragma unroll
    for (int i = 0; i < KPL4; ++i) {
      const uint32_t k = lane * 4 + 128 * i;
      if (k < K) {
        const float4 x = ldg_stream4(pa + k);
        const float4 z = ldg_stream4(pb + k);
        q[i] = make_float4(x.x * z.x, x.y * z.y, x.z * z.z, x.w * z.w);
      } else {
        q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const bool y = set_has(a.set, make_edge(min(u, v), max(u, v)));
    float pi_sum = 0.f, probs_sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPL4; ++i) {
      const uint32_t k = lane * 4 + 128 * i;
      if (k < K) {
        const float4 b = *reinterpret_cast<const float4*>(s_beta + k);
        pi_sum += (q[i].x + q[i].y) + (q[i].z + q[i].w);
        // probs_k = y ? beta_k * f : (1 - beta_k) * f   (beta.cc:116-120)
        q[i].x *= y ? b.x : 1.0f - b.x;
        q[i].y *= y ? b.y : 1.0f - b.y;
        q[i].z *= y ? b.z : 1.0f - b.z;
        q[i].w *= y ? b.w : 1.0f - b.w;
        probs_sum += (q[i].x + q[i].y) + (q[i].z + q[i].w);
      }
    }
    pi_sum = warp_sum(pi_sum);
    probs_sum = warp_sum(probs_sum);
    // prob_0 = (y ? EPSILON : 1 - EPSILON) * (1 - pi_sum)   (beta.cc:124-125)
    probs_sum += (y ? a.epsilon : 1.0f - a.epsilon) * (1.0f - pi_sum);
    const float rS = 1.0f / probs_sum;
    if (y) {
#pragma unroll
      for (int i = 0; i < KPL4; ++i) {
        accB[i].x = fmaf(q[i].x, rS, accB[i].x);
        accB[i].y = fmaf(q[i].y, rS, accB[i].y);
        accB[i].z = fmaf(q[i].z, rS, accB[i].z);
        accB[i].w = fmaf(q[i].w, rS, accB[i].w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < KPL4; ++i) {
        accA[i].x = fmaf(q[i].x, rS, accA[i].x);
        accA[i].y = fmaf(q[i].y, rS, accA[i].y);
        accA[i].z = fmaf(q[i].z, rS, accA[i].z);
        accA[i].w = fmaf(q[i].w, rS, accA[i].w);
      }
    }
  }
  // combine the CTA's warps in warp order
  for (uint32_t w = 0; w < BETA_WARPS; ++w) {
    if (wib == w) {
#pragma unroll
      for (int i = 0; i < KPL4; ++i) {
        const uint32_t k = lane * 4 + 128 * i;
        if (k < K) {
          float4* pA = reinterpret_cast<float4*>(s_acc + k);
          float4* pB = reinterpret_cast<float4*>(s_acc + K + k);
          if (w == 0) {
            *pA = accA[i];
            *pB = accB[i];
          } else {
            float4 x = *pA, z = *pB;
            x.x += accA[i].x; x.y += accA[i].y; x.z += accA[i].z; x.w += accA[i].w;
            z.x += accB[i].x; z.y += accB[i].y; z.z += accB[i].z; z.w += accB[i].w;
            *pA = x;
            *pB = z;
          }
        }
      }
    }
    __syncthreads();
  }
  float* out = a.partial + (size_t)blockIdx.x * 2 * K;
  for (uint32_t k = threadIdx.x; k < 2 * K; k += blockDim.x) out[k] = s_acc[k];
}

// K > 1024: one CTA per edge (grid-strided), thread t owns the columns 4t..4t+3 (+1024 i), so
// the [2][K] accumulators stay in registers (a warp-per-edge layout would need 2K/32 registers
// per lane).  The two per-edge sums cross the CTA through shared memory, one barrier per edge
// (double-buffered), added in warp order.  The next edge's rows are loaded before that barrier.
template <int KPT4>
__global__ void __launch_bounds__(256) k_beta_partial_cta(const __grid_constant__ BetaArgs a) {
  __shared__ float s_red[2][2][8];
  const uint32_t K = a.K, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float4 accA[KPT4], accB[KPT4], bk[KPT4];
#pragma unroll
  for (int i = 0; i < KPT4; ++i) {
    accA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    accB[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t k = tid * 4 + 1024 * i;
    bk[i] = k < K ? make_float4(__ldg(&a.beta[2 * k + 1]), __ldg(&a.beta[2 * k + 3]), __ldg(&a.beta[2 * k + 5]),
                                __ldg(&a.beta[2 * k + 7]))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 q[KPT4], nq[KPT4];
  auto load_edge = [&](uint32_t e, float4* dst, uint32_t* pu, uint32_t* pv) {
    const uint64_t edge = __ldg(&a.edges[e]);
    const uint32_t u = (uint32_t)(edge >> 32), v = (uint32_t)(edge & 0xffffffffu);
    *pu = u;
    *pv = v;
    const float* pa = store_row(a.sv, u);
    const float* pb = store_row(a.sv, v);
#pragma unroll
    for (int i = 0; i < KPT4; ++i) {
      const uint32_t k = tid * 4 + 1024 * i;
      if (k < K) {
        const float4 x = ldg_stream4(pa + k);
        const float4 z = ldg_stream4(pb + k);
        dst[i] = make_float4(x.x * z.x, x.y * z.y, x.z * z.z, x.w * z.w);
      } else {
        dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  uint32_t u = 0, v = 0, nu = 0, nv = 0, par = 0;
  uint32_t e = blockIdx.x;
  if (e < a.E_mb) load_edge(e, q, &u, &v);
  for (; e < a.E_mb; e += gridDim.x, par ^= 1) {
    const uint32_t en = e + gridDim.x;
    if (en < a.E_mb) load_edge(en, nq, &nu, &nv);  // in flight across the reduction below
    const bool y = set_has(a.set, make_edge(min(u, v), max(u, v)));
    float pi_sum = 0.f, probs_sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPT4; ++i) {
      pi_sum += (q[i].x + q[i].y) + (q[i].z + q[i].w);
      q[i].x *= y ? bk[i].x : 1.0f - bk[i].x;
      q[i].y *= y ? bk[i].y : 1.0f - bk[i].y;
      q[i].z *= y ? bk[i].z : 1.0f - bk[i].z;
      q[i].w *= y ? bk[i].w : 1.0f - bk[i].w;
      probs_sum += (q[i].x + q[i].y) + (q[i].z + q[i].w);
    }
    pi_sum = warp_sum(pi_sum);
    probs_sum = warp_sum(probs_sum);
    if (lane == 0) {
      s_red[par][0][wid] = pi_sum;
      s_red[par][1][wid] = probs_sum;
    }
    __syncthreads();
    pi_sum = 0.f;
    probs_sum = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      pi_sum += s_red[par][0][ww];
      probs_sum += s_red[par][1][ww];
    }
    probs_sum += (y ? a.epsilon : 1.0f - a.epsilon) * (1.0f - pi_sum);
    const float rS = 1.0f / probs_sum;
    if (y) {
#pragma unroll
      for (int i = 0; i < KPT4; ++i) {
        accB[i].x = fmaf(q[i].x, rS, accB[i].x);
        accB[i].y = fmaf(q[i].y, rS, accB[i].y);
        accB[i].z = fmaf(q[i].z, rS, accB[i].z);
        accB[i].w = fmaf(q[i].w, rS, accB[i].w);
      }
    } else {
#pragma unroll
      for (int i = 0; i < KPT4; ++i) {
        accA[i].x = fmaf(q[i].x, rS, accA[i].x);
        accA[i].y = fmaf(q[i].y, rS, accA[i].y);
        accA[i].z = fmaf(q[i].z, rS, accA[i].z);
        accA[i].w = fmaf(q[i].w, rS, accA[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < KPT4; ++i) q[i] = nq[i];
    u = nu;
    v = nv;
  }
  float* out = a.partial + (size_t)blockIdx.x * 2 * K;
#pragma unroll
  for (int i = 0; i < KPT4; ++i) {
    const uint32_t k = tid * 4 + 1024 * i;
    if (k < K) {
      *reinterpret_cast<float4*>(out + k) = accA[i];
      *reinterpret_cast<float4*>(out + K + k) = accB[i];
    }
  }
}

// generic-K fallback (K not a multiple of 4 or K > 4096): scalar loads, lane-strided
__global__ void __launch_bounds__(BETA_WARPS * 32)
    k_beta_partial_generic(const __grid_constant__ BetaArgs a) {
  extern __shared__ __align__(16) float s_mem[];
  const uint32_t K = a.K;
  const uint32_t WARPS = blockDim.x >> 5;
  float* s_acc = s_mem;  // [WARPS][2][K]
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* accA = s_acc + (size_t)wib * 2 * K;
  float* accB = accA + K;
  for (uint32_t k = lane; k < K; k += 32) {
    accA[k] = 0.f;
    accB[k] = 0.f;
  }
  const uint32_t gwarp = blockIdx.x * WARPS + wib;
  const uint32_t nwarps = gridDim.x * WARPS;
  for (uint32_t e = gwarp; e < a.E_mb; e += nwarps) {
    const uint64_t edge = __ldg(&a.edges[e]);
    const uint32_t u = (uint32_t)(edge >> 32), v = (uint32_t)(edge & 0xffffffffu);
    const float* pa = store_row(a.sv, u);
    const float* pb = store_row(a.sv, v);
    const bool y = set_has(a.set, make_edge(min(u, v), max(u, v)));
    float pi_sum = 0.f, probs_sum = 0.f;
    for (uint32_t k = lane; k < K; k += 32) {
      const float f = pa[k] * pb[k];
      const float b = __ldg(&a.beta[2 * k + 1]);
      pi_sum += f;
      probs_sum += (y ? b : 1.0f - b) * f;
    }
    pi_sum = warp_sum(pi_sum);
    probs_sum = warp_sum(probs_sum);
    probs_sum += (y ? a.epsilon : 1.0f - a.epsilon) * (1.0f - pi_sum);
    const float rS = 1.0f / probs_sum;
    float* acc = y ? accB : accA;
    for (uint32_t k = lane; k < K; k += 32) {
      const float b = __ldg(&a.beta[2 * k + 1]);
      acc[k] = fmaf((y ? b : 1.0f - b) * (pa[k] * pb[k]), rS, acc[k]);
    }
  }
  __syncthreads();
  float* out = a.partial + (size_t)blockIdx.x * 2 * K;
  for (uint32_t k = threadIdx.x; k < 2 * K; k += blockDim.x) {
    float s = s_acc[k];
    for (uint32_t w = 1; w < WARPS; ++w) s += s_acc[(size_t)w * 2 * K + k];
    out[k] = s;
  }
}

// one Langevin step of theta_k and beta_k = theta_k / (theta_k0 + theta_k1): update_theta
// (beta.cc:51-82) + CopyTo(beta) + WG_NORMALIZE_KERNEL rows of 2 (beta.cc:378-379,
// normalize.cc:13-32).  State index k, two normals per k.
__device__ __forceinline__ void theta_step(float* __restrict__ theta, float* __restrict__ beta, float g0, float g1,
                                           uint32_t k, float eps_t, float eta0, float eta1, float scale,
                                           ulonglong2* pool) {
  Rng s = rng_load(pool, k);
  const float half = __fdiv_rn(eps_t, 2.0f);
  const float r0 = rng_randn(s);
  float t0 = theta[2 * k];
  const float f0 = __fsqrt_rn(__fmul_rn(eps_t, t0));
  t0 = fabsf(__fadd_rn(__fadd_rn(t0, __fmul_rn(half, __fadd_rn(__fsub_rn(eta0, t0), __fmul_rn(scale, g0)))),
                       __fmul_rn(f0, r0)));
  t0 = fmaxf(t0, 1e-24f);
  const float r1 = rng_randn(s);
  float t1 = theta[2 * k + 1];
  const float f1 = __fsqrt_rn(__fmul_rn(eps_t, t1));
  t1 = fabsf(__fadd_rn(__fadd_rn(t1, __fmul_rn(half, __fadd_rn(__fsub_rn(eta1, t1), __fmul_rn(scale, g1)))),
                       __fmul_rn(f1, r1)));
  t1 = fmaxf(t1, 1e-24f);
  rng_store(pool, k, s);
  theta[2 * k] = t0;
  theta[2 * k + 1] = t1;
  const float sum = __fadd_rn(__fadd_rn(0.f, t0), t1);
  beta[2 * k] = __fdiv_rn(t0, sum);
  beta[2 * k + 1] = __fdiv_rn(t1, sum);
}

// sum_theta (beta.cc:30-37) + the tail of calculate_grads_partial/sum_grads:
//   g_k0 = A_k (1/theta_k0 - 1/thetaSum_k) - B_k / thetaSum_k
//   g_k1 = B_k (1/theta_k1 - 1/thetaSum_k) - A_k / thetaSum_k        (beta.cc:130-135)
#define BETA_RED_PY 16
struct ThetaStep {  // the update_theta launch that follows, folded into the reduction's last stage
  float* theta;     // nullptr: reduce only (multi-GPU: the gradient is all-reduced first)
  float* beta;
  float eps_t, eta0, eta1, scale;
  ulonglong2* pool;
};
__global__ void __launch_bounds__(32 * BETA_RED_PY)
    k_beta_reduce(const float* __restrict__ partial, uint32_t P, uint32_t K,
                  const float* __restrict__ theta, float* __restrict__ theta_sum,
                  float* __restrict__ grads, const ThetaStep ts_) {
  // 32 consecutive k per CTA; the P partials are split over BETA_RED_PY rows of threads
  // (p = py, py + PY, ...), combined in row order: a fixed association for a given P
  __shared__ float sA[BETA_RED_PY][32], sB[BETA_RED_PY][32];
  const uint32_t kx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const uint32_t k = blockIdx.x * 32 + kx;
  float A = 0.f, B = 0.f;
  if (k < K) {
    for (uint32_t p = py; p < P; p += BETA_RED_PY) {
      A += partial[(size_t)p * 2 * K + k];
      B += partial[(size_t)p * 2 * K + K + k];
    }
  }
  sA[py][kx] = A;
  sB[py][kx] = B;
  __syncthreads();
  if (py != 0 || k >= K) return;
  for (uint32_t r = 1; r < BETA_RED_PY; ++r) {
    A += sA[r][kx];
    B += sB[r][kx];
  }
  const float t0 = theta[2 * k], t1 = theta[2 * k + 1];
  const float ts = __fadd_rn(t0, t1);
  const float rts = __fdiv_rn(1.0f, ts);
  theta_sum[k] = ts;
  const float g0 = A * (__fdiv_rn(1.0f, t0) - rts) + B * (0.0f - rts);
  const float g1 = A * (0.0f - rts) + B * (__fdiv_rn(1.0f, t1) - rts);
  grads[2 * k] = g0;
  grads[2 * k + 1] = g1;
  // theta_k is read and written by this thread only, and beta is no longer read by anyone
  if (ts_.theta != nullptr) theta_step(ts_.theta, ts_.beta, g0, g1, k, ts_.eps_t, ts_.eta0, ts_.eta1, ts_.scale, ts_.pool);
}

// update_theta (beta.cc:51-82) + CopyTo(beta) + WG_NORMALIZE_KERNEL rows of 2
// (beta.cc:378-379, normalize.cc:13-32).  State index k, two normals per k.
__global__ void k_update_theta(float* __restrict__ theta, float* __restrict__ beta,
                               const float* __restrict__ grads, uint32_t K, float eps_t, float eta0,
                               float eta1, float scale, ulonglong2* pool) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  theta_step(theta, beta, grads[2 * k], grads[2 * k + 1], k, eps_t, eta0, eta1, scale, pool);
}

static uint32_t beta_max_ctas(const ammsb_ctx* c) { return (uint32_t)c->sm_count * 3; }

extern "C" int ammsb_beta_workspace_bytes(ammsb_ctx* c, uint32_t K, size_t* bytes) {
  *bytes = sizeof(float) * 2 * (size_t)K * beta_max_ctas(c);
  return 0;
}

static int beta_grads_impl(ammsb_ctx* c, const ammsb_params* p, const float* d_theta,
                           const float* d_beta, ammsb_store* store, ammsb_set* train,
                           const uint64_t* d_edges, uint32_t E_mb, float* d_theta_sum,
                           float* d_grads, void* d_ws, size_t ws_bytes, const ThetaStep& step) {
  AMMSB_REQUIRE(p->K == store->K, "params do not match the store");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  const uint32_t K = p->K;
  uint32_t ctas = (E_mb + BETA_WARPS - 1) / BETA_WARPS;
  if (ctas > beta_max_ctas(c)) ctas = beta_max_ctas(c);
  AMMSB_REQUIRE(ws_bytes >= sizeof(float) * 2 * (size_t)K * (ctas ? ctas : 1), "beta workspace too small");
  if (ctas > 0) {
    BetaArgs a;
    a.sv = store->view();
    a.set = train->view();
    a.beta = d_beta;
    a.edges = d_edges;
    a.E_mb = E_mb;
    a.K = K;
    a.epsilon = p->epsilon;
    a.partial = (float*)d_ws;
    const uint32_t kpl4 = (K + 127) / 128;
    if ((K % 4) == 0 && kpl4 <= 8) {
      const size_t smem = sizeof(float) * 3 * (size_t)K;
      if (kpl4 <= 1) k_beta_partial<1><<<ctas, BETA_WARPS * 32, smem, c->stream>>>(a);
      else if (kpl4 <= 2) k_beta_partial<2><<<ctas, BETA_WARPS * 32, smem, c->stream>>>(a);
      else if (kpl4 <= 4) k_beta_partial<4><<<ctas, BETA_WARPS * 32, smem, c->stream>>>(a);
      else k_beta_partial<8><<<ctas, BETA_WARPS * 32, smem, c->stream>>>(a);
    } else if ((K % 4) == 0 && K <= 4096) {
      // one CTA per edge; the grid is capped by the partial-sum workspace like the warp kernel
      ctas = E_mb < beta_max_ctas(c) ? E_mb : beta_max_ctas(c);
      AMMSB_REQUIRE(ws_bytes >= sizeof(float) * 2 * (size_t)K * ctas, "beta workspace too small");
      if (K <= 2048) k_beta_partial_cta<2><<<ctas, 256, 0, c->stream>>>(a);
      else k_beta_partial_cta<4><<<ctas, 256, 0, c->stream>>>(a);
    } else {
      uint32_t warps = BETA_WARPS;
      while (warps > 1 && sizeof(float) * 2 * (size_t)K * warps > 160 * 1024) warps >>= 1;
      const size_t smem = sizeof(float) * 2 * (size_t)K * warps;
      AMMSB_REQUIRE(smem <= c->smem_optin, "K too large for update_beta");
      AMMSB_CHECK_CUDA(cudaFuncSetAttribute(k_beta_partial_generic,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_beta_partial_generic<<<ctas, warps * 32, smem, c->stream>>>(a);
    }
    AMMSB_LAUNCH_CHECK();
  }
  k_beta_reduce<<<(K + 31) / 32, 32 * BETA_RED_PY, 0, c->stream>>>((const float*)d_ws, ctas, K, d_theta,
                                                        d_theta_sum, d_grads, step);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_beta_grads(ammsb_ctx* c, const ammsb_params* p, const float* d_theta,
                                const float* d_beta, ammsb_store* store, ammsb_set* train,
                                const uint64_t* d_edges, uint32_t E_mb, float* d_theta_sum,
                                float* d_grads, void* d_ws, size_t ws_bytes) {
  ThetaStep none = {};
  return beta_grads_impl(c, p, d_theta, d_beta, store, train, d_edges, E_mb, d_theta_sum, d_grads, d_ws, ws_bytes,
                         none);
}

extern "C" int ammsb_update_theta(ammsb_ctx* c, const ammsb_params* p, float* d_theta, float* d_beta,
                                  const float* d_grads, float scale, uint32_t step_count,
                                  ammsb_rng* pool) {
  AMMSB_REQUIRE(pool && pool->n >= p->K, "beta RNG pool smaller than K");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  k_update_theta<<<(p->K + 127) / 128, 128, 0, c->stream>>>(d_theta, d_beta, d_grads, p->K,
                                                            ammsb_eps_t(p, step_count), p->eta0,
                                                            p->eta1, scale, pool->d_state);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

// single-GPU path: the Langevin step of theta_k runs in the thread that finished the reduction
// of column k (one launch less per iteration; same arithmetic as ammsb_beta_grads followed by
// ammsb_update_theta)
extern "C" int ammsb_update_beta(ammsb_ctx* c, const ammsb_params* p, float* d_theta, float* d_beta,
                                 ammsb_store* store, ammsb_set* train, const uint64_t* d_edges,
                                 uint32_t E_mb, float scale, uint32_t step_count, ammsb_rng* pool,
                                 float* d_theta_sum, float* d_grads, void* d_ws, size_t ws_bytes) {
  AMMSB_REQUIRE(pool && pool->n >= p->K, "beta RNG pool smaller than K");
  ThetaStep step;
  step.theta = d_theta;
  step.beta = d_beta;
  step.eps_t = ammsb_eps_t(p, step_count);
  step.eta0 = p->eta0;
  step.eta1 = p->eta1;
  step.scale = scale;
  step.pool = pool->d_state;
  return beta_grads_impl(c, p, d_theta, d_beta, store, train, d_edges, E_mb, d_theta_sum, d_grads, d_ws, ws_bytes,
                         step);
}
