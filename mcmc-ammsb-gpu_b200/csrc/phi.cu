// phi.cu -- update_phi / update_pi (reference: mcmc/phi.cc).
//
// Two kernels implement update_phi:
//   k_update_phi_fast    production path.  One warp per mini-batch slot, lane l owns
//                        the community indices k = l, l+32, ... (exactly the ownership
//                        of the reference's default launch, WG-NAIVE with phi_wg_size
//                        32), pi rows are staged into shared memory by the TMA engine
//                        (cp.async.bulk + mbarrier ring), one reciprocal per neighbor.
//   k_update_phi_strict  any K / wg / mode; evaluates every expression in the
//                        reference's order with IEEE round-to-nearest ops and no FMA
//                        contraction.  It is the on-device twin of the CPU oracle.
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

struct PhiArgs {
  StoreView sv;
  SetView set;
  const float* beta;
  const uint32_t* nodes;
  const uint32_t* neighbors;
  uint32_t V, n, K;
  uint32_t units;      // reference work-items (THREAD) or groups (WG) that own RNG state
  uint32_t mode, wg;   // reference launch being reproduced
  uint32_t part_index, part_count;  // this rank's share of the units (unit % count == index)
  uint32_t disable_noise;
  float eps_t, alpha, epsilon, Nn;
  ulonglong2* pool;
  float* phi_vec;
  float* phi_sum;
};

// ------------------------------------------------------------------ strict ---

// Serial/tree summation of s_val[0..K) in the reference's association:
// THREAD -> serial in k (phi.cc:100-108); WG -> WG_SUM (sum.cc:20-42).
__device__ float strict_row_sum(const float* s_val, float* s_aux, uint32_t K, uint32_t mode,
                                uint32_t wg) {
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  if (mode == AMMSB_MODE_THREAD) {
    if (tid == 0) {
      float s = 0.f;
      for (uint32_t k = 0; k < K; ++k) s = __fadd_rn(s, s_val[k]);
      s_aux[0] = s;
    }
    __syncthreads();
  } else {
    for (uint32_t vl = tid; vl < wg; vl += T) {
      float ps = 0.f;
      for (uint32_t k = vl; k < K; k += wg) ps = __fadd_rn(ps, s_val[k]);
      s_aux[vl] = ps;
    }
    __syncthreads();
    uint32_t p2 = wg;  // power_of_2(wg) >> 1, sum.cc:11-18
    p2 |= p2 >> 1; p2 |= p2 >> 2; p2 |= p2 >> 4; p2 |= p2 >> 8; p2 |= p2 >> 16;
    p2 = (p2 + 1) >> 1;
    for (; p2 > 0; p2 >>= 1) {
      for (uint32_t lid = tid; lid < p2; lid += T)
        if (lid + p2 < wg) s_aux[lid] = __fadd_rn(s_aux[lid], s_aux[lid + p2]);
      __syncthreads();
    }
  }
  const float r = s_aux[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(128) k_update_phi_strict(const __grid_constant__ PhiArgs a) {
  extern __shared__ float s_mem[];
  const uint32_t K = a.K, tid = threadIdx.x, T = blockDim.x;
  float* s_pa = s_mem;
  float* s_grads = s_pa + K;
  float* s_probs = s_grads + K;
  float* s_aux = s_probs + K;  // max(wg,1)
  const uint32_t vw = a.mode == AMMSB_MODE_THREAD ? 1u : a.wg;
  for (uint32_t unit = a.part_index + a.part_count * blockIdx.x; unit < a.units && unit < a.V;
       unit += a.part_count * gridDim.x) {
    for (uint32_t slot = unit; slot < a.V; slot += a.units) {
      const uint32_t node = a.nodes[slot];
      const float* pi = store_row(a.sv, node);
      const float phi_sum = *store_phi(a.sv, node);
      for (uint32_t k = tid; k < K; k += T) {
        s_pa[k] = pi[k];
        s_grads[k] = 0.f;
      }
      __syncthreads();
      for (uint32_t i = 0; i < a.n; ++i) {
        const uint32_t nb = a.neighbors[(size_t)slot * a.n + i];
        const float* pin = store_row(a.sv, nb);
        const bool y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
        const float e = y ? a.epsilon : __fsub_rn(1.0f, a.epsilon);
        for (uint32_t k = tid; k < K; k += T) {
          const float beta_k = a.beta[2 * k + 1];
          const float f = y ? __fsub_rn(beta_k, a.epsilon) : __fsub_rn(a.epsilon, beta_k);
          s_probs[k] = __fmul_rn(s_pa[k], __fadd_rn(__fmul_rn(pin[k], f), e));
        }
        __syncthreads();
        const float probs_sum = strict_row_sum(s_probs, s_aux, K, a.mode, a.wg);
        for (uint32_t k = tid; k < K; k += T) {
          const float term =
              __fsub_rn(__fdiv_rn(__fdiv_rn(s_probs[k], probs_sum), __fmul_rn(s_pa[k], phi_sum)),
                        __fdiv_rn(1.0f, phi_sum));
          s_grads[k] = __fadd_rn(s_grads[k], term);
        }
        __syncthreads();
      }
      // noise in the reference's per-state draw order
      float* s_noise = s_probs;
      if (a.disable_noise) {
        for (uint32_t k = tid; k < K; k += T) s_noise[k] = 1.0f;
      } else {
        for (uint32_t vl = tid; vl < vw; vl += T) {
          Rng st = rng_load(a.pool, (uint64_t)unit * vw + vl);
          for (uint32_t k = vl; k < K; k += vw) s_noise[k] = rng_randn(st);
          rng_store(a.pool, (uint64_t)unit * vw + vl, st);
        }
      }
      __syncthreads();
      float* out = a.phi_vec + (size_t)slot * K;
      for (uint32_t k = tid; k < K; k += T) {
        const float phi_k = __fmul_rn(s_pa[k], phi_sum);
        const float drift = __fmul_rn(
            __fdiv_rn(a.eps_t, 2.0f),
            __fadd_rn(__fsub_rn(a.alpha, phi_k), __fmul_rn(a.Nn, s_grads[k])));
        const float v = fabsf(__fadd_rn(__fadd_rn(phi_k, drift),
                                        __fmul_rn(__fsqrt_rn(__fmul_rn(a.eps_t, phi_k)), s_noise[k])));
        const float r = fmaxf(v, 1e-24f);
        out[k] = r;
        s_pa[k] = r;  // for the row sum below
      }
      __syncthreads();
      const float sum = strict_row_sum(s_pa, s_aux, K, a.mode, a.wg);
      if (tid == 0) a.phi_sum[slot] = sum;
      __syncthreads();
    }
  }
}

// -------------------------------------------------------------------- fast ---

// ---- Langevin-noise producer warps (warp specialisation) ----
//
// A unit's noise stream is sequential (ziggurat over xorshift128+, ~32 dependent draws per lane
// and slot) and touches no memory: when the warp that gathers rows draws it itself it has no
// loads in flight for ~12% of its time.  With NW > 0 the CTA carries NW extra warps that only
// draw noise: producer p serves compute warps p, p+NW, ..., walks exactly their unit/slot
// sequences (so state i is advanced by the same draws as in the reference launch), and hands
// each slot's K normals over through a shared-memory row guarded by a full/empty mbarrier pair.
// The compute warps then never stall on the RNG and keep their TMA ring busy.
template <int WARPS, int NW, class Tab>
__device__ __forceinline__ void noise_producer(const PhiArgs& a, uint32_t p, float* s_nz, uint64_t* full,
                                               uint64_t* empty, uint32_t lane, const Tab zig) {
  constexpr int PER = WARPS / NW;  // compute warps served by this producer
  const uint32_t K = a.K;
  const bool fast_noise = (a.mode == AMMSB_MODE_WG && a.wg == 32);
  const uint32_t vw = a.mode == AMMSB_MODE_THREAD ? 1u : a.wg;
  const uint32_t active = a.units < a.V ? a.units : a.V;
  const uint32_t total_warps = gridDim.x * WARPS;
  uint32_t unit[PER], slot[PER], item[PER];
  Rng st[PER];
  bool open[PER];  // a unit is in progress (its state is live in st)
#pragma unroll
  for (int c = 0; c < PER; ++c) {
    const uint32_t w = p + c * NW;
    unit[c] = a.part_index + a.part_count * (blockIdx.x * WARPS + w);
    slot[c] = unit[c];
    item[c] = 0;
    open[c] = false;
  }
  for (;;) {
    bool any = false;
#pragma unroll
    for (int c = 0; c < PER; ++c) {
      if (unit[c] >= active) continue;
      any = true;
      const uint32_t w = p + c * NW;
      float* out = s_nz + (size_t)w * K;
      mbar_wait(&empty[w], (item[c] & 1) ^ 1);  // the consumer is done with the previous row
      if (fast_noise) {
        if (!open[c]) {
          st[c] = rng_load(a.pool, (uint64_t)unit[c] * 32 + lane);
          open[c] = true;
        }
        for (uint32_t k = lane; k < K; k += 32) out[k] = rng_randn_t(st[c], zig);
      } else {
        for (uint32_t vl = lane; vl < vw; vl += 32) {
          Rng vs = rng_load(a.pool, (uint64_t)unit[c] * vw + vl);
          for (uint32_t k = vl; k < K; k += vw) out[k] = rng_randn_t(vs, zig);
          rng_store(a.pool, (uint64_t)unit[c] * vw + vl, vs);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[w]);  // release: the row is visible to the consumer
      ++item[c];
      slot[c] += a.units;
      if (slot[c] >= a.V) {  // unit finished: persist its state, move to the warp's next unit
        if (fast_noise) rng_store(a.pool, (uint64_t)unit[c] * 32 + lane, st[c]);
        open[c] = false;
        unit[c] += a.part_count * total_warps;
        slot[c] = unit[c];
      }
    }
    if (!any) break;
  }
}

// NB = neighbors per loop trip.  For short rows (K <= 512) the per-neighbor fixed cost (barrier
// wait, shuffle tree, reciprocal, refill) dominates the issue slots; NB = 2 interleaves two
// neighbors (two shuffle trees in flight, own row kept in registers) without changing any
// result: every sum is formed in the same order as with NB = 1.
// EARLY (NW == 0 only): the gather warp draws a slot's noise itself, but at the START of the
// slot -- right after it has requested the own row and the first STAGES neighbor rows, while it
// would otherwise just wait for them -- into a shared-memory row of its own, instead of after the
// neighbor loop with nothing in flight.  Same draws in the same order (a unit's slots are visited
// in order either way).
template <int KPL, int STAGES, int WARPS, bool EXACT, int NB, int NW, bool EARLY>
__global__ void __launch_bounds__((WARPS + NW) * 32)
    k_update_phi_fast(const __grid_constant__ PhiArgs a) {
  static_assert(!(EARLY && NW > 0), "EARLY is the producer-less way to take the noise off the critical path");
  constexpr bool NZ_ROWS = NW > 0 || EARLY;  // a noise row per gather warp
  extern __shared__ __align__(128) unsigned char s_raw[];
  const uint32_t K = a.K;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t row_bytes = K * 4;
  // layout: [WARPS][(own + STAGES rows)] | NZ_ROWS ? [WARPS][noise row] | [WARPS][STAGES+1] stage
  // barriers | NW ? full[WARPS], empty[WARPS] | ziggurat tables
  float* s_nz = reinterpret_cast<float*>(s_raw) + (size_t)WARPS * (STAGES + 1) * K;
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(s_raw + ((size_t)WARPS * (STAGES + 1) + (NZ_ROWS ? WARPS : 0)) * row_bytes);
  uint64_t* nz_full = bar_base + WARPS * (STAGES + 1);
  uint64_t* nz_empty = nz_full + WARPS;
  // ziggurat tables, after the last barrier word
  uint32_t* s_zig = reinterpret_cast<uint32_t*>(bar_base + WARPS * (STAGES + 1) + (NW ? 2 * WARPS : 0));
  const ZigShared zig{s_zig};
  const bool ws_noise = NW > 0 && !a.disable_noise;
  if (NW == 0) zig_stage(s_zig);  // gather warps that draw their own noise; producers read the global tables
  if (NW > 0 && threadIdx.x == 0) {
    for (int w = 0; w < WARPS; ++w) {
      mbar_init(&nz_full[w], 1);
      mbar_init(&nz_empty[w], 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (NW > 0 && wib >= WARPS) {  // producer warps: noise only, no row traffic
    // the producers read the tables through the read-only path: at K = 1024 shared memory is the
    // busy resource (measured: 0.3838 ms with the shared copy, 0.3798 ms without)
    if (ws_noise)
      noise_producer<WARPS, (NW > 0 ? NW : 1)>(a, wib - WARPS, s_nz, nz_full, nz_empty, lane, ZigGlobal());
    return;
  }
  float* s_own = reinterpret_cast<float*>(s_raw) + (size_t)wib * (STAGES + 1) * K;
  float* s_stage = s_own + K;
  uint64_t* bars = bar_base + wib * (STAGES + 1);
  if (lane == 0) {
    for (int s = 0; s <= STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  uint32_t phase = 0;  // bit s: parity to wait on for barrier s (bit STAGES = own row)
  uint32_t nz_item = 0;  // noise rows consumed so far (parity of the full barrier)

  // f_k = beta_k - epsilon for the lane's k (phi.cc:237-239); the non-link factor is -f_k
  float fb[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const uint32_t k = lane + 32 * i;
    fb[i] = (EXACT || k < K) ? __ldg(&a.beta[2 * k + 1]) - a.epsilon : 0.f;
  }
  const float e_link = a.epsilon, e_non = 1.0f - a.epsilon;
  const float half_eps = a.eps_t / 2;
  const bool fast_noise = (a.mode == AMMSB_MODE_WG && a.wg == 32);
  const uint32_t vw = a.mode == AMMSB_MODE_THREAD ? 1u : a.wg;

  const uint32_t gwarp = blockIdx.x * WARPS + wib;
  const uint32_t total_warps = gridDim.x * WARPS;
  for (uint32_t unit = a.part_index + a.part_count * gwarp; unit < a.units && unit < a.V;
       unit += a.part_count * total_warps) {
    Rng st;
    if (NW == 0 && fast_noise && !a.disable_noise) st = rng_load(a.pool, (uint64_t)unit * 32 + lane);
    for (uint32_t slot = unit; slot < a.V; slot += a.units) {
      const uint32_t node = __ldg(&a.nodes[slot]);
      const float phi_sum = *store_phi(a.sv, node);
      const float rphi = 1.0f / phi_sum;
      __syncwarp();  // previous slot's reads of s_own / stages are complete
      if (lane == 0) {
        mbar_expect_tx(&bars[STAGES], row_bytes);
        bulk_g2s(s_own, store_row(a.sv, node), row_bytes, &bars[STAGES]);
      }
      // neighbor chunk registers: cur = neighbors [c*32, c*32+32), nxt = the next 32
      const uint32_t* nbr = a.neighbors + (size_t)slot * a.n;
      const float* cur_ptr = nullptr;
      const float* nxt_ptr = nullptr;
      uint32_t cur_mask = 0, nxt_mask = 0;
      {
        bool y = false;
        if (lane < a.n) {
          const uint32_t nb = __ldg(&nbr[lane]);
          cur_ptr = store_row(a.sv, nb);
          if (lane < STAGES) {
            mbar_expect_tx(&bars[lane], row_bytes);
            bulk_g2s(s_stage + (size_t)lane * K, cur_ptr, row_bytes, &bars[lane]);
          }
          y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
        }
        cur_mask = __ballot_sync(FULL_MASK, y);
        y = false;
        if (32 + lane < a.n) {
          const uint32_t nb = __ldg(&nbr[32 + lane]);
          nxt_ptr = store_row(a.sv, nb);
          y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
        }
        nxt_mask = __ballot_sync(FULL_MASK, y);
      }
      if (EARLY && !a.disable_noise) {  // the rows requested above are in flight meanwhile
        float* nz = s_nz + (size_t)wib * K;
        if (fast_noise) {
          for (uint32_t k = lane; k < K; k += 32) nz[k] = rng_randn_t(st, zig);  // read back by this lane only
        } else {
          for (uint32_t vl = lane; vl < vw; vl += 32) {
            Rng vs = rng_load(a.pool, (uint64_t)unit * vw + vl);
            for (uint32_t k = vl; k < K; k += vw) nz[k] = rng_randn_t(vs, zig);
            rng_store(a.pool, (uint64_t)unit * vw + vl, vs);
          }
          __syncwarp();
        }
      }
      float g[KPL];
#pragma unroll
      for (int i = 0; i < KPL; ++i) g[i] = 0.f;

      mbar_wait(&bars[STAGES], (phase >> STAGES) & 1);
      phase ^= 1u << STAGES;

      float own[NB == 2 ? KPL : 1];
      if (NB == 2) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
          const uint32_t k = lane + 32 * i;
          own[NB == 2 ? i : 0] = (EXACT || k < K) ? s_own[k] : 0.f;
        }
      }
      for (uint32_t j = 0; j < a.n; j += NB) {
        const uint32_t jj = j & 31;
        if (jj == 0 && j > 0) {
          cur_ptr = nxt_ptr;
          cur_mask = nxt_mask;
          bool y = false;
          nxt_ptr = nullptr;
          if (j + 32 + lane < a.n) {
            const uint32_t nb = __ldg(&nbr[j + 32 + lane]);
            nxt_ptr = store_row(a.sv, nb);
            y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
          }
          nxt_mask = __ballot_sync(FULL_MASK, y);
        }
        float t[NB][KPL];
        float S[NB];
        bool live[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          live[u] = (u == 0) || (j + u < a.n);
          S[u] = 0.f;
          if (live[u]) {
            const uint32_t s = (j + u) % STAGES;
            const bool y = (cur_mask >> (jj + u)) & 1;
            const float e = y ? e_link : e_non;
            mbar_wait(&bars[s], (phase >> s) & 1);
            phase ^= 1u << s;
            const float* row = s_stage + (size_t)s * K;
            if (y) {  // warp-uniform: the sign of f_k folds into the FMA
#pragma unroll
              for (int i = 0; i < KPL; ++i) {
                const uint32_t k = lane + 32 * i;
                if (EXACT || k < K) {
                  t[u][i] = fmaf(row[k], fb[i], e);
                  S[u] = fmaf(NB == 2 ? own[NB == 2 ? i : 0] : s_own[k], t[u][i], S[u]);
                } else {
                  t[u][i] = 0.f;
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < KPL; ++i) {
                const uint32_t k = lane + 32 * i;
                if (EXACT || k < K) {
                  t[u][i] = fmaf(row[k], -fb[i], e);
                  S[u] = fmaf(NB == 2 ? own[NB == 2 ? i : 0] : s_own[k], t[u][i], S[u]);
                } else {
                  t[u][i] = 0.f;
                }
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < KPL; ++i) t[u][i] = 0.f;
          }
        }
        __syncwarp();  // every lane has consumed the stage(s) -> refill
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const uint32_t q = j + u + STAGES;
          if (live[u] && q < a.n && lane == (q & 31)) {
            const uint32_t s = (j + u) % STAGES;
            const float* src = ((q >> 5) == (j >> 5)) ? cur_ptr : nxt_ptr;
            mbar_expect_tx(&bars[s], row_bytes);
            bulk_g2s(s_stage + (size_t)s * K, src, row_bytes, &bars[s]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < NB; ++u) S[u] += __shfl_xor_sync(FULL_MASK, S[u], o);
        }
        // (probs_k / probs_sum) / (pi_k * phi_sum) - 1 / phi_sum with probs_k / pi_k = t_k
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          if (live[u]) {
            const float inv = 1.0f / (S[u] * phi_sum);
#pragma unroll
            for (int i = 0; i < KPL; ++i) g[i] += fmaf(t[u][i], inv, -rphi);
          }
        }
      }

      // Langevin noise (phi.cc:266-274) in the reference's per-state draw order
      float* s_noise = NZ_ROWS ? s_nz + (size_t)wib * K : s_stage;  // else: all stages are drained here
      if (NW > 0) {
        if (ws_noise) mbar_wait(&nz_full[wib], nz_item & 1);  // the producer's row for this slot
      } else if (!EARLY && !a.disable_noise) {
        if (fast_noise) {
          for (uint32_t k = lane; k < K; k += 32) s_noise[k] = rng_randn_t(st, zig);
        } else {
          for (uint32_t vl = lane; vl < vw; vl += 32) {
            Rng vs = rng_load(a.pool, (uint64_t)unit * vw + vl);
            for (uint32_t k = vl; k < K; k += vw) s_noise[k] = rng_randn_t(vs, zig);
            rng_store(a.pool, (uint64_t)unit * vw + vl, vs);
          }
          __syncwarp();
        }
      }
      float* out = a.phi_vec + (size_t)slot * K;
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const uint32_t k = lane + 32 * i;
        if (EXACT || k < K) {
          const float noise = a.disable_noise ? 1.0f : s_noise[k];
          const float v = phi_langevin(s_own[k], phi_sum, g[i], noise, half_eps, a.eps_t, a.alpha, a.Nn);
          out[k] = v;
          lsum += v;
        }
      }
      if (NW > 0 && ws_noise) {
        __syncwarp();  // every lane has read the noise row
        if (lane == 0) mbar_arrive(&nz_empty[wib]);
        ++nz_item;
      }
      lsum = warp_sum(lsum);
      if (lane == 0) a.phi_sum[slot] = lsum;
    }
    // with producer warps (NW > 0) the state lives in, and is persisted by, the producer
    if (NW == 0 && fast_noise && !a.disable_noise) rng_store(a.pool, (uint64_t)unit * 32 + lane, st);
  }
}

// ------------------------------------------------------------------ team ---
//
// K > 1024: a slot is processed by a TEAM of T warps (T = 2 or 4).  Warp w owns the column
// segment [w*KS, (w+1)*KS), KS = K/T <= 1024, and stages only that segment of every row, so the
// per-warp register budget and the bytes in flight per SM are those of the K = 1024 kernel.
// The one cross-warp dependency per neighbor -- sum_k probs_k -- goes through shared memory and
// a named barrier of the team (partials added in warp order: a fixed association).
// Langevin noise comes from 2 producer warps per CTA (see noise_producer above): a unit's stream
// is sequential over all K columns, so one producer draws a whole row; with T = 4 the two
// producers alternate slots, with T = 2 each serves one team.  Rows are double-buffered in
// shared memory behind full/empty mbarriers.  One slot per unit only (V <= 65535); larger V
// falls back to the strict kernel.
__device__ __forceinline__ void team_barrier(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int KPL, int STAGES, int T, bool EXACT>
__global__ void __launch_bounds__(192) k_update_phi_team(const __grid_constant__ PhiArgs a) {
  constexpr int WARPS = 4;          // gather warps
  constexpr int NW = 2;             // noise producer warps
  constexpr int TEAMS = WARPS / T;
  constexpr int PPT = NW / TEAMS;   // producers per team
  extern __shared__ __align__(128) unsigned char s_raw[];
  const uint32_t K = a.K, KS = K / T;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t seg_bytes = KS * 4;
  // layout: [WARPS][(own + STAGES) segments] | [TEAMS][2][K] noise | stage barriers
  // [WARPS][STAGES+1] | full[TEAMS][2] | empty[TEAMS][2] | partial sums [TEAMS][3T] floats
  float* s_nz_all = reinterpret_cast<float*>(s_raw) + (size_t)WARPS * (STAGES + 1) * KS;
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(s_nz_all + (size_t)TEAMS * 2 * K);
  uint64_t* nz_full_all = bar_base + WARPS * (STAGES + 1);
  uint64_t* nz_empty_all = nz_full_all + TEAMS * 2;
  float* s_part_all = reinterpret_cast<float*>(nz_empty_all + TEAMS * 2);
  const bool ws_noise = !a.disable_noise;
  if (threadIdx.x == 0) {
    for (int i = 0; i < TEAMS * 2; ++i) {
      mbar_init(&nz_full_all[i], 1);
      mbar_init(&nz_empty_all[i], T);  // every gather warp of the team releases the row
    }
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t active = a.units < a.V ? a.units : a.V;
  const uint32_t total_teams = gridDim.x * TEAMS;

  if (wib >= WARPS) {
    // ---- producer: items i = q, q + PPT, ... of one team; row buffer i & 1 ----
    if (!ws_noise) return;
    const uint32_t p = wib - WARPS;
    const uint32_t team = p / PPT, q = p % PPT;
    const uint32_t gteam = blockIdx.x * TEAMS + team;
    const bool fast_noise = (a.mode == AMMSB_MODE_WG && a.wg == 32);
    const uint32_t vw = a.mode == AMMSB_MODE_THREAD ? 1u : a.wg;
    for (uint32_t i = q;; i += PPT) {
      const uint32_t unit = a.part_index + a.part_count * (gteam + i * total_teams);
      if (unit >= active) break;
      const uint32_t b = i & 1;
      float* out = s_nz_all + ((size_t)team * 2 + b) * K;
      mbar_wait(&nz_empty_all[team * 2 + b], ((i >> 1) & 1) ^ 1);
      if (fast_noise) {
        Rng st = rng_load(a.pool, (uint64_t)unit * 32 + lane);
        for (uint32_t k = lane; k < K; k += 32) out[k] = rng_randn(st);
        rng_store(a.pool, (uint64_t)unit * 32 + lane, st);
      } else {
        for (uint32_t vl = lane; vl < vw; vl += 32) {
          Rng vs = rng_load(a.pool, (uint64_t)unit * vw + vl);
          for (uint32_t k = vl; k < K; k += vw) out[k] = rng_randn(vs);
          rng_store(a.pool, (uint64_t)unit * vw + vl, vs);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&nz_full_all[team * 2 + b]);
    }
    return;
  }

  const uint32_t team = wib / T, w = wib % T;
  const uint32_t seg = w * KS;  // first column of this warp's segment
  float* s_own = reinterpret_cast<float*>(s_raw) + (size_t)wib * (STAGES + 1) * KS;
  float* s_stage = s_own + KS;
  uint64_t* bars = bar_base + wib * (STAGES + 1);
  float* s_part = s_part_all + team * 3 * T;  // [2][T] + [T]
  if (lane == 0) {
    for (int s = 0; s <= STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  uint32_t phase = 0;
  const uint32_t bar_id = 1 + team, bar_threads = T * 32;

  float fb[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    const uint32_t kk = lane + 32 * i;
    fb[i] = (EXACT || kk < KS) ? __ldg(&a.beta[2 * (seg + kk) + 1]) - a.epsilon : 0.f;
  }
  const float e_link = a.epsilon, e_non = 1.0f - a.epsilon;
  const float half_eps = a.eps_t / 2;
  const uint32_t gteam = blockIdx.x * TEAMS + team;

  for (uint32_t it = 0;; ++it) {
    const uint32_t slot = a.part_index + a.part_count * (gteam + it * total_teams);  // slot == unit
    if (slot >= active) break;  // team-uniform
    const uint32_t node = __ldg(&a.nodes[slot]);
    const float phi_sum = *store_phi(a.sv, node);
    const float rphi = 1.0f / phi_sum;
    __syncwarp();
    if (lane == 0) {
      mbar_expect_tx(&bars[STAGES], seg_bytes);
      bulk_g2s(s_own, store_row(a.sv, node) + seg, seg_bytes, &bars[STAGES]);
    }
    const uint32_t* nbr = a.neighbors + (size_t)slot * a.n;
    const float* cur_ptr = nullptr;
    const float* nxt_ptr = nullptr;
    uint32_t cur_mask = 0, nxt_mask = 0;
    {
      bool y = false;
      if (lane < a.n) {
        const uint32_t nb = __ldg(&nbr[lane]);
        cur_ptr = store_row(a.sv, nb) + seg;
        if (lane < STAGES) {
          mbar_expect_tx(&bars[lane], seg_bytes);
          bulk_g2s(s_stage + (size_t)lane * KS, cur_ptr, seg_bytes, &bars[lane]);
        }
        y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
      }
      cur_mask = __ballot_sync(FULL_MASK, y);
      y = false;
      if (32 + lane < a.n) {
        const uint32_t nb = __ldg(&nbr[32 + lane]);
        nxt_ptr = store_row(a.sv, nb) + seg;
        y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
      }
      nxt_mask = __ballot_sync(FULL_MASK, y);
    }
    float g[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) g[i] = 0.f;
    mbar_wait(&bars[STAGES], (phase >> STAGES) & 1);
    phase ^= 1u << STAGES;

    for (uint32_t j = 0; j < a.n; ++j) {
      const uint32_t q32 = j & 31;
      if (q32 == 0 && j > 0) {
        cur_ptr = nxt_ptr;
        cur_mask = nxt_mask;
        bool y = false;
        nxt_ptr = nullptr;
        if (j + 32 + lane < a.n) {
          const uint32_t nb = __ldg(&nbr[j + 32 + lane]);
          nxt_ptr = store_row(a.sv, nb) + seg;
          y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
        }
        nxt_mask = __ballot_sync(FULL_MASK, y);
      }
      const uint32_t s = j % STAGES;
      const bool y = (cur_mask >> q32) & 1;
      const float e = y ? e_link : e_non;
      mbar_wait(&bars[s], (phase >> s) & 1);
      phase ^= 1u << s;
      const float* row = s_stage + (size_t)s * KS;
      float t[KPL];
      float S = 0.f;
      if (y) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
          const uint32_t kk = lane + 32 * i;
          if (EXACT || kk < KS) {
            t[i] = fmaf(row[kk], fb[i], e);
            S = fmaf(s_own[kk], t[i], S);
          } else {
            t[i] = 0.f;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
          const uint32_t kk = lane + 32 * i;
          if (EXACT || kk < KS) {
            t[i] = fmaf(row[kk], -fb[i], e);
            S = fmaf(s_own[kk], t[i], S);
          } else {
            t[i] = 0.f;
          }
        }
      }
      __syncwarp();
      {
        const uint32_t q = j + STAGES;
        if (q < a.n && lane == (q & 31)) {
          const float* src = ((q >> 5) == (j >> 5)) ? cur_ptr : nxt_ptr;
          mbar_expect_tx(&bars[s], seg_bytes);
          bulk_g2s(s_stage + (size_t)s * KS, src, seg_bytes, &bars[s]);
        }
      }
      S = warp_sum(S);
      float* part = s_part + (j & 1) * T;
      if (lane == 0) part[w] = S;
      team_barrier(bar_id, bar_threads);
      S = 0.f;
#pragma unroll
      for (int ww = 0; ww < T; ++ww) S += part[ww];
      const float inv = 1.0f / (S * phi_sum);
#pragma unroll
      for (int i = 0; i < KPL; ++i) g[i] += fmaf(t[i], inv, -rphi);
    }

    // the slot's noise row (all K columns) from the producer; this warp reads its segment
    const uint32_t b = it & 1;
    const float* nz = s_nz_all + ((size_t)team * 2 + b) * K + seg;
    if (ws_noise) mbar_wait(&nz_full_all[team * 2 + b], (it >> 1) & 1);
    float* out = a.phi_vec + (size_t)slot * K + seg;
    float lsum = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const uint32_t kk = lane + 32 * i;
      if (EXACT || kk < KS) {
        const float noise = a.disable_noise ? 1.0f : nz[kk];
        const float v = phi_langevin(s_own[kk], phi_sum, g[i], noise, half_eps, a.eps_t, a.alpha, a.Nn);
        out[kk] = v;
        lsum += v;
      }
    }
    if (ws_noise) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&nz_empty_all[team * 2 + b]);
    }
    lsum = warp_sum(lsum);
    float* part = s_part + 2 * T;
    if (lane == 0) part[w] = lsum;
    team_barrier(bar_id, bar_threads);
    if (w == 0 && lane == 0) {
      float tot = 0.f;
      for (int ww = 0; ww < T; ++ww) tot += part[ww];
      a.phi_sum[slot] = tot;
    }
  }
}

// ------------------------------------------------------------------ split ---
//
// Few slots (link mini-batches: V = 1 + deg(u), a handful): the launch is a latency chain, not
// bandwidth -- with one warp per slot the n neighbors are handled one after the other by a
// single warp while 140 SMs idle.  Here a slot gets a whole CTA: its n neighbors are dealt
// round-robin to WPS gather warps (warp w takes neighbors w, w + WPS, ...; a ring of R rows in
// flight per warp, so all n rows of a 32-neighbor slot are requested at once), one more warp
// draws the slot's Langevin noise meanwhile, and the per-warp gradient partials are added in
// warp order (a fixed association, but not the neighbor-by-neighbor one of the kernels above:
// results agree with them to fp32 rounding, not bit for bit -- which is why this kernel is chosen
// by the mini-batch's slot count alone, never by the share a rank owns).  One slot per unit.
template <int KPL, int WPS, int R, bool EXACT>
__global__ void __launch_bounds__((WPS + 1) * 32) k_update_phi_split(const __grid_constant__ PhiArgs a) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const uint32_t K = a.K;
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t row_bytes = K * 4;
  // layout: own row | noise row (later: the new phi row) | [WPS][R] stage rows | bar_own,
  // [WPS][R] stage barriers
  float* s_own = reinterpret_cast<float*>(s_raw);
  float* s_noise = s_own + K;
  float* s_stage_all = s_noise + K;
  uint64_t* bar_own = reinterpret_cast<uint64_t*>(s_stage_all + (size_t)WPS * R * K);
  uint64_t* bars_all = bar_own + 1;
  uint32_t* s_zig = reinterpret_cast<uint32_t*>(bars_all + WPS * R);  // ziggurat tables
  const ZigShared zig{s_zig};
  const uint32_t slot = a.part_index + a.part_count * blockIdx.x;  // slot == unit
  if (slot >= a.V || slot >= a.units) return;                      // CTA-uniform
  const uint32_t node = __ldg(&a.nodes[slot]);
  zig_stage(s_zig);
  if (threadIdx.x == 0) {
    mbar_init(bar_own, 1);
    for (int i = 0; i < WPS * R; ++i) mbar_init(&bars_all[i], 1);
    mbar_fence_init();
    mbar_expect_tx(bar_own, row_bytes);
    bulk_g2s(s_own, store_row(a.sv, node), row_bytes, bar_own);
  }
  __syncthreads();
  const float phi_sum = *store_phi(a.sv, node);

  if (wib == WPS) {
    // ---- noise warp: the unit's stream in the reference's per-state draw order ----
    if (!a.disable_noise) {
      if (a.mode == AMMSB_MODE_WG && a.wg == 32) {
        Rng st = rng_load(a.pool, (uint64_t)slot * 32 + lane);
        for (uint32_t k = lane; k < K; k += 32) s_noise[k] = rng_randn_t(st, zig);
        rng_store(a.pool, (uint64_t)slot * 32 + lane, st);
      } else {
        const uint32_t vw = a.mode == AMMSB_MODE_THREAD ? 1u : a.wg;
        for (uint32_t vl = lane; vl < vw; vl += 32) {
          Rng vs = rng_load(a.pool, (uint64_t)slot * vw + vl);
          for (uint32_t k = vl; k < K; k += vw) s_noise[k] = rng_randn_t(vs, zig);
          rng_store(a.pool, (uint64_t)slot * vw + vl, vs);
        }
      }
    }
  } else {
    // ---- gather warp w: neighbors w, w + WPS, ... (lane i holds the i-th of them) ----
    const uint32_t w = wib;
    float* s_stage = s_stage_all + (size_t)w * R * K;
    uint64_t* bars = bars_all + w * R;
    const uint32_t cnt = a.n > w ? (a.n - w + WPS - 1) / WPS : 0;  // <= 32 (checked by the host)
    const float* my_ptr = nullptr;
    bool y = false;
    if (lane < cnt) {
      const uint32_t nb = __ldg(&a.neighbors[(size_t)slot * a.n + w + lane * WPS]);
      my_ptr = store_row(a.sv, nb);
      if (lane < R) {
        mbar_expect_tx(&bars[lane], row_bytes);
        bulk_g2s(s_stage + (size_t)lane * K, my_ptr, row_bytes, &bars[lane]);
      }
      y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
    }
    const uint32_t mask = __ballot_sync(FULL_MASK, y);
    float fb[KPL], g[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const uint32_t k = lane + 32 * i;
      fb[i] = (EXACT || k < K) ? __ldg(&a.beta[2 * k + 1]) - a.epsilon : 0.f;
      g[i] = 0.f;
    }
    const float e_link = a.epsilon, e_non = 1.0f - a.epsilon;
    const float rphi = 1.0f / phi_sum;
    uint32_t phase = 0;
    mbar_wait(bar_own, 0);
    for (uint32_t j = 0; j < cnt; ++j) {
      const uint32_t s = j % R;
      const bool link = (mask >> j) & 1;
      const float e = link ? e_link : e_non;
      const float sgn = link ? 1.0f : -1.0f;
      mbar_wait(&bars[s], (phase >> s) & 1);
      phase ^= 1u << s;
      const float* row = s_stage + (size_t)s * K;
      float t[KPL];
      float S = 0.f;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const uint32_t k = lane + 32 * i;
        if (EXACT || k < K) {
          t[i] = fmaf(row[k], sgn * fb[i], e);
          S = fmaf(s_own[k], t[i], S);
        } else {
          t[i] = 0.f;
        }
      }
      __syncwarp();  // every lane has consumed the stage -> refill
      if (j + R < cnt && lane == j + R) {
        mbar_expect_tx(&bars[s], row_bytes);
        bulk_g2s(s_stage + (size_t)s * K, my_ptr, row_bytes, &bars[s]);
      }
      S = warp_sum(S);
      const float inv = 1.0f / (S * phi_sum);
#pragma unroll
      for (int i = 0; i < KPL; ++i) g[i] += fmaf(t[i], inv, -rphi);
    }
    // partial gradient of this warp into its (drained) first stage row
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
      const uint32_t k = lane + 32 * i;
      if (EXACT || k < K) s_stage[k] = g[i];
    }
  }
  __syncthreads();  // partials and the noise row are complete
  if (wib == WPS) mbar_wait(bar_own, 0);  // the noise warp has not yet observed the own-row copy

  // Langevin step (phi.cc:266-274): every thread of the CTA takes columns t, t + threads, ...
  const float half_eps = a.eps_t / 2;
  float* out = a.phi_vec + (size_t)slot * K;
  for (uint32_t k = threadIdx.x; k < K; k += (WPS + 1) * 32) {
    float gk = 0.f;
#pragma unroll
    for (int w = 0; w < WPS; ++w) gk += s_stage_all[(size_t)w * R * K + k];
    const float noise = a.disable_noise ? 1.0f : s_noise[k];
    const float v = phi_langevin(s_own[k], phi_sum, gk, noise, half_eps, a.eps_t, a.alpha, a.Nn);
    out[k] = v;
    s_noise[k] = v;  // this thread's own column: no other reader of the noise value
  }
  __syncthreads();
  if (wib == 0) {  // row sum in the association of the one-warp-per-slot kernels (= WG_SUM, wg 32)
    float lsum = 0.f;
    for (uint32_t k = lane; k < K; k += 32) lsum += s_noise[k];
    lsum = warp_sum(lsum);
    if (lane == 0) a.phi_sum[slot] = lsum;
  }
}

// cudaFuncSetAttribute + the occupancy query cost microseconds of host time per launch; their
// result depends only on (kernel, device, block, smem), so it is computed once per combination.
template <class Kern>
static int resident_ctas_per_sm(Kern kern, int device, int block, size_t smem, int* occ_out) {
  struct Entry { const void* fn; int device, block; size_t smem; int occ; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  const void* fn = reinterpret_cast<const void*>(kern);
  {
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
      if (e.fn == fn && e.device == device && e.block == block && e.smem == smem) {
        *occ_out = e.occ;
        return 0;
      }
  }
  // the limit is a property of the function, not of one launch: it is raised to the device
  // maximum once and never lowered (a kernel is launched with different row sizes, and a cached
  // entry must stay launchable after a smaller one was added)
  int optin = 0;
  AMMSB_CHECK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  AMMSB_REQUIRE(smem <= (size_t)optin, "update_phi: shared memory request exceeds the device limit");
  AMMSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  int occ = 0;
  AMMSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
  std::lock_guard<std::mutex> lock(mu);
  cache.push_back(Entry{fn, device, block, smem, occ});
  *occ_out = occ;
  return 0;
}

// number of units (and so of concurrently useful warps / CTAs) this rank owns
static uint32_t my_units(const PhiArgs& a) {
  const uint32_t active = a.units < a.V ? a.units : a.V;
  return active > a.part_index ? (active - a.part_index + a.part_count - 1) / a.part_count : 0;
}

template <int KPL, int STAGES, int WARPS, int NB = 1, int NW = 0, bool EARLY = false>
static int launch_fast(ammsb_ctx* c, const PhiArgs& a) {
  const size_t smem = ((size_t)WARPS * (STAGES + 1) + ((NW || EARLY) ? WARPS : 0)) * a.K * 4 +
                      (size_t)WARPS * (STAGES + 1) * 8 + (NW ? 2 * WARPS * 8 : ZIG_WORDS * 4);
  const bool exact = (a.K == 32u * KPL);
  auto kern = exact ? k_update_phi_fast<KPL, STAGES, WARPS, true, NB, NW, EARLY>
                    : k_update_phi_fast<KPL, STAGES, WARPS, false, NB, NW, EARLY>;
  int occ = 0;
  if (resident_ctas_per_sm(kern, c->device, (WARPS + NW) * 32, smem, &occ)) return 1;
  AMMSB_REQUIRE(occ > 0, "update_phi: kernel does not fit on an SM");
  const uint32_t active = my_units(a);
  if (active == 0) return 0;
  uint32_t blocks = (active + WARPS - 1) / WARPS;
  const uint32_t resident = (uint32_t)occ * c->sm_count;
  if (blocks > resident) blocks = resident;  // persistent: one wave, warps stride over units
  kern<<<blocks, (WARPS + NW) * 32, smem, c->stream>>>(a);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

template <int KPL>
static int launch_split(ammsb_ctx* c, const PhiArgs& a) {
  constexpr int WPS = 8, R = 4;
  const size_t smem = (size_t)(2 + WPS * R) * a.K * 4 + (size_t)(1 + WPS * R) * 8 + ZIG_WORDS * 4;
  const bool exact = (a.K == 32u * KPL);
  auto kern = exact ? k_update_phi_split<KPL, WPS, R, true> : k_update_phi_split<KPL, WPS, R, false>;
  int occ = 0;
  if (resident_ctas_per_sm(kern, c->device, (WPS + 1) * 32, smem, &occ)) return 1;
  AMMSB_REQUIRE(occ > 0, "update_phi: split kernel does not fit on an SM");
  const uint32_t blocks = my_units(a);
  if (blocks == 0) return 0;
  kern<<<blocks, (WPS + 1) * 32, smem, c->stream>>>(a);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

template <int KPL, int STAGES, int T>
static int launch_team(ammsb_ctx* c, const PhiArgs& a) {
  const uint32_t KS = a.K / T;
  const uint32_t teams_per_cta = 4 / T;
  const size_t smem = (size_t)4 * (STAGES + 1) * KS * 4 + (size_t)teams_per_cta * 2 * a.K * 4 +
                      (size_t)4 * (STAGES + 1) * 8 + (size_t)teams_per_cta * 4 * 8 + (size_t)teams_per_cta * 3 * T * 4;
  const bool exact = (KS == 32u * KPL);
  auto kern = exact ? k_update_phi_team<KPL, STAGES, T, true> : k_update_phi_team<KPL, STAGES, T, false>;
  int occ = 0;
  if (resident_ctas_per_sm(kern, c->device, 192, smem, &occ)) return 1;
  AMMSB_REQUIRE(occ > 0, "update_phi: team kernel does not fit on an SM");
  const uint32_t active = my_units(a);
  if (active == 0) return 0;
  uint32_t blocks = (active + teams_per_cta - 1) / teams_per_cta;
  const uint32_t resident = (uint32_t)occ * c->sm_count;
  if (blocks > resident) blocks = resident;
  kern<<<blocks, 192, smem, c->stream>>>(a);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

// work-items (THREAD) or work-groups (WG) of the reference launch, phi.cc:740-747
static uint32_t phi_units(uint32_t mode, uint32_t wg, uint32_t V) {
  if (mode == AMMSB_MODE_THREAD) {
    uint32_t g = V / wg + (V % wg ? 1 : 0);
    if (g > 65535u) g = 65535u;
    return g * wg;
  }
  return V < 65535u ? V : 65535u;
}

extern "C" int ammsb_update_phi(ammsb_ctx* c, const ammsb_params* p, const ammsb_phi_opts* o,
                                const float* d_beta, ammsb_store* store, ammsb_set* train,
                                const uint32_t* d_nodes, const uint32_t* d_neighbors, uint32_t V,
                                uint32_t step_count, ammsb_rng* pool, float* d_phi_vec,
                                float* d_phi_sum) {
  AMMSB_REQUIRE(V > 0, "mini-batch nodes size = 0!");  // phi.cc:732
  AMMSB_REQUIRE(p->K == store->K && p->N == store->N, "params do not match the store");
  AMMSB_REQUIRE(o->wg > 0, "work-group size must be > 0");
  AMMSB_REQUIRE(d_phi_sum != nullptr && d_phi_vec != nullptr, "phi_vec / phi_sum are required");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  PhiArgs a;
  a.sv = store->view();
  a.set = train->view();
  a.beta = d_beta;
  a.nodes = d_nodes;
  a.neighbors = d_neighbors;
  a.V = V;
  a.n = p->num_neighbors;
  a.K = p->K;
  a.mode = o->mode;
  a.wg = o->wg;
  a.disable_noise = o->disable_noise;
  a.part_count = o->part_count ? o->part_count : 1;
  a.part_index = o->part_index;
  AMMSB_REQUIRE(a.part_index < a.part_count, "part_index out of range");
  // phi.cc:740-747 launch geometry
  uint64_t states;
  a.units = phi_units(o->mode, o->wg, V);
  if (o->mode == AMMSB_MODE_THREAD) {
    states = a.units < V ? a.units : V;
  } else {
    states = (uint64_t)a.units * o->wg;
  }
  AMMSB_REQUIRE(o->disable_noise || (pool && pool->n >= states),
                "Num seeds smaller than global threads");  // phi.cc:748-750
  a.eps_t = ammsb_eps_t(p, step_count);
  a.alpha = p->alpha;
  a.epsilon = p->epsilon;
  a.Nn = (1.0f * p->N) / p->num_neighbors;  // phi.cc:113
  a.pool = pool ? pool->d_state : nullptr;
  a.phi_vec = d_phi_vec;
  a.phi_sum = d_phi_sum;

  const bool fast_ok = !o->strict && (p->K % 4 == 0) && p->K <= 1024;
  if (fast_ok) {
    const uint32_t kpl = (p->K + 31) / 32;
    // few slots (link mini-batches: V = 1 + deg(u)): a CTA per slot, neighbors split over its
    // warps.  Chosen by the slot count of the whole mini-batch, so that every rank of a
    // multi-GPU run takes the same path (the two paths differ in fp32 rounding).
    if (V <= a.units && V <= (uint32_t)c->sm_count && a.n <= 256 && !getenv("AMMSB_PHI_NOSPLIT")) {
      if (kpl <= 2) return launch_split<2>(c, a);
      if (kpl <= 4) return launch_split<4>(c, a);
      if (kpl <= 8) return launch_split<8>(c, a);
      if (kpl <= 16) return launch_split<16>(c, a);
      return launch_split<32>(c, a);
    }
    // K <= 512: no producer warps (they cost the occupancy short rows need); the gather warp
    // draws the slot's noise while its first rows are in flight (AMMSB_PHI_LATE_NOISE: after the
    // neighbor loop, the earlier behaviour -- kept for A/B measurements)
    const bool late = getenv("AMMSB_PHI_LATE_NOISE") != nullptr;
    if (kpl <= 2) return late ? launch_fast<2, 8, 4, 2>(c, a) : launch_fast<2, 8, 4, 2, 0, true>(c, a);
    if (kpl <= 4) return late ? launch_fast<4, 8, 4, 2>(c, a) : launch_fast<4, 8, 4, 2, 0, true>(c, a);
    if (kpl <= 8) return late ? launch_fast<8, 6, 4, 2>(c, a) : launch_fast<8, 6, 4, 2, 0, true>(c, a);
    if (kpl <= 16) {
      // NB = 2 costs occupancy here (measured -6%); producer warps do not pay either at K = 512
      // (8 gather warps/SM cannot cover 2 KB rows: 5.28 vs 5.32 TB/s)
      return late ? launch_fast<16, 4, 4>(c, a) : launch_fast<16, 4, 4, 1, 0, true>(c, a);
    }
    // 4 gather warps x 4 stages + 2 noise-producer warps, 2 CTAs/SM: 6.10 TB/s at K = 1024
    // (the all-in-one <32,3,4> kernel: 5.66 TB/s; without noise both reach 6.3 TB/s)
    if (getenv("AMMSB_PHI_NOWS")) return launch_fast<32, 3, 4>(c, a);
    // experiment: no producers, early noise as for K <= 512 (8 gather warps/SM, 4 stages each)
    if (getenv("AMMSB_PHI_EARLY_1024")) return launch_fast<32, 4, 4, 1, 0, true>(c, a);
    // few slots (link mini-batches: V = 1 + deg(u)): the launch is a latency chain of row
    // round trips, not bandwidth -- one gather warp per CTA with 16 rows in flight
    if (my_units(a) <= (uint32_t)c->sm_count && !getenv("AMMSB_PHI_NOSMALL")) return launch_fast<32, 16, 1, 1, 1>(c, a);
    return launch_fast<32, 4, 4, 1, 2>(c, a);
  }
  // K in (1024, 4096]: teams of 2 or 4 warps per slot (one slot per unit, i.e. V <= 65535)
  if (!o->strict && V <= a.units && p->K > 1024 && p->K <= 4096) {
    // measured on B200: K=2048 5.82 TB/s, K=4096 5.55 TB/s (4 stages; 3 stages: 5.76 / 5.48)
    if (p->K <= 2048 && p->K % 8 == 0) return launch_team<32, 4, 2>(c, a);
    if (p->K % 16 == 0) return launch_team<32, 4, 4>(c, a);
  }
  const uint32_t vw = o->mode == AMMSB_MODE_THREAD ? 1u : o->wg;
  const size_t smem = sizeof(float) * (3 * (size_t)p->K + vw);
  AMMSB_REQUIRE(smem <= c->smem_optin, "K too large for the strict update_phi kernel");
  AMMSB_CHECK_CUDA(cudaFuncSetAttribute(k_update_phi_strict,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  uint32_t blocks = my_units(a);
  if (blocks == 0) return 0;
  const uint32_t cap = (uint32_t)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_update_phi_strict<<<blocks, 128, smem, c->stream>>>(a);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

// --------------------------------------------------------------- update_pi ---

// update_pi, phi.cc:154-173 / :178-197: pi[node][:] = phi_vec[slot][:] / sum,
// phi[node] = sum.  The row sum was produced by update_phi in the association of
// the launch mode; the division is IEEE.
__global__ void __launch_bounds__(256)
    k_update_pi(const __grid_constant__ StoreView sv, const float* __restrict__ phi_vec, const float* __restrict__ phi_sum,
                const uint32_t* __restrict__ nodes, uint32_t V, uint32_t units, uint32_t part_index,
                uint32_t part_count) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t K = sv.K;
  for (; warp < V; warp += nwarps) {
    if (part_count > 1 && (warp % units) % part_count != part_index) continue;  // another rank's slot
    const uint32_t node = __ldg(&nodes[warp]);
    const float* src = phi_vec + (size_t)warp * K;
    float* dst = store_row(sv, node);
    float sum;
    if (phi_sum) {
      sum = __ldg(&phi_sum[warp]);
    } else {
      float ls = 0.f;
      for (uint32_t k = lane; k < K; k += 32) ls = __fadd_rn(ls, src[k]);
      sum = warp_sum(ls);
    }
    if ((K & 3) == 0) {
      // four 128-bit loads in flight per lane before the (IEEE) divides and the stores
      for (uint32_t k0 = lane * 4; k0 < K; k0 += 512) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t k = k0 + 128 * u;
          if (k < K) v[u] = *reinterpret_cast<const float4*>(src + k);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t k = k0 + 128 * u;
          if (k < K) {
            v[u].x = __fdiv_rn(v[u].x, sum);
            v[u].y = __fdiv_rn(v[u].y, sum);
            v[u].z = __fdiv_rn(v[u].z, sum);
            v[u].w = __fdiv_rn(v[u].w, sum);
            *reinterpret_cast<float4*>(dst + k) = v[u];
            for (uint32_t r = 0; r < sv.num_mirrors; ++r)  // replicated mode: NVLink peer stores
              *reinterpret_cast<float4*>(sv.mirror_pi[r] + (size_t)node * K + k) = v[u];
          }
        }
      }
    } else {
      for (uint32_t k = lane; k < K; k += 32) {
        const float v = __fdiv_rn(src[k], sum);
        dst[k] = v;
        for (uint32_t r = 0; r < sv.num_mirrors; ++r) sv.mirror_pi[r][(size_t)node * K + k] = v;
      }
    }
    if (lane == 0)
      for (uint32_t r = 0; r < sv.num_mirrors; ++r) sv.mirror_phi[r][node] = sum;
    if (lane == 0) *store_phi(sv, node) = sum;
  }
}

static int update_pi_impl(ammsb_ctx* c, uint32_t K, ammsb_store* store, const float* d_phi_vec,
                          const float* d_phi_sum, const uint32_t* d_nodes, uint32_t V, uint32_t units,
                          uint32_t part_index, uint32_t part_count) {
  AMMSB_REQUIRE(V > 0, "mini-batch nodes size = 0!");
  AMMSB_REQUIRE(K == store->K, "K does not match the store");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  uint32_t blocks = (V + 7) / 8;
  const uint32_t cap = (uint32_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_update_pi<<<blocks, 256, 0, c->stream>>>(store->view(), d_phi_vec, d_phi_sum, d_nodes, V, units, part_index,
                                             part_count);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_update_pi(ammsb_ctx* c, uint32_t K, ammsb_store* store, const float* d_phi_vec,
                               const float* d_phi_sum, const uint32_t* d_nodes, uint32_t V) {
  return update_pi_impl(c, K, store, d_phi_vec, d_phi_sum, d_nodes, V, V ? V : 1, 0, 1);
}

extern "C" int ammsb_update_pi_part(ammsb_ctx* c, uint32_t K, ammsb_store* store, const float* d_phi_vec,
                                    const float* d_phi_sum, const uint32_t* d_nodes, uint32_t V,
                                    const ammsb_phi_opts* o) {
  AMMSB_REQUIRE(o->wg > 0, "work-group size must be > 0");
  const uint32_t pc = o->part_count ? o->part_count : 1;
  AMMSB_REQUIRE(o->part_index < pc, "part_index out of range");
  return update_pi_impl(c, K, store, d_phi_vec, d_phi_sum, d_nodes, V, phi_units(o->mode, o->wg, V ? V : 1),
                        o->part_index, pc);
}
