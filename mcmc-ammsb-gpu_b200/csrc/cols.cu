// cols.cu -- the COLUMN-SHARDED ("lane-sharded") multi-GPU layout of pi and the kernels that
// work on it.  The reference is single-device (its only scale mechanism is RowPartitionedMatrix,
// partitioned-alloc.h:14-141); this file is the multi-GPU successor of phi.cc / beta.cc /
// perplexity.cc in which WORK PARTIALS cross NVLink instead of pi rows.
//
// Layout.  The reference's default update_phi launch gives a mini-batch slot to a work-group of 32
// work-items; item l owns the columns k = l, l+32, ... and one RNG state (phi.cc:214-302,
// 740-747).  With G GPUs (2, 4 or 8) GPU g owns the reference lanes l = g (mod G) of EVERY row:
// LPG = 32/G lanes, K/G columns, held as a local matrix [N][K/G].  Every row access of every
// kernel is therefore local HBM -- no pi row ever crosses NVLink -- and the Langevin noise of a
// column is drawn on the GPU that owns it, from the same state as in the reference launch.
//
// What crosses NVLink are the K-wide sums: per (slot, neighbor) the partial sum_k probs_k, per slot
// the partial row sum of the new phi, per mini-batch edge two partial sums, per held-out pair two.
// The reference's WG_SUM (sum.cc:20-42) adds the 32 lane partials in a tree of strides 16, 8, 4,
// 2, 1.  Lanes l and l' with l = l' (mod G) are combined by the strides >= G, all inside one GPU
// (a sub-warp shuffle tree); the remaining strides G/2 .. 1 combine the G per-GPU partials.  Every
// GPU receives every other GPU's partial and adds them in that same tree order, so the full sum is
// bit-identical on every GPU AND bit-identical to what the one-warp-per-slot kernel of phi.cu
// computes on one GPU: the result does not depend on G.
//
// The exchange is fused into the kernels (no collective call, no barrier kernel): a partial is a
// single 4-byte store into a mailbox slot of the peer (mapped peer memory, NVLink), and a mailbox
// word is self-validating -- it holds the sentinel 0xffffffff (a NaN pattern no partial can have)
// until the value arrives, and the consumer re-arms it after reading.  No flags, no fences, no
// ordering requirement between words.  Mailbox halves alternate with the step parity, so a word is
// re-armed two steps before it is written again.
//
// k_cols_phi keeps the pi row pieces of a (slot, neighbor) in shared memory across the exchange:
// a warp runs G slots at once (LPG lanes each, the same per-lane arithmetic as phi.cu's lane),
// stages one neighbor piece per slot and trip with TMA bulk copies into a ring of R stages, forms
// the partial sums of stage t (phase A), and D trips later -- when the peers' partials have crossed
// the switch -- finishes stage t - D (phase B: gradient accumulation) from the same staged piece.
// HBM traffic is the algorithmic traffic; NVLink carries 4 bytes per (slot, neighbor, peer).
//
// Emulation: a launch may compute several ranks ("virtual ranks") of one GPU -- one cooperative
// launch over all ranks' shards and mailboxes -- which is how the one-GPU tests pin the G-rank
// protocol, G-invariance and parity with phi.cu bit for bit.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

#define COLS_SENTINEL 0xffffffffu
#define COLS_SPIN_LIMIT (1u << 22)
#define COLS_HDR_BYTES 256u

struct ColsRankView {
  float* pi;        // [N][KG]   this rank's columns of every row
  float* phi;       // [N]       row sums (every rank holds the same values)
  float* phi_vec;   // [Vcap][KG] update_phi output, local column order
  float* ppx;       // [Hcap]    running mean per held-out pair (every rank the same values)
  unsigned char* box[AMMSB_MAX_SHARDS];  // every rank's mailbox as mapped on this rank's device
  ulonglong2* pool;  // RNG pool of the current operator (phi: V*32 states, beta: K states)
  float* ws;         // beta partial sums [ctas][2][KG]
  double* ws_d;      // perplexity partial sums [ctas][4]
  uint32_t rank;
};

// offsets (bytes) inside a mailbox; every region is [2 halves][G sources][cap words]
struct ColsBoxLayout {
  size_t theta, beta;          // [2K] floats each (interleaved like the reference's buffers)
  size_t S, R, B, P;           // region bases
  size_t S_src, R_src, B_src, P_src;   // bytes per source
  size_t NB, NB_third;         // sampled neighbor lists [3][Vcap * n] u32 (step % 3), written by the rank that owns the sampler state
  size_t bytes;
};

struct ammsb_cols {
  ammsb_ctx* ctx = nullptr;
  uint64_t N = 0;
  uint32_t K = 0, G = 1, rank = 0, n = 0;
  uint32_t Vcap = 0, Ecap = 0;
  uint64_t Hcap = 0;
  uint32_t KG = 0;
  float *d_pi = nullptr, *d_phi = nullptr, *d_phi_vec = nullptr, *d_ppx = nullptr, *d_ws = nullptr;
  float* d_nz = nullptr;  // Langevin noise rows of the groups in flight: [resident warps][2][G][KG]
  size_t nz_warps = 0;
  double* d_ws_d = nullptr;
  ColsBoxLayout lay;
  VmmAlloc local, remote[AMMSB_MAX_SHARDS];
  unsigned char* box[AMMSB_MAX_SHARDS] = {nullptr};
  uint32_t ws_ctas = 0;
  uint32_t ppx_ctas = 0;  // grid of the last perplexity launch: its four sums sit behind that many CTA partials
  ColsRankView view(ulonglong2* pool) const {
    ColsRankView v;
    v.pi = d_pi; v.phi = d_phi; v.phi_vec = d_phi_vec; v.ppx = d_ppx;
    for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) v.box[i] = box[i];
    v.pool = pool; v.ws = d_ws; v.ws_d = d_ws_d; v.rank = rank;
    return v;
  }
};

// ---- geometry shared by host and device ----
// group-pass id of a slot: the warp sub-group that owns unit u = slot % units handles its slots
// u, u + units, ... (passes); G consecutive units form a group.
__host__ __device__ __forceinline__ uint32_t cols_groups(uint32_t V, uint32_t units, uint32_t G) {
  const uint32_t active = units < V ? units : V;
  return (active + G - 1) / G;
}
__host__ __device__ __forceinline__ uint32_t cols_passes(uint32_t V, uint32_t units) {
  return units ? (V + units - 1) / units : 0;
}
static uint32_t cols_units(uint32_t V) { return V < 65535u ? V : 65535u; }  // phi.cc:740-747, WG modes

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static ColsBoxLayout cols_layout(uint32_t K, uint32_t G, uint32_t n, uint32_t Vcap, uint32_t Ecap, uint64_t Hcap) {
  ColsBoxLayout l;
  size_t off = COLS_HDR_BYTES;
  l.theta = off; off += align_up(sizeof(float) * 2 * K, 256);
  l.beta = off; off += align_up(sizeof(float) * 2 * K, 256);
  const uint32_t units = cols_units(Vcap);
  const size_t gp = (size_t)cols_groups(Vcap, units, G) * cols_passes(Vcap, units);
  l.S_src = align_up(gp * n * G * 4, 256);
  l.R_src = align_up(gp * G * 4, 256);
  l.B_src = align_up(((size_t)Ecap + G) * 2 * 4, 256);
  l.P_src = align_up((Hcap + G) * 2 * 4, 256);
  l.S = off; off += 2 * (size_t)G * l.S_src;
  l.R = off; off += 2 * (size_t)G * l.R_src;
  l.B = off; off += 2 * (size_t)G * l.B_src;
  l.P = off; off += 2 * (size_t)G * l.P_src;
  l.NB_third = align_up((size_t)Vcap * n * 4, 256);
  l.NB = off; off += 3 * l.NB_third;
  l.bytes = off;
  return l;
}

// local float index of global column k on its owner (see the header comment):
//   l = k % 32, i = k / 32, rank = l % G, li = l / G, c = ((i / 4) * LPG + li) * 4 + i % 4
__host__ __device__ __forceinline__ uint32_t cols_local_index(uint32_t k, uint32_t G) {
  const uint32_t l = k & 31, i = k >> 5, li = l / G, LPG = 32 / G;
  return ((i >> 2) * LPG + li) * 4 + (i & 3);
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t ld_mbox(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_mbox(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t partial_bits(float v) {
  const uint32_t b = __float_as_uint(v);
  return b == COLS_SENTINEL ? 0x7fc00000u : b;  // a NaN stays a NaN, never the sentinel
}
// wait for a mailbox word, re-arm it; *err is set when the wait gives up
__device__ __forceinline__ float poll_mbox(uint32_t* p, uint32_t* err) {
  uint32_t v = ld_mbox(p), spins = 0;
  while (v == COLS_SENTINEL) {
    if (++spins > COLS_SPIN_LIMIT) {
      atomicExch(err, 1u);
      break;
    }
    v = ld_mbox(p);
  }
  st_mbox(p, COLS_SENTINEL);
  return __uint_as_float(v);
}

// 16-byte asynchronous copy global -> shared (L1 bypassed) and its completion on an mbarrier
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one attempt at a mailbox word (sentinel = not there yet)
__device__ __forceinline__ uint32_t peek_mbox(const uint32_t* p) { return ld_mbox(p); }
// finish a poll whose first attempt returned `v`; re-arms the word
__device__ __forceinline__ float finish_poll(uint32_t* p, uint32_t v, uint32_t* err) {
  uint32_t spins = 0;
  while (v == COLS_SENTINEL) {
    if (++spins > COLS_SPIN_LIMIT) {
      atomicExch(err, 1u);
      break;
    }
    v = ld_mbox(p);
  }
  st_mbox(p, COLS_SENTINEL);
  return __uint_as_float(v);
}

// Combine the G per-rank partials in the reference's tree order (strides G/2 .. 1 of WG_SUM).
// On entry lane `li` of every LPG-lane sub-group holds, for G > LPG, the partials of ranks li
// and li + LPG (v0, v1); for G <= LPG the lanes li < G hold the partial of rank li in v0.  On exit
// every lane of the sub-group holds the full sum.
template <int G>
__device__ __forceinline__ float cols_rank_tree(float v0, float v1, uint32_t lane) {
  constexpr int LPG = 32 / G;
  float v;
  if (G > LPG) {
    v = v0 + v1;  // stride LPG = G/2 (G = 8)
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  } else {
    const uint32_t li = lane & (LPG - 1);
    v = __shfl_sync(FULL_MASK, v0, (lane & ~(uint32_t)(LPG - 1)) | (li & (G - 1)));
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  }
  return v;
}

struct ColsPhiArgs {
  ColsRankView r[AMMSB_MAX_SHARDS];
  SetView set;
  ColsBoxLayout lay;
  const uint32_t* nodes;
  const uint32_t* neighbors;
  uint32_t nv, ctas_per_rank;
  uint32_t V, n, units;
  uint32_t R, D, MB;  // ring depth, A -> B distance (trips), metadata buffers
  uint32_t parity, disable_noise, loopback, split;
  uint32_t nb_poll, nb_third;  // the neighbor lists are the mailbox region NB (third step % 3): words arrive from the rank that sampled them
  uint32_t fake_lat;  // loopback only: cycles a warp waits after its sends, as if the peers' partials took that long (AMMSB_COLS_FAKE_LAT, ns)
  uint32_t debug;  // timing ablations (AMMSB_COLS_DEBUG): 1 no cuckoo, 2 no phase A math, 4 no phase B math, 8 no loads, 16 no Langevin step
  float eps_t, alpha, epsilon, Nn;
};

// bytes of shared memory per warp / per CTA (host and device agree through these)
template <int KPL, int G>
struct ColsPhiSmem {
  static constexpr int LPG = 32 / G;
  static constexpr int KG = KPL * LPG;
  static constexpr int PIECE = KG * 4;
  static constexpr int PSTRIDE = PIECE + ((LPG == 4 && (PIECE % 128) != 64) ? 64 : 0);
  static constexpr int STAGE = G * PSTRIDE;
  static constexpr int OWN = G * PIECE;
  static constexpr int META = G * 32 * 4 + 4 * G * 4;  // nb[32][G], ymask[G], slot[G], node[G], phi_sum[G]
  __host__ __device__ static size_t per_warp(uint32_t R, uint32_t MB) {
    return (size_t)R * STAGE + OWN + (size_t)MB * META + (size_t)R * G * 4 + ((size_t)R + 1) * 8 + 64;
  }
};

// One cursor over this warp's stages.  The three cursors of k_cols_phi (LOAD, phase A, phase B)
// walk the same sequence -- the warp's live group-passes in (group, pass) order, n stages each --
// at fixed distances, NB stages (a "trip") at a time; everything a trip needs is kept
// incrementally (no divisions).
struct ColsCursor {
  uint32_t group, pass;  // current group-pass; group >= ngroups: exhausted
  uint32_t j;            // first neighbor index of the trip within the group-pass
  uint32_t buf;          // ring buffer of the trip's first stage
  uint32_t meta;         // metadata buffer of the current segment (segment number mod MB)
};

__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async16_u32(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of this thread's most recent copy groups are still in flight
__device__ __forceinline__ void cp_async_wait_pending(uint32_t pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

// NB = stages per trip: the loop body handles NB consecutive neighbors in each of its three parts,
// which amortises the per-stage control over NB stages and gives the sum chains NB-fold ILP.
// Requires n % NB == 0, R % NB == 0, D % NB == 0.
template <int KPL, int G, int NB>
__global__ void __launch_bounds__(256, 1) k_cols_phi(const __grid_constant__ ColsPhiArgs a) {
  using SM = ColsPhiSmem<KPL, G>;
  constexpr int LPG = SM::LPG, KG = SM::KG, Q = KPL / 4;
  extern __shared__ __align__(128) unsigned char s_raw[];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t sub = lane / LPG, li = lane % LPG;
  const uint32_t vr = blockIdx.x / a.ctas_per_rank, cta = blockIdx.x % a.ctas_per_rank;
  const uint32_t rank = a.r[vr].rank;
  const uint32_t n = a.n, R = a.R, D = a.D, MB = a.MB;
  float* const my_pi = a.r[vr].pi;
  const float* const my_phi = a.r[vr].phi;
  float* const my_vec = a.r[vr].phi_vec;
  ulonglong2* const my_pool = a.r[vr].pool;

  // ---- shared memory: ziggurat tables | per warp: ring, own pieces, metadata, self partials, barriers
  uint32_t* s_zig = reinterpret_cast<uint32_t*>(s_raw);
  const ZigShared zig{s_zig};
  zig_stage(s_zig);
  const size_t pw = (SM::per_warp(R, MB) + 127) / 128 * 128;
  unsigned char* wbase = s_raw + 1536 + (size_t)wib * pw;
  unsigned char* s_ring = wbase;
  float* s_own = reinterpret_cast<float*>(s_ring + (size_t)R * SM::STAGE);
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_own) + SM::OWN);
  float* s_self = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_meta) + (size_t)MB * SM::META);
  __syncthreads();
  const uint32_t ring_u32 = smem_u32(s_ring), own_u32 = smem_u32(s_own);

  unsigned char* mybox = a.r[vr].box[rank];
  uint32_t* err = reinterpret_cast<uint32_t*>(mybox);
  const float* beta = reinterpret_cast<const float*>(mybox + a.lay.beta);
  const size_t half_S = (size_t)a.parity * G * a.lay.S_src, half_R = (size_t)a.parity * G * a.lay.R_src;
  // mailbox words this lane writes (its sub-group's partial to the ranks p = li, li + LPG) and reads
  uint32_t* send0 = nullptr;
  uint32_t* send1 = nullptr;
  uint32_t* poll0 = nullptr;
  uint32_t* poll1 = nullptr;
  if (li < G) {
    if (li != rank && !a.loopback) {
      send0 = reinterpret_cast<uint32_t*>(a.r[vr].box[li] + a.lay.S + half_S + (size_t)rank * a.lay.S_src);
      poll0 = reinterpret_cast<uint32_t*>(mybox + a.lay.S + half_S + (size_t)li * a.lay.S_src);
    }
  }
  if (G > LPG) {
    const uint32_t p1 = li + LPG;
    if (p1 != rank && !a.loopback) {
      send1 = reinterpret_cast<uint32_t*>(a.r[vr].box[p1] + a.lay.S + half_S + (size_t)rank * a.lay.S_src);
      poll1 = reinterpret_cast<uint32_t*>(mybox + a.lay.S + half_S + (size_t)p1 * a.lay.S_src);
    }
  }

  // f_k = beta_k - epsilon for the thread's columns k = l + 32 i, l = rank + G li  (phi.cc:237-239)
  const uint32_t l_ref = rank + G * li;
  float fb[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) fb[i] = beta[2 * (l_ref + 32 * i) + 1] - a.epsilon;
  const float e_link = a.epsilon, e_non = 1.0f - a.epsilon;
  const float half_eps = a.eps_t / 2;

  // ---- this warp's group-passes ----
  const uint32_t active_units = a.units < a.V ? a.units : a.V;
  const uint32_t ngroups = (active_units + G - 1) / G;
  const uint32_t passes = (a.V + a.units - 1) / a.units;
  const uint32_t gwarp = cta * warps + wib, total_warps = a.ctas_per_rank * warps;

  // a group-pass with no live slot is skipped by every cursor (slot of sub-group 0 is its smallest)
  auto gp_live = [&](uint32_t group, uint32_t pass) -> bool {
    return group < ngroups && (size_t)group * G + (size_t)pass * a.units < a.V;
  };
  auto next_gp = [&](uint32_t& group, uint32_t& pass) {
    do {
      if (++pass == passes) {
        pass = 0;
        group += total_warps;
      }
    } while (group < ngroups && !gp_live(group, pass));
  };
  auto step_cursor = [&](ColsCursor& c) {  // one trip on
    c.buf += NB;
    if (c.buf == R) c.buf = 0;
    c.j += NB;
    if (c.j == n) {
      c.j = 0;
      next_gp(c.group, c.pass);
      if (++c.meta == MB) c.meta = 0;
    } else if ((c.j & 31) == 0) {
      if (++c.meta == MB) c.meta = 0;
    }
  };
  auto slot_of = [&](uint32_t group, uint32_t pass, uint32_t s) -> uint32_t {
    const uint32_t unit = group * G + s;
    const uint32_t slot = unit + pass * a.units;
    return (unit < active_units && slot < a.V) ? slot : 0xffffffffu;
  };

  // metadata of a segment (<= 32 neighbors of the G slots of a group-pass): neighbor ids, the
  // cuckoo answers y (phi.cc:230-234), slot / node / phi_sum of every sub-group
  auto prep_segment = [&](const ColsCursor& c) {
    uint32_t* m = s_meta + (size_t)c.meta * (SM::META / 4);
    uint32_t* m_nb = m;
    uint32_t* m_y = m + G * 32;
    uint32_t* m_slot = m_y + G;
    uint32_t* m_node = m_slot + G;
    float* m_phi = reinterpret_cast<float*>(m_node + G);
    const uint32_t j0 = c.j, cnt = min(32u, n - j0);
    uint32_t my_node = 0;
    if (lane < G) {
      const uint32_t slot = slot_of(c.group, c.pass, lane);
      m_slot[lane] = slot;
      if (slot != 0xffffffffu) {
        my_node = __ldg(&a.nodes[slot]);
        m_node[lane] = my_node;
        m_phi[lane] = my_phi[my_node];
      } else {
        m_node[lane] = 0;
        m_phi[lane] = 1.0f;
      }
    }
#pragma unroll 4
    for (int s = 0; s < G; ++s) {
      const uint32_t slot = slot_of(c.group, c.pass, s);  // warp-uniform
      const uint32_t node = __shfl_sync(FULL_MASK, my_node, s);
      bool y = false;
      uint32_t nb = 0;
      if (slot != 0xffffffffu && lane < cnt) {
        nb = __ldg(&a.neighbors[(size_t)slot * n + j0 + lane]);
        if (!(a.debug & 1)) y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
      }
      m_nb[lane * G + s] = nb;
      const uint32_t mask = __ballot_sync(FULL_MASK, y);
      if (lane == 0) m_y[s] = mask;
    }
    __syncwarp();
  };

  ColsCursor cl, ca, cb;  // LOAD, phase A, phase B
  cl.group = gwarp;
  cl.pass = 0;
  cl.j = cl.buf = cl.meta = 0;
  if (cl.group < ngroups && !gp_live(cl.group, 0)) next_gp(cl.group, cl.pass);
  ca = cb = cl;

  float ownA[KPL], ownB[KPL], g[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) ownA[i] = ownB[i] = g[i] = 0.f;
  Rng st;
  st.x = st.y = 0;
  // per-segment / per-group-pass state of the A and B sides
  uint32_t ymaskA = 0, ymaskB = 0, slotB = 0xffffffffu;
  bool liveA = false, liveB = false;
  float phi_sumB = 1.0f, rphiB = 1.0f;
  size_t sidxA = 0, sidxB = 0;  // mailbox index of (group-pass, j = 0, sub)
  uint32_t pk0[NB], pk1[NB];    // early look at the next trip's mailbox words
#pragma unroll
  for (int u = 0; u < NB; ++u) pk0[u] = pk1[u] = COLS_SENTINEL;
  const uint32_t row_off = sub * SM::PSTRIDE + li * 16;  // this thread's first float4 within a stage

  const int trips_D = (int)(D / NB), trips_R = (int)(R / NB);
  for (int it = trips_D - trips_R;; ++it) {
    // ------------------------------------------------ phase B of the trip D stages back ----
    if (it >= trips_D) {
      if (cb.group >= ngroups) break;
      const uint32_t j = cb.j, b = cb.buf;
      if ((j & 31) == 0) {
        const uint32_t* m = s_meta + (size_t)cb.meta * (SM::META / 4);
        ymaskB = m[G * 32 + sub];
        if (j == 0) {  // the B side enters a group-pass: A is still inside it (D < n)
          slotB = m[G * 32 + G + sub];
          liveB = slotB != 0xffffffffu;
          phi_sumB = __uint_as_float(m[G * 32 + 3 * G + sub]);
          rphiB = 1.0f / phi_sumB;
          sidxB = ((size_t)cb.group * passes + cb.pass) * n * G + sub;
#pragma unroll
          for (int i = 0; i < KPL; ++i) {
            ownB[i] = ownA[i];
            g[i] = 0.f;
          }
          if (cb.pass == 0 && !a.disable_noise) {  // a new group: its units' RNG states (pass 0: slot == unit)
            const uint32_t unit = cb.group * G + sub;
            if (unit < active_units) st = rng_load(my_pool, (uint64_t)unit * 32 + l_ref);
          }
        }
      }
      // the G partials of every (slot, j + u): own from the ring, the peers' from the mailbox.  The
      // first look at this trip's words was taken one trip ago (pk*): the L2 round trip of a poll is
      // off the critical path.
      float S[NB];
      {
        float v0[NB], v1[NB];
        const size_t idx = sidxB + (size_t)j * G;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const float mine = s_self[(b + u) * G + sub];
          v0[u] = (li < G) ? mine : 0.f;
          v1[u] = (G > LPG) ? mine : 0.f;
        }
        if (liveB) {
          uint32_t w0[NB], w1[NB];
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            w0[u] = pk0[u];
            w1[u] = pk1[u];
          }
          if (j == 0) {
#pragma unroll
            for (int u = 0; u < NB; ++u) {
              if (poll0) w0[u] = peek_mbox(poll0 + idx + u * G);
              if (G > LPG && poll1) w1[u] = peek_mbox(poll1 + idx + u * G);
            }
          }
          if (j + NB < n) {
#pragma unroll
            for (int u = 0; u < NB; ++u) {
              if (poll0) pk0[u] = peek_mbox(poll0 + idx + (NB + u) * G);
              if (G > LPG && poll1) pk1[u] = peek_mbox(poll1 + idx + (NB + u) * G);
            }
          }
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            if (poll0) v0[u] = finish_poll(poll0 + idx + u * G, w0[u], err);
            if (G > LPG && poll1) v1[u] = finish_poll(poll1 + idx + u * G, w1[u], err);
          }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) S[u] = cols_rank_tree<G>(v0[u], v1[u], lane);
      }
      // (probs_k / probs_sum) / (pi_k * phi_sum) - 1 / phi_sum with probs_k / pi_k = t_k
      float inv[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) inv[u] = 1.0f / (S[u] * phi_sumB);
      const float nrphi = -rphiB;
      const uint32_t ybits = (ymaskB >> (j & 31)) & ((1u << NB) - 1u);
      const unsigned char* base = s_ring + (size_t)b * SM::STAGE + row_off;
      if (a.debug & 4) {
      } else if (!__any_sync(FULL_MASK, ybits != 0)) {  // no training link among the pairs (the usual case)
#pragma unroll
        for (int u = 0; u < NB; ++u) {  // neighbor order: the gradient is summed as on one GPU
          const float4* row = reinterpret_cast<const float4*>(base + (size_t)u * SM::STAGE);
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 r4 = row[q * LPG];
            g[4 * q] += fmaf(fmaf(r4.x, -fb[4 * q], e_non), inv[u], nrphi);
            g[4 * q + 1] += fmaf(fmaf(r4.y, -fb[4 * q + 1], e_non), inv[u], nrphi);
            g[4 * q + 2] += fmaf(fmaf(r4.z, -fb[4 * q + 2], e_non), inv[u], nrphi);
            g[4 * q + 3] += fmaf(fmaf(r4.w, -fb[4 * q + 3], e_non), inv[u], nrphi);
          }
        }
      } else {
#pragma unroll 1
        for (int u = 0; u < NB; ++u) {
          const bool y = (ybits >> u) & 1;
          const float e = y ? e_link : e_non;
          const uint32_t sgn = y ? 0u : 0x80000000u;
          const float4* row = reinterpret_cast<const float4*>(base + (size_t)u * SM::STAGE);
          const float iu = inv[u];
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 r4 = row[q * LPG];
            const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int i = 4 * q + c;
              g[i] += fmaf(fmaf(rr[c], __uint_as_float(__float_as_uint(fb[i]) ^ sgn), e), iu, nrphi);
            }
          }
        }
      }
      __syncwarp();  // every lane is done with the trip's stage buffers
      if (j + NB == n && !(a.debug & 16)) {
        // ---- Langevin step of the group-pass (phi.cc:266-274), noise in the state's draw order ----
        float* out = my_vec + (size_t)(liveB ? slotB : 0) * KG;
        float lsum = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          float o[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = 4 * q + c;
            const float noise = (a.disable_noise || !liveB) ? 1.0f : rng_randn_t(st, zig);
            o[c] = phi_langevin(ownB[i], phi_sumB, g[i], noise, half_eps, a.eps_t, a.alpha, a.Nn);
            lsum += o[c];
          }
          if (liveB) *reinterpret_cast<float4*>(out + (size_t)(q * LPG + li) * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) lsum += __shfl_xor_sync(FULL_MASK, lsum, o);
        // this rank's partial row sum to every rank (its own mailbox included): update_pi reads them
        if (liveB) {
          const size_t ridx = ((size_t)cb.group * passes + cb.pass) * G + sub;
          const uint32_t bits = partial_bits(lsum);
          for (uint32_t p = li; p < G; p += LPG)  // loopback: all G source regions of the own mailbox
            st_mbox(reinterpret_cast<uint32_t*>(a.r[vr].box[a.loopback ? rank : p] + a.lay.R + half_R +
                                                (size_t)(a.loopback ? p : rank) * a.lay.R_src) + ridx,
                    bits);
        }
        uint32_t ng = cb.group, np = cb.pass;
        next_gp(ng, np);
        if (ng != cb.group && !a.disable_noise) {  // the group is finished: persist its states
          const uint32_t unit = cb.group * G + sub;
          if (unit < active_units) rng_store(my_pool, (uint64_t)unit * 32 + l_ref, st);
        }
      }
      step_cursor(cb);
    }
    // ------------------------------------------------- LOAD of the trip R - D stages ahead of A ----
    if (cl.group < ngroups) {
      const uint32_t j = cl.j, b = cl.buf;
      if ((j & 31) == 0) prep_segment(cl);
      const uint32_t* m = s_meta + (size_t)cl.meta * (SM::META / 4);
      // Every lane copies 16-byte chunks (cp.async, L1 bypassed): chunk c*32 + lane of a stage's G
      // pieces.  An idle sub-group's id is 0 -- row 0 is copied and ignored -- so nothing here is
      // predicated.  Each lane's copies arrive on the stage barrier when they have landed.
      constexpr int CP = SM::PIECE / 16;      // chunks per piece
      constexpr int NC = (G * CP + 31) / 32;  // copies per lane and stage
      if (j == 0) {                           // the own pieces of the group-pass
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint32_t gc = c * 32 + lane, s = gc / CP, w = gc % CP;
          if (NC * 32 == G * CP || gc < G * CP)
            cp_async16_u32(own_u32 + s * SM::PIECE + w * 16, my_pi + (size_t)m[G * 32 + 2 * G + s] * KG + w * 4);
        }
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const uint32_t* ids = m + ((j & 31) + u) * G;
        const uint32_t dst = ring_u32 + (b + u) * SM::STAGE;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint32_t gc = c * 32 + lane, s = gc / CP, w = gc % CP;
          if ((NC * 32 == G * CP || gc < G * CP) && !(a.debug & 8))
            cp_async16_u32(dst + s * SM::PSTRIDE + w * 16, my_pi + (size_t)ids[s] * KG + w * 4);
        }
      }
      step_cursor(cl);
    }
    // one copy group per iteration, also when nothing was requested: phase A counts groups
    cp_async_commit();
    // ---------------------------------------------------------------------------- phase A ----
    if (it >= 0 && ca.group < ngroups) {
      const uint32_t j = ca.j, b = ca.buf;
      // the trip's pieces (and at j = 0 the own pieces) were requested R - D stages ago: all but
      // the (R - D) / NB youngest groups of every lane have landed
      cp_async_wait_pending((uint32_t)(trips_R - trips_D));
      __syncwarp();
      if ((j & 31) == 0) {
        const uint32_t* m = s_meta + (size_t)ca.meta * (SM::META / 4);
        ymaskA = m[G * 32 + sub];
        if (j == 0) {  // own pieces of the group-pass into registers
          liveA = m[G * 32 + G + sub] != 0xffffffffu;
          sidxA = ((size_t)ca.group * passes + ca.pass) * n * G + sub;
          const float4* o4 = reinterpret_cast<const float4*>(s_own + (size_t)sub * KG);
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 v = liveA ? o4[q * LPG + li] : make_float4(0.f, 0.f, 0.f, 0.f);
            ownA[4 * q] = v.x; ownA[4 * q + 1] = v.y; ownA[4 * q + 2] = v.z; ownA[4 * q + 3] = v.w;
          }
          __syncwarp();  // s_own may be refilled by the next group-pass's request
        }
      }
      const uint32_t ybits = (ymaskA >> (j & 31)) & ((1u << NB) - 1u);
      const unsigned char* base = s_ring + (size_t)b * SM::STAGE + row_off;
      float S[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) S[u] = 0.f;
      if (a.debug & 2) {
      } else if (!__any_sync(FULL_MASK, ybits != 0)) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
#pragma unroll
          for (int u = 0; u < NB; ++u) {  // NB independent chains, each in the order of one GPU's lane
            const float4 r4 = reinterpret_cast<const float4*>(base + (size_t)u * SM::STAGE)[q * LPG];
            S[u] = fmaf(ownA[4 * q], fmaf(r4.x, -fb[4 * q], e_non), S[u]);
            S[u] = fmaf(ownA[4 * q + 1], fmaf(r4.y, -fb[4 * q + 1], e_non), S[u]);
            S[u] = fmaf(ownA[4 * q + 2], fmaf(r4.z, -fb[4 * q + 2], e_non), S[u]);
            S[u] = fmaf(ownA[4 * q + 3], fmaf(r4.w, -fb[4 * q + 3], e_non), S[u]);
          }
        }
      } else {
#pragma unroll 1
        for (int u = 0; u < NB; ++u) {
          const bool y = (ybits >> u) & 1;
          const float e = y ? e_link : e_non;
          const uint32_t sgn = y ? 0u : 0x80000000u;
          const float4* row = reinterpret_cast<const float4*>(base + (size_t)u * SM::STAGE);
          float Su = 0.f;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float4 r4 = row[q * LPG];
            const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int i = 4 * q + c;
              Su = fmaf(ownA[i], fmaf(rr[c], __uint_as_float(__float_as_uint(fb[i]) ^ sgn), e), Su);
            }
          }
#pragma unroll
          for (int v = 0; v < NB; ++v)
            if (v == u) S[v] = Su;
        }
      }
      // strides 16 .. G of WG_SUM: the lanes of this GPU
#pragma unroll
      for (int o = LPG / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < NB; ++u) S[u] += __shfl_xor_sync(FULL_MASK, S[u], o);
      }
      if (li == 0) {
#pragma unroll
        for (int u = 0; u < NB; ++u) s_self[(b + u) * G + sub] = S[u];
      }
      if (liveA) {
        const size_t idx = sidxA + (size_t)j * G;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          const uint32_t bits = partial_bits(S[u]);
          if (send0) st_mbox(send0 + idx + u * G, bits);
          if (G > LPG && send1) st_mbox(send1 + idx + u * G, bits);
        }
      }
      __syncwarp();
      step_cursor(ca);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// k_cols_phi2 -- update_phi on the column shards, slot-at-a-time mapping.
//
// A warp owns a group of G consecutive units (the reference work-groups that own the slots and
// their RNG states) and takes the group's slots one after the other.  For a slot it stages the
// own piece and the pieces of (up to) 32 sampled neighbors in shared memory (cp.async, 16 bytes
// per lane, a piece per instruction: coalesced) and keeps them there across the exchange:
//   phase A   lane b owns NEIGHBOR b: it walks the K/G local columns with LPG independent
//             accumulators -- one per reference lane li, each summed over i in the order of that
//             lane on one GPU -- and combines them with the strides >= G of the WG_SUM tree, all
//             in registers: the 32 partial sums of the slot are ready without a single shuffle.
//   exchange  one coalesced 128-byte store per peer (the 32 partials of the slot), then each lane
//             collects its neighbor's G partials and finishes the tree (strides G/2 .. 1).
//   phase B   lane f owns COLUMNS (float4 f of the piece): the gradient is accumulated neighbor
//             by neighbor as on one GPU, then the Langevin step and the partial row sum.
// The Langevin noise of the group's G slots is drawn at the start of the group with lane = (slot,
// reference lane) -- every lane busy, states advanced in the reference's order -- into a scratch
// row in global memory (it stays in L2) that the Langevin step reads back.
// Register use is small (no per-column state lives across phases), so 12+ warps share an SM and
// hide each other's load and exchange latencies; the instruction stream has no per-neighbor
// control flow.
struct ColsPhi2Smem {
  __host__ __device__ static size_t pstr(uint32_t KG) { return (size_t)KG * 4 + 16; }  // piece stride: conflict-free float4 rows
  __host__ __device__ static size_t per_warp(uint32_t KG) { return (33 * pstr(KG) + 128 + (size_t)KG * 4 + 127) / 128 * 128; }
  __host__ __device__ static size_t per_cta(uint32_t KG) { return 1536 + (((size_t)KG * 4 + 127) / 128 * 128); }
  // warps of a CTA (one CTA per SM): what 227 KB of shared memory hold, at most 16; registers come in
  // units of four warps, so 13-15 warps would cost the 128-register budget of 16 for little: 12 (168)
  __host__ __device__ static constexpr uint32_t max_warps(uint32_t KG) {
    const size_t pw = (33 * ((size_t)KG * 4 + 16) + 128 + (size_t)KG * 4 + 127) / 128 * 128, pc = 1536 + (((size_t)KG * 4 + 127) / 128 * 128);
    const size_t fit = (232448 - pc) / pw;
    return fit >= 16 ? 16u : (fit > 12 ? 12u : (fit < 1 ? 1u : (uint32_t)fit));
  }
};

template <int KPL, int G>
__global__ void __launch_bounds__(ColsPhi2Smem::max_warps(KPL * (32 / G)) * 32, 1) k_cols_phi2(const __grid_constant__ ColsPhiArgs a, float* __restrict__ nz_scratch) {
  constexpr int LPG = 32 / G, KG = KPL * LPG, F4 = KG / 4;
  constexpr int PSTR = KG * 4 + 16;
  constexpr int FPL = (F4 + 31) / 32;  // float4 per lane in the column phases
  constexpr int PPI = F4 >= 32 ? 1 : 32 / F4;  // pieces per copy instruction
  constexpr int IPP = F4 >= 32 ? F4 / 32 : 1;  // copy instructions per piece
  extern __shared__ __align__(128) unsigned char s_raw[];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t vr = blockIdx.x / a.ctas_per_rank, cta = blockIdx.x % a.ctas_per_rank;
  const uint32_t rank = a.r[vr].rank;
  const uint32_t n = a.n;
  float* const my_pi = a.r[vr].pi;
  const float* const my_phi = a.r[vr].phi;
  float* const my_vec = a.r[vr].phi_vec;
  ulonglong2* const my_pool = a.r[vr].pool;

  uint32_t* s_zig = reinterpret_cast<uint32_t*>(s_raw);
  const ZigShared zig{s_zig};
  zig_stage(s_zig);
  float* s_fb = reinterpret_cast<float*>(s_raw + 1536);  // beta_k - epsilon, local column order
  unsigned char* mybox = a.r[vr].box[rank];
  uint32_t* err = reinterpret_cast<uint32_t*>(mybox);
  {
    const float* beta = reinterpret_cast<const float*>(mybox + a.lay.beta);
    for (uint32_t c = threadIdx.x; c < KG; c += blockDim.x) {
      const uint32_t f = c >> 2, e = c & 3, q = f / LPG, li = f % LPG;
      const uint32_t k = (rank + G * li) + 32 * (4 * q + e);
      s_fb[c] = beta[2 * k + 1] - a.epsilon;  // phi.cc:237-239
    }
  }
  __syncthreads();
  unsigned char* wbase = s_raw + ColsPhi2Smem::per_cta(KG) + (size_t)wib * ColsPhi2Smem::per_warp(KG);
  unsigned char* s_rows = wbase;                                   // [32] neighbor pieces
  unsigned char* s_own = wbase + 32 * PSTR;                        // own piece, later the new phi piece
  float* s_inv = reinterpret_cast<float*>(wbase + 33 * PSTR);      // [32] 1 / (probs_sum * phi_sum)
  float* s_new = reinterpret_cast<float*>(wbase + 33 * PSTR + 128);  // the new phi piece (row-sum scratch)
  const uint32_t rows_u32 = smem_u32(s_rows);

  const size_t half_S = (size_t)a.parity * G * a.lay.S_src, half_R = (size_t)a.parity * G * a.lay.R_src;
  const float e_link = a.epsilon, e_non = 1.0f - a.epsilon;
  const float half_eps = a.eps_t / 2;

  const uint32_t active_units = a.units < a.V ? a.units : a.V;
  const uint32_t ngroups = (active_units + G - 1) / G;
  const uint32_t passes = (a.V + a.units - 1) / a.units;
  const uint32_t gwarp = cta * warps + wib, total_warps = a.ctas_per_rank * warps;
  float* const my_nz = nz_scratch + ((size_t)vr * total_warps + gwarp) * 2 * G * KG;  // [2][G][KG]
  // a.split (few slots: the link mini-batches): the G slots of a group go to G different warps --
  // warp w takes sub-slot w % G of group w / G -- instead of one warp taking them one after the other
  const uint32_t g_first = a.split ? gwarp / G : gwarp, g_stride = a.split ? total_warps / G : total_warps;
  const uint32_t sub_lo = a.split ? gwarp % G : 0, sub_hi = a.split ? sub_lo + 1 : G;

  // noise lanes: lane = (slot s_n of the group, reference lane li_n)
  const uint32_t s_n = lane / LPG, li_n = lane % LPG, l_ref = rank + G * li_n;
  const bool nz_mine = !a.split || s_n == sub_lo;  // this lane's slot of the group is this warp's

  // The Langevin noise (phi.cc:266-274) of a group-pass is drawn one group-pass AHEAD of its use, a
  // float4 of every lane's stream at a time, inside the waits for the peers' partial sums: the
  // sequential draw chains fill time in which the warp would otherwise only poll.  `nz_*` is that
  // cursor; the noise of a group-pass alternates between the two halves of the scratch rows.
  constexpr int PIECES = KPL / 4;
  auto gp_live = [&](uint32_t group, uint32_t pass) -> bool {
    return group < ngroups && (size_t)group * G + (size_t)pass * a.units < a.V;
  };
  auto next_gp = [&](uint32_t& group, uint32_t& pass) {
    do {
      if (++pass == passes) {
        pass = 0;
        group += g_stride;
      }
    } while (group < ngroups && !gp_live(group, pass));
  };
  uint32_t nz_group = g_first, nz_pass = 0, nz_piece = 0, nz_half = 0;
  Rng st;
  st.x = st.y = 0;
  const bool noisy = !a.disable_noise && !(a.debug & 16);
  if (noisy && nz_mine && nz_group < ngroups && nz_group * G + s_n < active_units)
    st = rng_load(my_pool, (uint64_t)(nz_group * G + s_n) * 32 + l_ref);
  // one float4 (4 consecutive draws of every lane's stream) of the cursor's group-pass
  auto noise_piece = [&]() {
    if (!noisy || nz_group >= ngroups || nz_piece >= (uint32_t)PIECES) return;
    const uint32_t unit = nz_group * G + s_n, slot = unit + nz_pass * a.units;
    if (nz_mine && unit < active_units && slot < a.V) {
      float4 z;
      z.x = rng_randn_t(st, zig);
      z.y = rng_randn_t(st, zig);
      z.z = rng_randn_t(st, zig);
      z.w = rng_randn_t(st, zig);
      reinterpret_cast<float4*>(my_nz + ((size_t)nz_half * G + s_n) * KG)[nz_piece * LPG + li_n] = z;
    }
    ++nz_piece;
  };
  // the cursor moves on to the next group-pass (its current one must be complete)
  auto noise_advance = [&]() {
    if (nz_group >= ngroups) return;
    const uint32_t old = nz_group;
    next_gp(nz_group, nz_pass);
    nz_piece = 0;
    nz_half ^= 1;
    if (noisy && nz_group != old) {  // the old group's streams are finished: persist, load the new ones
      if (nz_mine && old * G + s_n < active_units) rng_store(my_pool, (uint64_t)(old * G + s_n) * 32 + l_ref, st);
      if (nz_mine && nz_group < ngroups && nz_group * G + s_n < active_units)
        st = rng_load(my_pool, (uint64_t)(nz_group * G + s_n) * 32 + l_ref);
    }
  };
  // prologue: the whole noise of the warp's first group-pass
  if (g_first < ngroups) {
    for (int k = 0; k < PIECES; ++k) noise_piece();
    __syncwarp();
    noise_advance();
  }

  // the pieces of a chunk: neighbor b -> row b, (first chunk) the own piece -> row 32; one commit group
  auto issue_copies = [&](uint32_t node, uint32_t nb, uint32_t cnt, bool own) {
    if (F4 >= 32) {
#pragma unroll 8
      for (uint32_t r = 0; r < 32; ++r) {
        const uint32_t id = __shfl_sync(FULL_MASK, nb, r);
        if (r < cnt) {
#pragma unroll
          for (int h = 0; h < IPP; ++h)
            cp_async16_u32(rows_u32 + r * PSTR + (h * 32 + lane) * 16, my_pi + (size_t)id * KG + (h * 32 + lane) * 4);
        }
      }
      if (own) {
#pragma unroll
        for (int h = 0; h < IPP; ++h)
          cp_async16_u32(rows_u32 + 32 * PSTR + (h * 32 + lane) * 16, my_pi + (size_t)node * KG + (h * 32 + lane) * 4);
      }
    } else {
      const uint32_t sub = lane / F4, w = lane % F4;
#pragma unroll 8
      for (uint32_t r0 = 0; r0 < 32; r0 += PPI) {
        const uint32_t r = r0 + sub;
        const uint32_t id = __shfl_sync(FULL_MASK, nb, r);
        if (r < cnt) cp_async16_u32(rows_u32 + r * PSTR + w * 16, my_pi + (size_t)id * KG + w * 4);
      }
      if (own && lane < F4) cp_async16_u32(rows_u32 + 32 * PSTR + lane * 16, my_pi + (size_t)node * KG + lane * 4);
    }
    cp_async_commit();
  };

  uint32_t group = g_first, pass = 0;
  for (; group < ngroups; next_gp(group, pass)) {
    {
      const uint32_t use_half = nz_half ^ 1;  // the half the cursor filled before it moved on
      {
      // the next slot's node, row sum, neighbor id and cuckoo answer are fetched while the current
      // slot waits for its partials and runs phase B (n <= 32: one chunk per slot)
      bool pf_valid = false, pf_y = false, copies_ahead = false;
      uint32_t pf_node = 0, pf_nb = 0;
      float pf_phi = 0.f;
      for (uint32_t s = sub_lo; s < sub_hi; ++s) {
        const uint32_t unit = group * G + s, slot = unit + pass * a.units;
        if (unit >= active_units || slot >= a.V) break;  // warp-uniform; later sub-slots are dead too
        const uint32_t node = pf_valid ? pf_node : __ldg(&a.nodes[slot]);
        const float phi_sum = pf_valid ? pf_phi : my_phi[node];
        const float rphi = 1.0f / phi_sum;
        const bool have_pf = pf_valid;
        const bool pf_next = n <= 32 && s + 1 < sub_hi && unit + 1 < active_units && slot + 1 < a.V && !a.debug;
        const size_t ridx = ((size_t)group * passes + pass) * G + s;
        float4 g4[FPL];
#pragma unroll
        for (int r = 0; r < FPL; ++r) g4[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t c0 = 0; c0 < n; c0 += 32) {
          const uint32_t cnt = min(32u, n - c0);
          // ---- stage the pieces: neighbor b of the chunk -> row b; the own piece -> row 32 ----
          uint32_t nb = node;
          if (have_pf) {
            nb = pf_nb;
          } else if (lane < cnt) {
            if (a.nb_poll) {  // sampled on the rank that owns the slot's sampler state, delivered by peer stores
              uint32_t* w = reinterpret_cast<uint32_t*>(mybox + a.lay.NB + (size_t)a.nb_third * a.lay.NB_third) + (size_t)slot * n + c0 + lane;
              nb = __float_as_uint(finish_poll(w, peek_mbox(w), err));
            } else {
              nb = __ldg(&a.neighbors[(size_t)slot * n + c0 + lane]);
            }
          }
          if (!(a.debug & 8) && !(have_pf && copies_ahead)) issue_copies(node, nb, cnt, c0 == 0);
          // the cuckoo answer for this lane's neighbor while the pieces travel (phi.cc:230-234)
          bool y = false;
          if (have_pf) y = pf_y;
          else if (lane < cnt && !(a.debug & 1)) y = set_has(a.set, make_edge(min(node, nb), max(node, nb)));
          const uint32_t ymask = __ballot_sync(FULL_MASK, y);
          cp_async_wait_pending(0);
          __syncwarp();

          // ---- phase A: lane b = neighbor b; LPG chains, one per reference lane of this GPU ----
          float S = 0.f;
          if (!(a.debug & 2)) {
            float ch[LPG];
#pragma unroll
            for (int t = 0; t < LPG; ++t) ch[t] = 0.f;
            const float4* own4 = reinterpret_cast<const float4*>(s_own);
            const float4* fb4 = reinterpret_cast<const float4*>(s_fb);
            const float4* row4 = reinterpret_cast<const float4*>(s_rows + (size_t)(lane < cnt ? lane : 0) * PSTR);
            if (ymask == 0) {  // no training link among the chunk's pairs (the usual case)
#pragma unroll
              for (int f = 0; f < F4; ++f) {
                const float4 o = own4[f], fb = fb4[f], r = row4[f];
                float& acc = ch[f % LPG];
                acc = fmaf(o.x, fmaf(-r.x, fb.x, e_non), acc);  // fma(r, -f, e) == fma(-r, f, e) bit for bit
                acc = fmaf(o.y, fmaf(-r.y, fb.y, e_non), acc);
                acc = fmaf(o.z, fmaf(-r.z, fb.z, e_non), acc);
                acc = fmaf(o.w, fmaf(-r.w, fb.w, e_non), acc);
              }
            } else {
              const float e = y ? e_link : e_non;
              const float sg = y ? 1.0f : -1.0f;
#pragma unroll
              for (int f = 0; f < F4; ++f) {
                const float4 o = own4[f], fb = fb4[f], r = row4[f];
                float& acc = ch[f % LPG];
                acc = fmaf(o.x, fmaf(r.x * sg, fb.x, e), acc);
                acc = fmaf(o.y, fmaf(r.y * sg, fb.y, e), acc);
                acc = fmaf(o.z, fmaf(r.z * sg, fb.z, e), acc);
                acc = fmaf(o.w, fmaf(r.w * sg, fb.w, e), acc);
              }
            }
            // strides 16 .. G of WG_SUM over the reference lanes l = rank + G li: li strides LPG/2 .. 1
#pragma unroll
            for (int o = LPG / 2; o > 0; o >>= 1) {
#pragma unroll
              for (int t = 0; t < o; ++t) ch[t] += ch[t + o];
            }
            S = ch[0];
          }
          // ---- exchange: the slot's partials to every peer; collect this neighbor's G partials ----
          const size_t sidx = ridx * n + c0 + lane;
          float P[G];
#pragma unroll
          for (int p = 0; p < G; ++p) P[p] = S;
          if (lane < cnt && !a.loopback) {
            const uint32_t bits = partial_bits(S);
#pragma unroll
            for (int p = 0; p < G; ++p)
              if ((uint32_t)p != rank)
                st_mbox(reinterpret_cast<uint32_t*>(a.r[vr].box[p] + a.lay.S + half_S + (size_t)rank * a.lay.S_src) + sidx, bits);
          }
          const long long t_sent = a.fake_lat ? clock64() : 0;
          uint32_t* pf_w = nullptr;
          if (pf_next) {  // first-level loads of the next slot, in flight across the noise piece and the wait
            pf_node = __ldg(&a.nodes[slot + 1]);
            pf_nb = pf_node;
            if (lane < n) {
              if (a.nb_poll) {
                pf_w = reinterpret_cast<uint32_t*>(mybox + a.lay.NB + (size_t)a.nb_third * a.lay.NB_third) + (size_t)(slot + 1) * n + lane;
                pf_nb = peek_mbox(pf_w);
              } else {
                pf_nb = __ldg(&a.neighbors[(size_t)(slot + 1) * n + lane]);
              }
            }
          }
          if (c0 == 0) noise_piece();  // while the partials cross the switch: part of the next group-pass's noise
          if (a.fake_lat) {
            while (clock64() - t_sent < (long long)a.fake_lat) {}
          }
          if (lane < cnt && !a.loopback) {
            uint32_t w[G];
#pragma unroll
            for (int p = 0; p < G; ++p)
              if ((uint32_t)p != rank)
                w[p] = peek_mbox(reinterpret_cast<uint32_t*>(mybox + a.lay.S + half_S + (size_t)p * a.lay.S_src) + sidx);
#pragma unroll
            for (int p = 0; p < G; ++p)
              if ((uint32_t)p != rank)
                P[p] = finish_poll(reinterpret_cast<uint32_t*>(mybox + a.lay.S + half_S + (size_t)p * a.lay.S_src) + sidx, w[p], err);
          }
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) {  // strides G/2 .. 1 of WG_SUM
#pragma unroll
            for (int p = 0; p < o; ++p) P[p] += P[p + o];
          }
          // (probs_k / probs_sum) / (pi_k * phi_sum) - 1 / phi_sum with probs_k / pi_k = t_k
          s_inv[lane] = 1.0f / (P[0] * phi_sum);
          __syncwarp();
          if (pf_next) {  // second level: the row sum and the cuckoo bins, in flight across phase B
            pf_phi = my_phi[pf_node];
            if (pf_w != nullptr) pf_nb = __float_as_uint(finish_poll(pf_w, pf_nb, err));
            pf_y = lane < n ? set_has(a.set, make_edge(min(pf_node, pf_nb), max(pf_node, pf_nb))) : false;
          }
          pf_valid = pf_next;

          // ---- phase B: lane = columns (float4 f = lane + 32 r); neighbors in order ----
          if (!(a.debug & 4)) {
            float4 fb[FPL];
#pragma unroll
            for (int r = 0; r < FPL; ++r)
              fb[r] = (lane + 32 * r < F4) ? reinterpret_cast<const float4*>(s_fb)[lane + 32 * r] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float nrphi = -rphi;
            if (ymask == 0) {  // no training link among the chunk's pairs (the usual case)
#pragma unroll 8
              for (uint32_t b = 0; b < cnt; ++b) {
                const float inv = s_inv[b];
#pragma unroll
                for (int r = 0; r < FPL; ++r) {
                  if (lane + 32 * r < F4) {
                    const float4 x = reinterpret_cast<const float4*>(s_rows + (size_t)b * PSTR)[lane + 32 * r];
                    g4[r].x += fmaf(fmaf(-x.x, fb[r].x, e_non), inv, nrphi);
                    g4[r].y += fmaf(fmaf(-x.y, fb[r].y, e_non), inv, nrphi);
                    g4[r].z += fmaf(fmaf(-x.z, fb[r].z, e_non), inv, nrphi);
                    g4[r].w += fmaf(fmaf(-x.w, fb[r].w, e_non), inv, nrphi);
                  }
                }
              }
            } else {
#pragma unroll 4
              for (uint32_t b = 0; b < cnt; ++b) {
                const float inv = s_inv[b];
                const bool yb = (ymask >> b) & 1;
                const float e = yb ? e_link : e_non, sg = yb ? 1.0f : -1.0f;
#pragma unroll
                for (int r = 0; r < FPL; ++r) {
                  if (lane + 32 * r < F4) {
                    const float4 x = reinterpret_cast<const float4*>(s_rows + (size_t)b * PSTR)[lane + 32 * r];
                    g4[r].x += fmaf(fmaf(x.x * sg, fb[r].x, e), inv, nrphi);
                    g4[r].y += fmaf(fmaf(x.y * sg, fb[r].y, e), inv, nrphi);
                    g4[r].z += fmaf(fmaf(x.z * sg, fb[r].z, e), inv, nrphi);
                    g4[r].w += fmaf(fmaf(x.w * sg, fb[r].w, e), inv, nrphi);
                  }
                }
              }
            }
          }
          __syncwarp();  // the rows may be overwritten by the next chunk / slot
        }
        // ---- Langevin step (phi.cc:266-274): lane = columns.  The own piece moves to registers and the
        //      staging buffer is handed to the NEXT slot's copies first (its node and neighbor ids were
        //      prefetched): they travel while this slot finishes ----
        if (!(a.debug & 16)) {
          const float* nzrow = my_nz + ((size_t)use_half * G + s) * KG;
          float4 o4[FPL];
#pragma unroll
          for (int r = 0; r < FPL; ++r)
            o4[r] = (lane + 32 * r < F4) ? reinterpret_cast<const float4*>(s_own)[lane + 32 * r] : make_float4(0.f, 0.f, 0.f, 0.f);
          __syncwarp();
          copies_ahead = pf_valid && !(a.debug & 8);
          if (copies_ahead) issue_copies(pf_node, pf_nb, n, true);
#pragma unroll
          for (int r = 0; r < FPL; ++r) {
            const uint32_t f = lane + 32 * r;
            if (f < F4) {
              const float4 o = o4[r];
              float4 z = make_float4(1.f, 1.f, 1.f, 1.f);
              if (!a.disable_noise) z = __ldcg(reinterpret_cast<const float4*>(nzrow) + f);
              float4 v;
              v.x = phi_langevin(o.x, phi_sum, g4[r].x, z.x, half_eps, a.eps_t, a.alpha, a.Nn);
              v.y = phi_langevin(o.y, phi_sum, g4[r].y, z.y, half_eps, a.eps_t, a.alpha, a.Nn);
              v.z = phi_langevin(o.z, phi_sum, g4[r].z, z.z, half_eps, a.eps_t, a.alpha, a.Nn);
              v.w = phi_langevin(o.w, phi_sum, g4[r].w, z.w, half_eps, a.eps_t, a.alpha, a.Nn);
              reinterpret_cast<float4*>(my_vec + (size_t)slot * KG)[f] = v;
              reinterpret_cast<float4*>(s_new)[f] = v;
            }
          }
          __syncwarp();
          // partial row sum: reference lane li adds its columns in order (i = 0 .. KPL-1), then the
          // strides >= G of the tree; lanes li < LPG hold one reference lane each
          float ls = 0.f;
          if (lane < LPG) {
            const float* v = s_new;
#pragma unroll 8
            for (int i = 0; i < KPL; ++i) ls += v[((i >> 2) * LPG + lane) * 4 + (i & 3)];
          }
#pragma unroll
          for (int o = LPG / 2; o > 0; o >>= 1) ls += __shfl_xor_sync(FULL_MASK, ls, o);
          ls = __shfl_sync(FULL_MASK, ls, 0);
          if (lane < (uint32_t)G) {  // this rank's partial to every rank (its own mailbox included)
            const uint32_t p = lane;
            st_mbox(reinterpret_cast<uint32_t*>(a.r[vr].box[a.loopback ? rank : p] + a.lay.R + half_R +
                                                (size_t)(a.loopback ? p : rank) * a.lay.R_src) + ridx,
                    partial_bits(ls));
          }
          __syncwarp();  // s_new is free for the next slot
        } else {
          copies_ahead = false;
        }
      }
      }
      // what is left of the next group-pass's noise, then the cursor moves on
      while (noisy && nz_group < ngroups && nz_piece < (uint32_t)PIECES) noise_piece();
      __syncwarp();
      noise_advance();
    }
  }
  // the last group's streams (the cursor has run past the end: nothing left to persist otherwise)
}


// ---- update_pi on the column shards (phi.cc:154-197): pi[node][own columns] = phi_vec / sum ----
struct ColsPiArgs {
  ColsRankView r[AMMSB_MAX_SHARDS];
  ColsBoxLayout lay;
  const uint32_t* nodes;
  uint32_t nv, ctas_per_rank, G, KG;
  uint32_t V, units, parity;
};

__global__ void __launch_bounds__(256) k_cols_pi(const __grid_constant__ ColsPiArgs a) {
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t vr = blockIdx.x / a.ctas_per_rank, cta = blockIdx.x % a.ctas_per_rank;
  const ColsRankView& me = a.r[vr];
  const uint32_t G = a.G, KG = a.KG;
  unsigned char* mybox = me.box[me.rank];
  uint32_t* err = reinterpret_cast<uint32_t*>(mybox);
  const uint32_t passes = (a.V + a.units - 1) / a.units;
  const size_t half_R = (size_t)a.parity * G * a.lay.R_src;
  for (uint32_t slot = cta * warps + wib; slot < a.V; slot += a.ctas_per_rank * warps) {
    const uint32_t pass = slot / a.units, unit = slot - pass * a.units;
    const size_t ridx = ((size_t)(unit / G) * passes + pass) * G + unit % G;
    float v = 0.f;
    if (lane < G) v = poll_mbox(reinterpret_cast<uint32_t*>(mybox + a.lay.R + half_R + (size_t)lane * a.lay.R_src) + ridx, err);
    for (uint32_t o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);  // strides G/2 .. 1 of WG_SUM
    const float sum = __shfl_sync(FULL_MASK, v, 0);
    const uint32_t node = __ldg(&a.nodes[slot]);
    const float4* src = reinterpret_cast<const float4*>(me.phi_vec + (size_t)slot * KG);
    float4* dst = reinterpret_cast<float4*>(me.pi + (size_t)node * KG);
    for (uint32_t f = lane; f < KG / 4; f += 32) {
      float4 x = src[f];
      x.x = __fdiv_rn(x.x, sum);
      x.y = __fdiv_rn(x.y, sum);
      x.z = __fdiv_rn(x.z, sum);
      x.w = __fdiv_rn(x.w, sum);
      dst[f] = x;
    }
    if (lane == 0) me.phi[node] = sum;
  }
}

// ---- update_beta on the column shards (beta.cc:30-137, 334-384) ----
// A sub-group of LPG lanes per mini-batch edge, G edges per warp trip.  The two K-wide sums of an
// edge (sum pi_u pi_v and sum probs) are exchanged as partials like in k_cols_phi; both row pieces
// stay in registers across the exchange.  Accumulators A_k = sum_{y=0} p_k / S, B_k = sum_{y=1} p_k / S
// per thread, combined across sub-groups, warps and CTAs in a fixed order.
struct ColsBetaArgs {
  ColsRankView r[AMMSB_MAX_SHARDS];
  SetView set;
  ColsBoxLayout lay;
  const uint64_t* edges;
  uint32_t nv, ctas_per_rank;
  uint32_t E_mb, parity, loopback;
  float epsilon;
};

// two mailbox words at once (a partial pair is 8-byte aligned); every word still validates itself
__device__ __forceinline__ uint2 ld_mbox2(const uint32_t* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_mbox2(uint32_t* p, uint32_t x, uint32_t y) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 finish_poll2(uint32_t* p, uint2 v, uint32_t* err) {
  uint32_t spins = 0;
  while (v.x == COLS_SENTINEL || v.y == COLS_SENTINEL) {
    if (++spins > COLS_SPIN_LIMIT) {
      atomicExch(err, 1u);
      break;
    }
    v = ld_mbox2(p);
  }
  st_mbox2(p, COLS_SENTINEL, COLS_SENTINEL);
  return v;
}

// The exchange is taken off the critical path: a trip's partial pair is SENT when its rows are
// first read, and the trip is FINISHED (peers' pairs collected, tree, accumulation) BETA_D trips
// later, from the same rows read again -- they are in L2 -- so a warp never sits through an NVLink
// round trip, and the ranks need not run in lockstep trip by trip.  (Measured on 8 GPUs with the
// finish right behind the send: 0.37 ms for 131072 edges against 0.06 ms without the waits.)
#define BETA_D 3
template <int KPL, int G>
__global__ void __launch_bounds__(128) k_cols_beta(const __grid_constant__ ColsBetaArgs a) {
  constexpr int LPG = 32 / G, KG = KPL * LPG, Q = KPL / 4, WARPS = 4;
  __shared__ float s_acc[2][KG];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t sub = lane / LPG, li = lane % LPG;
  const uint32_t vr = blockIdx.x / a.ctas_per_rank, cta = blockIdx.x % a.ctas_per_rank;
  const ColsRankView& me = a.r[vr];
  const uint32_t rank = me.rank;
  unsigned char* mybox = me.box[rank];
  uint32_t* err = reinterpret_cast<uint32_t*>(mybox);
  const float* beta = reinterpret_cast<const float*>(mybox + a.lay.beta);
  const size_t half_B = (size_t)a.parity * G * a.lay.B_src;
  const uint32_t l_ref = rank + G * li;
  float bk[KPL], accA[KPL], accB[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
    bk[i] = beta[2 * (l_ref + 32 * i) + 1];
    accA[i] = accB[i] = 0.f;
  }
  const uint32_t trips = (a.E_mb + G - 1) / G;
  const uint32_t first = cta * WARPS + wib, stride = a.ctas_per_rank * WARPS;
  const uint32_t my_trips = first < trips ? (trips - first + stride - 1) / stride : 0;
  uint32_t ybits = 0;  // y of the trips in flight, bit = trip number mod 32

  // rows of edge e -> q[i] = probs_k of the own columns, the two partial sums (sub-group reduced)
  auto partials = [&](uint32_t e, bool live, bool have_y, bool& y, float* q, float& pi_sum, float& probs_sum) {
    pi_sum = 0.f;
    probs_sum = 0.f;
    if (live) {
      const uint64_t edge = __ldg(&a.edges[e]);
      const uint32_t u = (uint32_t)(edge >> 32), v = (uint32_t)(edge & 0xffffffffu);
      const float4* pa = reinterpret_cast<const float4*>(me.pi + (size_t)u * KG);
      const float4* pb = reinterpret_cast<const float4*>(me.pi + (size_t)v * KG);
#pragma unroll
      for (int qq = 0; qq < Q; ++qq) {
        const float4 x = ldg_stream4(reinterpret_cast<const float*>(pa + qq * LPG + li));
        const float4 z = ldg_stream4(reinterpret_cast<const float*>(pb + qq * LPG + li));
        q[4 * qq] = x.x * z.x; q[4 * qq + 1] = x.y * z.y; q[4 * qq + 2] = x.z * z.z; q[4 * qq + 3] = x.w * z.w;
      }
      if (!have_y) y = set_has(a.set, make_edge(min(u, v), max(u, v)));
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        pi_sum += q[i];
        q[i] *= y ? bk[i] : 1.0f - bk[i];  // probs_k (beta.cc:116-120)
        probs_sum += q[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < KPL; ++i) q[i] = 0.f;
    }
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) {
      pi_sum += __shfl_xor_sync(FULL_MASK, pi_sum, o);
      probs_sum += __shfl_xor_sync(FULL_MASK, probs_sum, o);
    }
  };

  for (uint32_t it = 0; it < my_trips + BETA_D; ++it) {
    if (it < my_trips) {  // ---- send: partial pair of trip `it` to every peer ----
      const uint32_t e = (first + it * stride) * G + sub;
      const bool live = e < a.E_mb;
      float q[KPL];
      bool y = false;
      float pi_sum, probs_sum;
      partials(e, live, false, y, q, pi_sum, probs_sum);
      ybits = (ybits & ~(1u << (it & 31))) | ((uint32_t)y << (it & 31));
      if (live && !a.loopback) {
        const uint32_t pb_ = partial_bits(pi_sum), qb_ = partial_bits(probs_sum);
        for (uint32_t p = li; p < G; p += LPG)
          if (p != rank)
            st_mbox2(reinterpret_cast<uint32_t*>(me.box[p] + a.lay.B + half_B + (size_t)rank * a.lay.B_src) + (size_t)e * 2,
                     pb_, qb_);
      }
    }
    if (it >= BETA_D) {  // ---- finish trip it - BETA_D ----
      const uint32_t jt = it - BETA_D;
      const uint32_t e = (first + jt * stride) * G + sub;
      const bool live = e < a.E_mb;
      const size_t idx = (size_t)e * 2;
      // the peers' pairs first (they have had BETA_D trips to arrive), then the rows again
      const uint32_t p0 = li, p1 = li + LPG;
      const bool poll0 = live && !a.loopback && p0 < (uint32_t)G && p0 != rank;
      const bool poll1 = live && !a.loopback && G > LPG && p1 != rank;
      uint32_t* src0 = reinterpret_cast<uint32_t*>(mybox + a.lay.B + half_B + (size_t)p0 * a.lay.B_src) + idx;
      uint32_t* src1 = reinterpret_cast<uint32_t*>(mybox + a.lay.B + half_B + (size_t)(G > LPG ? p1 : 0) * a.lay.B_src) + idx;
      uint2 w0 = make_uint2(0u, 0u), w1 = make_uint2(0u, 0u);
      if (poll0) w0 = ld_mbox2(src0);
      if (poll1) w1 = ld_mbox2(src1);
      float q[KPL];
      bool y = (ybits >> (jt & 31)) & 1u;
      float pi_sum, probs_sum;
      partials(e, live, true, y, q, pi_sum, probs_sum);
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
      if (live) {
        if (li < (uint32_t)G) {
          if (poll0) {
            w0 = finish_poll2(src0, w0, err);
            a0 = __uint_as_float(w0.x); b0 = __uint_as_float(w0.y);
          } else {
            a0 = pi_sum; b0 = probs_sum;
          }
        }
        if (G > LPG) {
          if (poll1) {
            w1 = finish_poll2(src1, w1, err);
            a1 = __uint_as_float(w1.x); b1 = __uint_as_float(w1.y);
          } else {
            a1 = pi_sum; b1 = probs_sum;
          }
        }
      }
      pi_sum = cols_rank_tree<G>(a0, a1, lane);
      probs_sum = cols_rank_tree<G>(b0, b1, lane);
      // prob_0 = (y ? EPSILON : 1 - EPSILON) * (1 - pi_sum)   (beta.cc:124-125)
      probs_sum += (y ? a.epsilon : 1.0f - a.epsilon) * (1.0f - pi_sum);
      const float rS = live ? 1.0f / probs_sum : 0.f;
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        if (y) accB[i] = fmaf(q[i], rS, accB[i]);
        else accA[i] = fmaf(q[i], rS, accA[i]);
      }
    }
  }
  // combine: sub-groups of a warp (shuffle tree over the sub-group index), then the CTA's warps in
  // warp order, then (k_cols_beta_reduce) the CTAs in order
#pragma unroll
  for (int i = 0; i < KPL; ++i) {
#pragma unroll
    for (int o = LPG; o < 32; o <<= 1) {
      accA[i] += __shfl_xor_sync(FULL_MASK, accA[i], o);
      accB[i] += __shfl_xor_sync(FULL_MASK, accB[i], o);
    }
  }
  for (uint32_t w = 0; w < WARPS; ++w) {
    if (wib == w && sub == 0) {
#pragma unroll
      for (int i = 0; i < KPL; ++i) {
        const uint32_t c = ((i >> 2) * LPG + li) * 4 + (i & 3);
        if (w == 0) {
          s_acc[0][c] = accA[i];
          s_acc[1][c] = accB[i];
        } else {
          s_acc[0][c] += accA[i];
          s_acc[1][c] += accB[i];
        }
      }
    }
    __syncthreads();
  }
  float* out = me.ws + (size_t)cta * 2 * KG;
  for (uint32_t c = threadIdx.x; c < 2 * KG; c += blockDim.x) out[c] = s_acc[c / KG][c % KG];
}

// fixed-order sum over the CTAs' partials, gradient and Langevin step of the own columns'
// theta (beta.cc:39-82), beta = normalised theta; the new values are published to every rank's
// theta / beta arrays (a rank only ever READS its own columns in kernels, the host reads all).
struct ColsThetaArgs {
  ColsRankView r[AMMSB_MAX_SHARDS];
  ColsBoxLayout lay;
  uint32_t nv, G, K, P;
  float eps_t, eta0, eta1, scale;
};

__device__ __forceinline__ void cols_theta_step(float t0, float t1, float g0, float g1, uint32_t k, float eps_t,
                                                float eta0, float eta1, float scale, ulonglong2* pool, float* out4) {
  Rng s = rng_load(pool, k);
  const float half = __fdiv_rn(eps_t, 2.0f);
  const float r0 = rng_randn(s);
  const float f0 = __fsqrt_rn(__fmul_rn(eps_t, t0));
  t0 = fabsf(__fadd_rn(__fadd_rn(t0, __fmul_rn(half, __fadd_rn(__fsub_rn(eta0, t0), __fmul_rn(scale, g0)))),
                       __fmul_rn(f0, r0)));
  t0 = fmaxf(t0, 1e-24f);
  const float r1 = rng_randn(s);
  const float f1 = __fsqrt_rn(__fmul_rn(eps_t, t1));
  t1 = fabsf(__fadd_rn(__fadd_rn(t1, __fmul_rn(half, __fadd_rn(__fsub_rn(eta1, t1), __fmul_rn(scale, g1)))),
                       __fmul_rn(f1, r1)));
  t1 = fmaxf(t1, 1e-24f);
  rng_store(pool, k, s);
  const float sum = __fadd_rn(__fadd_rn(0.f, t0), t1);
  out4[0] = t0;
  out4[1] = t1;
  out4[2] = __fdiv_rn(t0, sum);
  out4[3] = __fdiv_rn(t1, sum);
}

__global__ void __launch_bounds__(128) k_cols_theta(const __grid_constant__ ColsThetaArgs a) {
  const uint32_t KG = a.K / a.G, LPG = 32 / a.G;
  const uint32_t per_rank = (KG + blockDim.x - 1) / blockDim.x;
  const uint32_t vr = blockIdx.x / per_rank;
  const uint32_t c = (blockIdx.x % per_rank) * blockDim.x + threadIdx.x;
  if (c >= KG) return;
  const ColsRankView& me = a.r[vr];
  // local index c -> global column k
  const uint32_t f = c >> 2, e = c & 3, q = f / LPG, li = f % LPG;
  const uint32_t k = (me.rank + a.G * li) + 32 * (4 * q + e);
  float A = 0.f, B = 0.f;
  for (uint32_t p = 0; p < a.P; ++p) {
    A += me.ws[(size_t)p * 2 * KG + c];
    B += me.ws[(size_t)p * 2 * KG + KG + c];
  }
  unsigned char* mybox = me.box[me.rank];
  const float* theta = reinterpret_cast<const float*>(mybox + a.lay.theta);
  const float t0 = theta[2 * k], t1 = theta[2 * k + 1];
  const float ts = __fadd_rn(t0, t1);
  const float rts = __fdiv_rn(1.0f, ts);
  const float g0 = A * (__fdiv_rn(1.0f, t0) - rts) + B * (0.0f - rts);
  const float g1 = A * (0.0f - rts) + B * (__fdiv_rn(1.0f, t1) - rts);
  float o[4];
  cols_theta_step(t0, t1, g0, g1, k, a.eps_t, a.eta0, a.eta1, a.scale, me.pool, o);
  for (uint32_t p = 0; p < a.G; ++p) {
    if (me.box[p] == nullptr) continue;
    float* th = reinterpret_cast<float*>(me.box[p] + a.lay.theta);
    float* be = reinterpret_cast<float*>(me.box[p] + a.lay.beta);
    *reinterpret_cast<float2*>(th + 2 * k) = make_float2(o[0], o[1]);
    *reinterpret_cast<float2*>(be + 2 * k) = make_float2(o[2], o[3]);
  }
}

// ---- held-out perplexity on the column shards (perplexity.cc:14-83, 251-274) ----
// Sub-group per pair; the two K-wide sums are exchanged; EVERY rank then finishes every pair
// (running mean, log, the four sums) -- redundant scalar work instead of a second exchange, and the
// replicated ppx_per_edge arrays stay identical.
struct ColsPpxArgs {
  ColsRankView r[AMMSB_MAX_SHARDS];
  SetView set;
  ColsBoxLayout lay;
  const uint64_t* edges;
  uint32_t nv, ctas_per_rank;
  uint32_t H, parity, call_count, loopback;
  float epsilon;
};

template <int KPL, int G>
__global__ void __launch_bounds__(128) k_cols_ppx(const __grid_constant__ ColsPpxArgs a) {
  constexpr int LPG = 32 / G, KG = KPL * LPG, Q = KPL / 4, WARPS = 4;
  __shared__ double s_part[WARPS][4];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t sub = lane / LPG, li = lane % LPG;
  const uint32_t vr = blockIdx.x / a.ctas_per_rank, cta = blockIdx.x % a.ctas_per_rank;
  const ColsRankView& me = a.r[vr];
  const uint32_t rank = me.rank;
  unsigned char* mybox = me.box[rank];
  uint32_t* err = reinterpret_cast<uint32_t*>(mybox);
  const float* beta = reinterpret_cast<const float*>(mybox + a.lay.beta);
  const size_t half_P = (size_t)a.parity * G * a.lay.P_src;
  const uint32_t l_ref = rank + G * li;
  float bk[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) bk[i] = beta[2 * (l_ref + 32 * i) + 1];
  double link_lik = 0.0, non_lik = 0.0;
  uint32_t link_cnt = 0, non_cnt = 0;
  const uint32_t trips = (a.H + G - 1) / G;
  for (uint32_t tr = cta * WARPS + wib; tr < trips; tr += a.ctas_per_rank * WARPS) {
    const uint32_t i = tr * G + sub;
    const bool live = i < a.H;
    float sb = 0.f, sq = 0.f;
    bool is_edge = false;
    if (live) {
      const uint64_t e = __ldg(&a.edges[i]);
      const uint32_t u = (uint32_t)(e >> 32), v = (uint32_t)(e & 0xffffffffu);
      const float4* pa = reinterpret_cast<const float4*>(me.pi + (size_t)u * KG);
      const float4* pb = reinterpret_cast<const float4*>(me.pi + (size_t)v * KG);
      is_edge = set_has(a.set, e);  // membership of the key as stored (perplexity.cc:45-47)
#pragma unroll
      for (int qq = 0; qq < Q; ++qq) {
        const float4 x = ldg_stream4(reinterpret_cast<const float*>(pa + qq * LPG + li));
        const float4 z = ldg_stream4(reinterpret_cast<const float*>(pb + qq * LPG + li));
        const float f0 = x.x * z.x, f1 = x.y * z.y, f2 = x.z * z.z, f3 = x.w * z.w;
        sq += f0; sq += f1; sq += f2; sq += f3;
        sb = fmaf(f0, bk[4 * qq], sb);
        sb = fmaf(f1, bk[4 * qq + 1], sb);
        sb = fmaf(f2, bk[4 * qq + 2], sb);
        sb = fmaf(f3, bk[4 * qq + 3], sb);
      }
    }
#pragma unroll
    for (int o = LPG / 2; o > 0; o >>= 1) {
      sb += __shfl_xor_sync(FULL_MASK, sb, o);
      sq += __shfl_xor_sync(FULL_MASK, sq, o);
    }
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    if (live) {
      const size_t idx = (size_t)i * 2;
      if (!a.loopback) {
        const uint32_t x0 = partial_bits(sb), x1 = partial_bits(sq);
        for (uint32_t p = li; p < G; p += LPG)
          if (p != rank) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(me.box[p] + a.lay.P + half_P + (size_t)rank * a.lay.P_src) + idx;
            st_mbox(dst, x0);
            st_mbox(dst + 1, x1);
          }
      }
      if (li < G) {
        if (li == rank || a.loopback) {
          a0 = sb; b0 = sq;
        } else {
          uint32_t* src = reinterpret_cast<uint32_t*>(mybox + a.lay.P + half_P + (size_t)li * a.lay.P_src) + idx;
          a0 = poll_mbox(src, err); b0 = poll_mbox(src + 1, err);
        }
      }
      if (G > LPG) {
        const uint32_t p1 = li + LPG;
        if (p1 == rank || a.loopback) {
          a1 = sb; b1 = sq;
        } else {
          uint32_t* src = reinterpret_cast<uint32_t*>(mybox + a.lay.P + half_P + (size_t)p1 * a.lay.P_src) + idx;
          a1 = poll_mbox(src, err); b1 = poll_mbox(src + 1, err);
        }
      }
    }
    sb = cols_rank_tree<G>(a0, a1, lane);
    sq = cols_rank_tree<G>(b0, b1, lane);
    if (live && li == 0) {
      // calculate_edge_likelihood, perplexity.cc:16-40
      float s = is_edge ? sb : (sq - sb) + (1.0f - sq) * (1.0f - a.epsilon);
      if (s < 1.0e-30f) s = 1.0e-30f;
      float ppx = me.ppx[i];  // running mean over calls, perplexity.cc:51-52
      ppx = __fdiv_rn(__fadd_rn(__fmul_rn(ppx, (float)(a.call_count - 1)), s), (float)a.call_count);
      me.ppx[i] = ppx;
      const float lg = logf(ppx);
      if (is_edge) {
        link_lik += (double)lg;
        ++link_cnt;
      } else {
        non_lik += (double)lg;
        ++non_cnt;
      }
    }
  }
  double c0 = (double)link_cnt, c1 = (double)non_cnt;
  link_lik = warp_sum_d(link_lik);
  non_lik = warp_sum_d(non_lik);
  c0 = warp_sum_d(c0);
  c1 = warp_sum_d(c1);
  if (lane == 0) {
    s_part[wib][0] = link_lik; s_part[wib][1] = non_lik; s_part[wib][2] = c0; s_part[wib][3] = c1;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < WARPS; ++w) s += s_part[w][threadIdx.x];
    me.ws_d[(size_t)cta * 4 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(128) k_cols_ppx_reduce(const __grid_constant__ ColsPpxArgs a, uint32_t P) {
  const ColsRankView& me = a.r[blockIdx.x];
  const uint32_t q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (uint32_t p = lane; p < P; p += 32) s += me.ws_d[(size_t)p * 4 + q];
  s = warp_sum_d(s);
  if (lane == 0) me.ws_d[(size_t)P * 4 + q] = s;
}

// ---- NeighborSampler (sample.cc:13-121) partitioned over the ranks ----
// The reference work-item gid owns sampler state gid and serves the slots gid, gid + gsize, ...;
// rank gid % G runs it (state ownership is static, so the lists do not depend on G) and delivers
// every list to ALL ranks' mailboxes with peer stores.  A list word is self-validating like a
// partial sum (ids are < N < 0xffffffff); the thirds of the region rotate with step % 3, so a word
// is re-armed by its reader two kernel boundaries before it is written again.
struct ColsNsArgs {
  ColsRankView r[AMMSB_MAX_SHARDS];
  ColsBoxLayout lay;
  NsGeom g;
  const uint32_t* nodes;
  uint32_t nv, G, V, gsize, third, per_rank_blocks;
};

// A warp per 32 of the rank's work-items: every lane draws its slot's ids into its own column of a
// shared-memory table (ns_draw_slot, common.cuh), then the warp packs the tables one by one (table
// order, first n entries: sample.cc:64-74) and delivers each list to every rank with ONE coalesced
// store per peer -- n scattered 4-byte peer stores per slot would be bound by the NVLink message
// rate and compete with the partial-sum exchange of update_phi.
__global__ void __launch_bounds__(32) k_cols_neighbor_sample(const __grid_constant__ ColsNsArgs a) {
  extern __shared__ uint32_t s_tab[];  // [capacity][33]
  constexpr uint32_t STRIDE = 33;
  const uint32_t vr = blockIdx.x / a.per_rank_blocks, blk = blockIdx.x % a.per_rank_blocks;
  const ColsRankView& me = a.r[vr];
  const uint32_t lane = threadIdx.x;
  const uint32_t gid0 = me.rank + a.G * (blk * 32);  // the reference work-items of this rank: rank, rank + G, ...
  const uint32_t gid = gid0 + a.G * lane;
  const uint32_t n = a.g.n, N = a.g.N, cap = a.g.capacity;
  const bool owner = gid < a.gsize && gid < a.V;
  Rng seed;
  seed.x = seed.y = 0;
  if (owner) seed = rng_load(me.pool, gid);
  const uint32_t lanes_lt = (1u << lane) - 1;
  const size_t region = a.lay.NB + (size_t)a.third * a.lay.NB_third;
  for (uint32_t pass = 0; gid0 + pass < a.V; pass += a.gsize) {  // warp-uniform: the passes of the warp's first work-item
    const uint32_t i = gid + pass;
    if (owner && i < a.V) ns_draw_slot(seed, __ldg(&a.nodes[i]), a.g, s_tab + lane, STRIDE);
    __syncwarp();
    for (uint32_t t = 0; t < 32; ++t) {
      const uint32_t gt = gid0 + a.G * t;
      if (gt >= a.gsize || gt + pass >= a.V) break;
      const size_t out = (size_t)(gt + pass) * n;
      uint32_t count = 0;
      for (uint32_t c = 0; c < cap && count < n; c += 32) {
        const uint32_t j = c + lane;
        const uint32_t v = j < cap ? s_tab[j * STRIDE + t] : N;
        const uint32_t m = __ballot_sync(FULL_MASK, v != N);
        const uint32_t pos = count + __popc(m & lanes_lt);
        if (v != N && pos < n) {
          for (uint32_t p = 0; p < a.G; ++p) st_mbox(reinterpret_cast<uint32_t*>(me.box[p] + region) + out + pos, v);
        }
        count += __popc(m);
      }
    }
    __syncwarp();
  }
  if (owner) rng_store(me.pool, gid, seed);
}

// ---- init / host access ----
// RandomGammaAndNormalize (random.cc:131-167) on the column shards: every rank walks the whole
// gamma stream of every row (group per row, state = group*32 + lane, pool seeded {11,113}) -- the
// row sum needs all 32 lanes, and recomputing them is cheaper than an exchange at set-up -- and
// stores the columns of its own lanes.
__global__ void k_cols_init_pi(float* pi, float* phi, uint32_t N, uint32_t K, uint32_t G, uint32_t rank, uint32_t groups,
                               float eta0, float eta1) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= groups) return;
  const uint64_t id = (uint64_t)warp * 32 + lane;
  Rng s;
  s.x = 11 + id;
  s.y = 113 + id;
  const uint32_t KG = K / G, LPG = 32 / G;
  const bool mine = lane % G == rank;
  const uint32_t li = lane / G;
  for (uint32_t row = warp; row < N; row += groups) {
    float* r = pi + (size_t)row * KG;
    float lsum = 0.f;
    for (uint32_t i = 0; i < K / 32; ++i) {
      const float g = rng_gamma(s, eta0, eta1);
      lsum = __fadd_rn(lsum, g);
      if (mine) r[((i >> 2) * LPG + li) * 4 + (i & 3)] = g;
    }
    const float sum = warp_sum(lsum);
    if (mine)
      for (uint32_t i = 0; i < K / 32; ++i) {
        float* p = &r[((i >> 2) * LPG + li) * 4 + (i & 3)];
        *p = __fdiv_rn(*p, sum);
      }
    if (lane == 0) phi[row] = sum;
  }
}

// full rows [nrows][K] (global column order) <-> the shard's own columns
__global__ void k_cols_scatter(float* pi, const float* full, uint64_t row0, uint64_t nrows, uint32_t K, uint32_t G,
                               uint32_t rank, bool gather, float* full_out) {
  const uint32_t KG = K / G;
  const uint64_t total = nrows * KG;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = i / KG;
    const uint32_t c = (uint32_t)(i % KG);
    const uint32_t LPG = 32 / G, f = c >> 2, e = c & 3, q = f / LPG, li = f % LPG;
    const uint32_t k = (rank + G * li) + 32 * (4 * q + e);
    if (gather) full_out[r * K + k] = pi[(row0 + r) * KG + c];
    else pi[(row0 + r) * KG + c] = full[r * K + k];
  }
}
#endif  // __CUDACC__

// ------------------------------------------------------------------ host API ----

static int cols_check_shape(uint32_t K, uint32_t G) {
  AMMSB_REQUIRE(G == 2 || G == 4 || G == 8, "column-sharded layout: world must be 2, 4 or 8");
  AMMSB_REQUIRE(K == 128 || K == 256 || K == 512 || K == 1024, "column-sharded layout: K must be 128, 256, 512 or 1024");
  return 0;
}

extern "C" int ammsb_cols_create(ammsb_ctx* c, uint64_t N, uint32_t K, uint32_t world, uint32_t rank,
                                 uint32_t num_neighbors, uint32_t max_nodes, uint32_t max_edges, uint64_t max_pairs,
                                 ammsb_cols** out) {
  if (cols_check_shape(K, world)) return 1;
  AMMSB_REQUIRE(rank < world, "rank out of range");
  AMMSB_REQUIRE(N > 0 && N < 0xffffffffull, "N must fit a 32-bit Vertex (types.h:32)");
  AMMSB_REQUIRE(num_neighbors >= 8 && max_nodes > 0, "num_neighbors must be >= 8");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ammsb_cols* s = new ammsb_cols();
  s->ctx = c;
  s->N = N; s->K = K; s->G = world; s->rank = rank; s->n = num_neighbors;
  s->Vcap = max_nodes; s->Ecap = max_edges; s->Hcap = max_pairs;
  s->KG = K / world;
  s->lay = cols_layout(K, world, num_neighbors, max_nodes, max_edges, max_pairs);
  s->ws_ctas = (uint32_t)c->sm_count * 8;
  cudaError_t e = cudaMalloc((void**)&s->d_pi, sizeof(float) * N * s->KG);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_phi, sizeof(float) * N);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_phi_vec, sizeof(float) * (size_t)max_nodes * s->KG);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_ppx, sizeof(float) * (max_pairs ? max_pairs : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_ws, sizeof(float) * 2 * s->KG * (size_t)s->ws_ctas);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_ws_d, sizeof(double) * 4 * ((size_t)s->ws_ctas + 1));
  s->nz_warps = (size_t)c->sm_count * 32;
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_nz, sizeof(float) * s->nz_warps * K * 2);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_ppx, 0, sizeof(float) * (max_pairs ? max_pairs : 1), c->stream);
  if (e != cudaSuccess || vmm_alloc(c->device, s->lay.bytes, &s->local)) {
    cudaFree(s->d_pi); cudaFree(s->d_phi); cudaFree(s->d_phi_vec); cudaFree(s->d_ppx); cudaFree(s->d_ws); cudaFree(s->d_ws_d); cudaFree(s->d_nz);
    delete s;
    if (e != cudaSuccess) AMMSB_CHECK_CUDA(e);
    return 1;
  }
  s->box[rank] = reinterpret_cast<unsigned char*>(s->local.ptr);
  // every mailbox word starts armed (sentinel); the header (error word) and theta/beta start at 0
  AMMSB_CHECK_CUDA(cudaMemsetAsync(s->box[rank], 0xff, s->lay.bytes, c->stream));
  AMMSB_CHECK_CUDA(cudaMemsetAsync(s->box[rank], 0, s->lay.S, c->stream));
  AMMSB_CHECK_CUDA(cudaStreamSynchronize(c->stream));
  *out = s;
  return 0;
}

extern "C" int ammsb_cols_destroy(ammsb_cols* s) {
  if (!s) return 0;
  cudaSetDevice(s->ctx->device);
  for (int i = 0; i < AMMSB_MAX_SHARDS; ++i) vmm_free(&s->remote[i]);
  vmm_free(&s->local);
  cudaFree(s->d_pi); cudaFree(s->d_phi); cudaFree(s->d_phi_vec); cudaFree(s->d_ppx); cudaFree(s->d_ws); cudaFree(s->d_ws_d); cudaFree(s->d_nz);
  delete s;
  return 0;
}

extern "C" int ammsb_cols_mailbox_bytes(const ammsb_cols* s, size_t* bytes) {
  *bytes = s->lay.bytes;
  return 0;
}

extern "C" int ammsb_cols_export_fd(ammsb_cols* s, int* fd) { return vmm_export_fd(s->local, fd); }

extern "C" int ammsb_cols_attach_fd(ammsb_cols* s, uint32_t peer_rank, int fd) {
  AMMSB_REQUIRE(peer_rank < s->G && peer_rank != s->rank, "bad peer rank");
  if (vmm_import_fd(s->ctx->device, fd, s->lay.bytes, &s->remote[peer_rank])) return 1;
  s->box[peer_rank] = reinterpret_cast<unsigned char*>(s->remote[peer_rank].ptr);
  return 0;
}

// same process: another rank's mailbox by direct (peer) access -- also how one GPU emulates G ranks
extern "C" int ammsb_cols_attach_local(ammsb_cols* s, uint32_t peer_rank, ammsb_cols* peer) {
  AMMSB_REQUIRE(peer_rank < s->G && peer_rank != s->rank && peer->rank == peer_rank, "bad peer rank");
  AMMSB_REQUIRE(peer->N == s->N && peer->K == s->K && peer->G == s->G && peer->lay.bytes == s->lay.bytes,
                "peer does not match");
  AMMSB_CHECK_CUDA(cudaSetDevice(s->ctx->device));
  if (peer->ctx->device != s->ctx->device) {
    int can = 0;
    AMMSB_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, s->ctx->device, peer->ctx->device));
    AMMSB_REQUIRE(can, "devices are not peer-accessible");
    if (vmm_grant(s->ctx->device, peer->local)) return 1;
  }
  s->box[peer_rank] = peer->box[peer_rank];
  return 0;
}

// timing diagnostics (AMMSB_COLS_LOOPBACK=1: the kernels then neither send to nor wait for a peer):
// every unattached peer mailbox pointer becomes an alias of the own mailbox, so that ONE rank's
// share of a G-GPU step can be timed on one GPU.  Results of such a run are meaningless.
extern "C" int ammsb_cols_alias_self(ammsb_cols* s) {
  for (uint32_t p = 0; p < s->G; ++p)
    if (s->box[p] == nullptr) s->box[p] = s->box[s->rank];
  return 0;
}

extern "C" int ammsb_cols_init_pi(ammsb_cols* s, float eta0, float eta1) {
  ammsb_ctx* c = s->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  const uint32_t groups = s->N < 65535 ? (uint32_t)s->N : 65535u;
  k_cols_init_pi<<<(groups * 32 + 127) / 128, 128, 0, c->stream>>>(s->d_pi, s->d_phi, (uint32_t)s->N, s->K, s->G, s->rank,
                                                                  groups, eta0, eta1);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

static int cols_rows_io(ammsb_cols* s, uint64_t row0, uint64_t nrows, float* h, bool gather) {
  AMMSB_REQUIRE(row0 + nrows <= s->N, "rows out of range");
  ammsb_ctx* c = s->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  const uint64_t chunk = (64ull << 20) / (sizeof(float) * s->K) + 1;
  float* d_full = nullptr;
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_full, sizeof(float) * s->K * (nrows < chunk ? nrows : chunk)));
  int rc = 0;
  for (uint64_t r = 0; r < nrows && !rc; r += chunk) {
    const uint64_t cnt = nrows - r < chunk ? nrows - r : chunk;
    if (gather) {
      // columns of other ranks keep what the caller's buffer holds: start from the host content
      rc = ammsb_h2d(c, d_full, h + r * s->K, sizeof(float) * cnt * s->K);
      if (rc) break;
    } else {
      rc = ammsb_h2d(c, d_full, h + r * s->K, sizeof(float) * cnt * s->K);
      if (rc) break;
    }
    k_cols_scatter<<<c->sm_count * 8, 256, 0, c->stream>>>(s->d_pi, d_full, row0 + r, cnt, s->K, s->G, s->rank, gather, d_full);
    g_launch_count.fetch_add(1);
    if (cudaGetLastError() != cudaSuccess) { rc = 1; ammsb_set_error("k_cols_scatter launch failed"); break; }
    if (gather) rc = ammsb_d2h(c, h + r * s->K, d_full, sizeof(float) * cnt * s->K);
    else if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = 1; ammsb_set_error("scatter failed"); }
  }
  cudaFree(d_full);
  return rc;
}

extern "C" int ammsb_cols_write_pi(ammsb_cols* s, uint64_t row0, uint64_t nrows, const float* h_rows) {
  return cols_rows_io(s, row0, nrows, const_cast<float*>(h_rows), false);
}
extern "C" int ammsb_cols_read_pi(ammsb_cols* s, uint64_t row0, uint64_t nrows, float* h_rows) {
  return cols_rows_io(s, row0, nrows, h_rows, true);
}
extern "C" int ammsb_cols_write_phi(ammsb_cols* s, uint64_t row0, uint64_t nrows, const float* h) {
  AMMSB_REQUIRE(row0 + nrows <= s->N, "rows out of range");
  return ammsb_h2d(s->ctx, s->d_phi + row0, h, sizeof(float) * nrows);
}
extern "C" int ammsb_cols_read_phi(ammsb_cols* s, uint64_t row0, uint64_t nrows, float* h) {
  AMMSB_REQUIRE(row0 + nrows <= s->N, "rows out of range");
  return ammsb_d2h(s->ctx, h, s->d_phi + row0, sizeof(float) * nrows);
}
extern "C" int ammsb_cols_write_theta(ammsb_cols* s, const float* h_theta, const float* h_beta) {
  int rc = ammsb_h2d(s->ctx, s->box[s->rank] + s->lay.theta, h_theta, sizeof(float) * 2 * s->K);
  if (!rc) rc = ammsb_h2d(s->ctx, s->box[s->rank] + s->lay.beta, h_beta, sizeof(float) * 2 * s->K);
  return rc;
}
extern "C" int ammsb_cols_read_theta(ammsb_cols* s, float* h_theta, float* h_beta) {
  int rc = 0;
  if (h_theta) rc = ammsb_d2h(s->ctx, h_theta, s->box[s->rank] + s->lay.theta, sizeof(float) * 2 * s->K);
  if (!rc && h_beta) rc = ammsb_d2h(s->ctx, h_beta, s->box[s->rank] + s->lay.beta, sizeof(float) * 2 * s->K);
  return rc;
}
extern "C" int ammsb_cols_beta_ptr(ammsb_cols* s, float** d_theta, float** d_beta) {
  if (d_theta) *d_theta = reinterpret_cast<float*>(s->box[s->rank] + s->lay.theta);
  if (d_beta) *d_beta = reinterpret_cast<float*>(s->box[s->rank] + s->lay.beta);
  return 0;
}
extern "C" int ammsb_cols_read_phi_vec(ammsb_cols* s, uint32_t V, float* h_rows /* [V][K], own columns filled */) {
  AMMSB_REQUIRE(V <= s->Vcap, "V exceeds the capacity");
  ammsb_ctx* c = s->ctx;
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  float* d_full = nullptr;
  AMMSB_CHECK_CUDA(cudaMalloc((void**)&d_full, sizeof(float) * (size_t)V * s->K));
  int rc = ammsb_h2d(c, d_full, h_rows, sizeof(float) * (size_t)V * s->K);
  if (!rc) {
    k_cols_scatter<<<c->sm_count * 8, 256, 0, c->stream>>>(s->d_phi_vec, d_full, 0, V, s->K, s->G, s->rank, true, d_full);
    g_launch_count.fetch_add(1);
    rc = ammsb_d2h(c, h_rows, d_full, sizeof(float) * (size_t)V * s->K);
  }
  cudaFree(d_full);
  return rc;
}
// 0 = no exchange wait has timed out on this rank so far
extern "C" int ammsb_cols_check(ammsb_cols* s, uint32_t* timed_out) {
  return ammsb_d2h(s->ctx, timed_out, s->box[s->rank], 4);
}

// The exchange kernels are persistent, fill the SMs they are given and wait for the other ranks.
// A library kernel that ALSO waits for a peer (an NCCL broadcast on another stream) must always
// find an SM, or the two waits can lock each other out across GPUs: the cooperative grids leave a
// few SMs free (AMMSB_COLS_SPARE_SMS, default 4; pair with NCCL_MAX_NCHANNELS <= that).
static uint32_t cols_usable_sms(const ammsb_ctx* c) {
  uint32_t spare = 4;
  if (const char* e = getenv("AMMSB_COLS_SPARE_SMS")) spare = (uint32_t)atoi(e);
  return (uint32_t)c->sm_count > spare + 1 ? (uint32_t)c->sm_count - spare : 1u;
}

static int cols_ready(ammsb_cols* const* ranks, uint32_t nv) {
  AMMSB_REQUIRE(nv >= 1 && nv <= AMMSB_MAX_SHARDS, "bad number of ranks in one launch");
  for (uint32_t i = 0; i < nv; ++i) {
    AMMSB_REQUIRE(ranks[i] && ranks[i]->ctx->device == ranks[0]->ctx->device && ranks[i]->G == ranks[0]->G &&
                      ranks[i]->K == ranks[0]->K && ranks[i]->N == ranks[0]->N,
                  "ranks of one launch must live on one device and agree in shape");
    for (uint32_t p = 0; p < ranks[i]->G; ++p) AMMSB_REQUIRE(ranks[i]->box[p] != nullptr, "a peer mailbox is not attached");
  }
  return 0;
}

template <class Kern, class Args>
static int cols_launch_coop(ammsb_ctx* c, Kern kern, const Args& a, uint32_t grid, uint32_t block, size_t smem) {
  void* params[] = {const_cast<Args*>(&a)};
  AMMSB_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(block), params, smem,
                                               c->stream));
  AMMSB_LAUNCH_CHECK();
  return 0;
}

struct ColsTune {  // launch shape of k_cols_phi, overridable for A/B measurements
  uint32_t warps, R, D;
};
static ColsTune cols_tune(uint32_t KPL, uint32_t G, uint32_t n) {
  ColsTune t;
  t.warps = 8; t.R = 4; t.D = 2;
  if (const char* e = getenv("AMMSB_COLS_WARPS")) t.warps = (uint32_t)atoi(e);
  if (const char* e = getenv("AMMSB_COLS_R")) t.R = (uint32_t)atoi(e);
  if (const char* e = getenv("AMMSB_COLS_D")) t.D = (uint32_t)atoi(e);
  if (t.R > n) t.R = n;
  if (t.R < 2) t.R = 2;
  if (t.D >= t.R) t.D = t.R - 1;
  if (t.D < 1) t.D = 1;  // phase B of a stage runs after its phase A
  if (t.warps < 1) t.warps = 1;
  if (t.warps > 8) t.warps = 8;
  return t;
}

template <int KPL, int G, int NB>
static int cols_phi_launch_nb(ammsb_ctx* c, ColsPhiArgs& a, uint32_t nv);

// k_cols_phi2 (slot at a time): as many warps as the staged slot buffers allow, every CTA resident
template <int KPL, int G>
static int cols_phi2_launch(ammsb_ctx* c, ColsPhiArgs& a, uint32_t nv, float* d_nz, size_t nz_warps) {
  constexpr uint32_t KG = KPL * (32 / G);
  uint32_t warps = ColsPhi2Smem::max_warps(KG);  // the staged slot buffers decide (12 at K/G = 128)
  if (const char* e = getenv("AMMSB_COLS_WARPS")) warps = (uint32_t)atoi(e);
  if (warps < 1) warps = 1;
  if (warps > ColsPhi2Smem::max_warps(KG)) warps = ColsPhi2Smem::max_warps(KG);
  size_t smem;
  for (;; --warps) {
    smem = ColsPhi2Smem::per_cta(KG) + (size_t)warps * ColsPhi2Smem::per_warp(KG);
    if (smem <= c->smem_optin || warps == 1) break;
  }
  AMMSB_REQUIRE(smem <= c->smem_optin, "column update_phi: shared memory request exceeds the device limit");
  const uint32_t active = a.units < a.V ? a.units : a.V;
  const uint32_t ngroups = (active + G - 1) / G;
  // few slots (a link mini-batch): a warp per SLOT instead of a warp per group of G slots, so that
  // the slots' load / exchange / Langevin chains run side by side (CTAs of a multiple of G warps)
  const uint32_t wsplit = warps / G * G;
  a.split = wsplit > 0 && (size_t)ngroups * G <= (size_t)cols_usable_sms(c) / nv * wsplit && !getenv("AMMSB_COLS_NOSPLIT");
  if (a.split) {
    warps = wsplit;
    smem = ColsPhi2Smem::per_cta(KG) + (size_t)warps * ColsPhi2Smem::per_warp(KG);
  }
  auto kern = k_cols_phi2<KPL, G>;
  static bool attr_set[64] = {false};  // per device
  if (!attr_set[c->device & 63]) {
    AMMSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
    attr_set[c->device & 63] = true;
  }
  int occ = 0;
  AMMSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
  AMMSB_REQUIRE(occ > 0, "column update_phi: kernel does not fit on an SM");
  const uint32_t resident = (uint32_t)occ * cols_usable_sms(c);
  uint32_t ctas = resident / nv;
  const uint32_t tasks = a.split ? ngroups * G : ngroups;
  if (ctas * warps > tasks) ctas = (tasks + warps - 1) / warps;
  if (ctas > 0 && !a.split) {  // an even share of groups per warp (a static schedule: the slowest warp ends the kernel)
    const uint32_t per = (ngroups + ctas * warps - 1) / (ctas * warps);
    const uint32_t need = (ngroups + per - 1) / per;
    ctas = (need + warps - 1) / warps;
  }
  AMMSB_REQUIRE(ctas > 0, "column update_phi: no resident CTA available per rank");
  AMMSB_REQUIRE((size_t)ctas * nv * warps <= nz_warps, "column update_phi: noise scratch too small");
  a.ctas_per_rank = ctas;
  a.nv = nv;
  void* params[] = {&a, &d_nz};
  AMMSB_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(ctas * nv), dim3(warps * 32), params,
                                               smem, c->stream));
  AMMSB_LAUNCH_CHECK();
  return 0;
}

template <int KPL, int G>
static int cols_phi_launch(ammsb_ctx* c, ColsPhiArgs& a, uint32_t nv) {
  // stages per trip: 2 when the neighbor count allows it (AMMSB_COLS_NB=1 for A/B measurements)
  uint32_t nb = (a.n % 2 == 0) ? 2 : 1;
  if (const char* e = getenv("AMMSB_COLS_NB")) nb = (uint32_t)atoi(e);
  if (nb == 4 && a.n % 4 == 0) return cols_phi_launch_nb<KPL, G, 4>(c, a, nv);
  if (nb >= 2 && a.n % 2 == 0) return cols_phi_launch_nb<KPL, G, 2>(c, a, nv);
  return cols_phi_launch_nb<KPL, G, 1>(c, a, nv);
}

template <int KPL, int G, int NB>
static int cols_phi_launch_nb(ammsb_ctx* c, ColsPhiArgs& a, uint32_t nv) {
  using SM = ColsPhiSmem<KPL, G>;
  ColsTune t = cols_tune(KPL, G, a.n);
  // whole trips: R and D multiples of NB, D >= NB, R > D
  t.D = (t.D + NB - 1) / NB * NB;
  if (t.D < (uint32_t)NB) t.D = NB;
  t.R = (t.R + NB - 1) / NB * NB;
  if (t.R <= t.D) t.R = t.D + NB;
  AMMSB_REQUIRE(t.R <= a.n && t.R <= 32, "column update_phi: too few neighbors for the stage ring");
  a.R = t.R;
  a.D = t.D;
  const uint32_t min_seg = a.n % 32 ? a.n % 32 : 32;
  a.MB = 2 + (t.R - 1 + min_seg - 1) / min_seg;
  uint32_t warps = t.warps;
  size_t smem;
  for (;; --warps) {
    smem = 1536 + (size_t)warps * ((SM::per_warp(a.R, a.MB) + 127) / 128 * 128);
    if (smem <= c->smem_optin || warps == 1) break;
  }
  AMMSB_REQUIRE(smem <= c->smem_optin, "column update_phi: shared memory request exceeds the device limit");
  auto kern = k_cols_phi<KPL, G, NB>;
  static bool attr_set[64] = {false};  // per device
  if (!attr_set[c->device & 63]) {
    AMMSB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_optin));
    attr_set[c->device & 63] = true;
  }
  int occ = 0;
  AMMSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
  AMMSB_REQUIRE(occ > 0, "column update_phi: kernel does not fit on an SM");
  const uint32_t resident = (uint32_t)occ * c->sm_count;
  const uint32_t active = a.units < a.V ? a.units : a.V;
  const uint32_t ngroups = (active + G - 1) / G;
  uint32_t ctas = resident / nv;
  // no more warps than groups, and an even share of groups per warp (a static schedule: the
  // slowest warp ends the kernel)
  const uint32_t want_warps = ngroups;
  if (ctas * warps > want_warps) ctas = (want_warps + warps - 1) / warps;
  if (ctas > 0) {
    const uint32_t per = (ngroups + ctas * warps - 1) / (ctas * warps);
    const uint32_t need = (ngroups + per - 1) / per;  // warps that get `per` groups (the last maybe fewer)
    ctas = (need + warps - 1) / warps;
  }
  AMMSB_REQUIRE(ctas > 0, "column update_phi: no resident CTA available per rank");
  a.ctas_per_rank = ctas;
  a.nv = nv;
  return cols_launch_coop(c, kern, a, ctas * nv, warps * 32, smem);
}

extern "C" int ammsb_cols_update_phi(ammsb_ctx* c, ammsb_cols* const* ranks, uint32_t nv, const ammsb_params* p,
                                     const ammsb_phi_opts* o, ammsb_set* train, const uint32_t* d_nodes,
                                     const uint32_t* d_neighbors, uint32_t V, uint32_t step_count,
                                     ammsb_rng* const* pools) {
  if (cols_ready(ranks, nv)) return 1;
  const ammsb_cols* s0 = ranks[0];
  AMMSB_REQUIRE(V > 0, "mini-batch nodes size = 0!");  // phi.cc:732
  AMMSB_REQUIRE(V <= s0->Vcap, "mini-batch larger than the capacity the store was created with");
  AMMSB_REQUIRE(p->K == s0->K && p->N == s0->N && p->num_neighbors == s0->n, "params do not match the store");
  AMMSB_REQUIRE(o->mode == AMMSB_MODE_WG && o->wg == 32, "the column layout reproduces the WG launch with wg = 32");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ColsPhiArgs a;
  memset(&a, 0, sizeof a);
  a.units = cols_units(V);
  const uint64_t states = (uint64_t)a.units * 32;
  for (uint32_t i = 0; i < nv; ++i) {
    AMMSB_REQUIRE(o->disable_noise || (pools && pools[i] && pools[i]->n >= states), "Num seeds smaller than global threads");
    a.r[i] = ranks[i]->view(pools && pools[i] ? pools[i]->d_state : nullptr);
  }
  a.set = train->view();
  a.lay = s0->lay;
  a.nodes = d_nodes;
  a.neighbors = d_neighbors;
  a.nb_poll = d_neighbors == nullptr;  // the lists of ammsb_cols_neighbor_sample (per rank: set below)
  a.nb_third = step_count % 3;
  a.V = V;
  a.n = p->num_neighbors;
  a.parity = step_count & 1;
  a.disable_noise = o->disable_noise;
  a.loopback = getenv("AMMSB_COLS_LOOPBACK") != nullptr;
  a.debug = getenv("AMMSB_COLS_DEBUG") ? (uint32_t)atoi(getenv("AMMSB_COLS_DEBUG")) : 0;
  a.fake_lat = (a.loopback && getenv("AMMSB_COLS_FAKE_LAT")) ? (uint32_t)(atof(getenv("AMMSB_COLS_FAKE_LAT")) * 1.965) : 0;
  a.eps_t = ammsb_eps_t(p, step_count);
  a.alpha = p->alpha;
  a.epsilon = p->epsilon;
  a.Nn = (1.0f * p->N) / p->num_neighbors;  // phi.cc:113
  const uint32_t kpl = p->K / 32, G = s0->G;
  const bool staged = getenv("AMMSB_COLS_STAGED") != nullptr;  // the per-neighbor stage-ring kernel (A/B measurements)
  AMMSB_REQUIRE(!(staged && a.nb_poll), "the stage-ring kernel needs the neighbor lists as an argument");
#define COLS_PHI_CASE(KPL_, G_)                                                                              \
  if (kpl == KPL_ && G == G_)                                                                                \
    return staged ? cols_phi_launch<KPL_, G_>(c, a, nv)                                                      \
                  : cols_phi2_launch<KPL_, G_>(c, a, nv, ranks[0]->d_nz, ranks[0]->nz_warps);
  COLS_PHI_CASE(4, 2) COLS_PHI_CASE(4, 4) COLS_PHI_CASE(4, 8)
  COLS_PHI_CASE(8, 2) COLS_PHI_CASE(8, 4) COLS_PHI_CASE(8, 8)
  COLS_PHI_CASE(16, 2) COLS_PHI_CASE(16, 4) COLS_PHI_CASE(16, 8)
  COLS_PHI_CASE(32, 2) COLS_PHI_CASE(32, 4) COLS_PHI_CASE(32, 8)
#undef COLS_PHI_CASE
  AMMSB_REQUIRE(false, "unsupported (K, world) for the column layout");
  return 1;
}

extern "C" int ammsb_cols_neighbor_sample(ammsb_ctx* c, ammsb_cols* const* ranks, uint32_t nv, const uint32_t* d_nodes,
                                          uint32_t V, uint32_t wg, uint32_t step_count, ammsb_rng* const* pools) {
  if (cols_ready(ranks, nv)) return 1;
  const ammsb_cols* s0 = ranks[0];
  AMMSB_REQUIRE(V > 0 && V <= s0->Vcap, "bad mini-batch size");  // learner.cc:179
  AMMSB_REQUIRE(wg > 0 && (uint64_t)s0->n < s0->N, "bad sampler geometry");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ColsNsArgs a;
  memset(&a, 0, sizeof a);
  // sample.cc:116-119: global = min(ceil(V/wg), 65535/wg) * wg
  uint32_t groups = V / wg + (V % wg ? 1 : 0);
  if (groups > 65535u / wg) groups = 65535u / wg;
  a.gsize = groups * wg;
  const uint32_t active = a.gsize < V ? a.gsize : V;
  for (uint32_t i = 0; i < nv; ++i) {
    AMMSB_REQUIRE(pools && pools[i] && pools[i]->n >= active, "Num seeds smaller than global threads");
    a.r[i] = ranks[i]->view(pools[i]->d_state);
  }
  a.lay = s0->lay;
  a.g = ns_geom((uint32_t)s0->N, s0->n);
  a.nodes = d_nodes;
  a.nv = nv;
  a.G = s0->G;
  a.V = V;
  a.third = step_count % 3;
  const uint32_t block = 32;
  const uint32_t mine = (active + s0->G - 1) / s0->G;  // work-items of one rank
  a.per_rank_blocks = (mine + block - 1) / block;
  const size_t smem = sizeof(uint32_t) * 2 * s0->n * 33;
  AMMSB_REQUIRE(smem <= c->smem_optin, "num_node_sample too large for the shared-memory table");
  if (smem > 48 * 1024)
    AMMSB_CHECK_CUDA(cudaFuncSetAttribute(k_cols_neighbor_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_cols_neighbor_sample<<<a.per_rank_blocks * nv, block, smem, c->stream>>>(a);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_cols_update_pi(ammsb_ctx* c, ammsb_cols* const* ranks, uint32_t nv, const uint32_t* d_nodes,
                                    uint32_t V, uint32_t step_count) {
  if (cols_ready(ranks, nv)) return 1;
  AMMSB_REQUIRE(V > 0 && V <= ranks[0]->Vcap, "bad mini-batch size");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ColsPiArgs a;
  memset(&a, 0, sizeof a);
  for (uint32_t i = 0; i < nv; ++i) a.r[i] = ranks[i]->view(nullptr);
  a.lay = ranks[0]->lay;
  a.nodes = d_nodes;
  a.nv = nv;
  a.G = ranks[0]->G;
  a.KG = ranks[0]->KG;
  a.V = V;
  a.units = cols_units(V);
  a.parity = step_count & 1;
  uint32_t ctas = (V + 7) / 8;
  const uint32_t cap = (uint32_t)c->sm_count * 4 / nv;  // every CTA resident: peers' partials may still be in flight
  if (ctas > cap) ctas = cap ? cap : 1;
  a.ctas_per_rank = ctas;
  k_cols_pi<<<ctas * nv, 256, 0, c->stream>>>(a);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

template <int KPL, int G>
static int cols_beta_launch(ammsb_ctx* c, ColsBetaArgs& a, uint32_t nv, uint32_t ws_ctas) {
  const uint32_t trips = (a.E_mb + G - 1) / G;
  uint32_t ctas = (trips + 3) / 4;
  int occ = 0;  // every CTA must be resident: its warps wait for the same warps of the peers
  AMMSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cols_beta<KPL, G>, 128, 0));
  uint32_t cap = cols_usable_sms(c) * (uint32_t)occ / nv;
  if (cap > ws_ctas) cap = ws_ctas;
  if (ctas > cap) ctas = cap;
  if (ctas == 0) ctas = 1;
  a.ctas_per_rank = ctas;
  a.nv = nv;
  return cols_launch_coop(c, k_cols_beta<KPL, G>, a, ctas * nv, 128, 0);
}

extern "C" int ammsb_cols_update_beta(ammsb_ctx* c, ammsb_cols* const* ranks, uint32_t nv, const ammsb_params* p,
                                      ammsb_set* train, const uint64_t* d_edges, uint32_t E_mb, float scale,
                                      uint32_t step_count, ammsb_rng* const* pools) {
  if (cols_ready(ranks, nv)) return 1;
  const ammsb_cols* s0 = ranks[0];
  AMMSB_REQUIRE(p->K == s0->K && p->N == s0->N, "params do not match the store");
  AMMSB_REQUIRE(E_mb <= s0->Ecap, "mini-batch larger than the capacity the store was created with");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ColsBetaArgs a;
  memset(&a, 0, sizeof a);
  for (uint32_t i = 0; i < nv; ++i) {
    AMMSB_REQUIRE(pools && pools[i] && pools[i]->n >= p->K, "beta RNG pool smaller than K");
    a.r[i] = ranks[i]->view(pools[i]->d_state);
  }
  a.set = train->view();
  a.lay = s0->lay;
  a.edges = d_edges;
  a.E_mb = E_mb;
  a.parity = step_count & 1;
  a.loopback = getenv("AMMSB_COLS_LOOPBACK") != nullptr;
  a.epsilon = p->epsilon;
  const uint32_t kpl = p->K / 32, G = s0->G;
  int rc = 1;
  bool found = false;
#define COLS_BETA_CASE(KPL_, G_) \
  if (!found && kpl == KPL_ && G == G_) { found = true; rc = cols_beta_launch<KPL_, G_>(c, a, nv, s0->ws_ctas); }
  COLS_BETA_CASE(4, 2) COLS_BETA_CASE(4, 4) COLS_BETA_CASE(4, 8)
  COLS_BETA_CASE(8, 2) COLS_BETA_CASE(8, 4) COLS_BETA_CASE(8, 8)
  COLS_BETA_CASE(16, 2) COLS_BETA_CASE(16, 4) COLS_BETA_CASE(16, 8)
  COLS_BETA_CASE(32, 2) COLS_BETA_CASE(32, 4) COLS_BETA_CASE(32, 8)
#undef COLS_BETA_CASE
  AMMSB_REQUIRE(found, "unsupported (K, world) for the column layout");
  if (rc) return rc;
  ColsThetaArgs t;
  memset(&t, 0, sizeof t);
  for (uint32_t i = 0; i < nv; ++i) t.r[i] = a.r[i];
  t.lay = s0->lay;
  t.nv = nv;
  t.G = G;
  t.K = p->K;
  t.P = a.ctas_per_rank;
  t.eps_t = ammsb_eps_t(p, step_count);
  t.eta0 = p->eta0;
  t.eta1 = p->eta1;
  t.scale = scale;
  const uint32_t per_rank = (s0->KG + 127) / 128;
  k_cols_theta<<<per_rank * nv, 128, 0, c->stream>>>(t);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

template <int KPL, int G>
static int cols_ppx_launch(ammsb_ctx* c, ColsPpxArgs& a, uint32_t nv, uint32_t ws_ctas) {
  const uint32_t trips = (a.H + G - 1) / G;
  uint32_t ctas = (trips + 3) / 4;
  int occ = 0;
  AMMSB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cols_ppx<KPL, G>, 128, 0));
  uint32_t cap = cols_usable_sms(c) * (uint32_t)occ / nv;
  if (cap > ws_ctas) cap = ws_ctas;
  if (ctas > cap) ctas = cap;
  if (ctas == 0) ctas = 1;
  a.ctas_per_rank = ctas;
  a.nv = nv;
  if (cols_launch_coop(c, k_cols_ppx<KPL, G>, a, ctas * nv, 128, 0)) return 1;
  k_cols_ppx_reduce<<<nv, 128, 0, c->stream>>>(a, ctas);
  AMMSB_LAUNCH_CHECK();
  return 0;
}

extern "C" int ammsb_cols_perplexity(ammsb_ctx* c, ammsb_cols* const* ranks, uint32_t nv, const ammsb_params* p,
                                     ammsb_set* heldout, const uint64_t* d_edges, uint32_t H, uint32_t call_count,
                                     double* h_sums /* [nv][4] or NULL */, double* h_avg /* [nv] or NULL */) {
  if (cols_ready(ranks, nv)) return 1;
  const ammsb_cols* s0 = ranks[0];
  AMMSB_REQUIRE(p->K == s0->K && p->N == s0->N, "params do not match the store");
  AMMSB_REQUIRE(H <= s0->Hcap, "more held-out pairs than the capacity the store was created with");
  AMMSB_REQUIRE(call_count >= 1, "call_count is 1-based (perplexity.cc:252)");
  AMMSB_CHECK_CUDA(cudaSetDevice(c->device));
  ColsPpxArgs a;
  memset(&a, 0, sizeof a);
  for (uint32_t i = 0; i < nv; ++i) a.r[i] = ranks[i]->view(nullptr);
  a.set = heldout->view();
  a.lay = s0->lay;
  a.edges = d_edges;
  a.H = H;
  a.parity = call_count & 1;
  a.call_count = call_count;
  a.loopback = getenv("AMMSB_COLS_LOOPBACK") != nullptr;
  a.epsilon = p->epsilon;
  const uint32_t kpl = p->K / 32, G = s0->G;
  int rc = 1;
  bool found = false;
#define COLS_PPX_CASE(KPL_, G_) \
  if (!found && kpl == KPL_ && G == G_) { found = true; rc = cols_ppx_launch<KPL_, G_>(c, a, nv, s0->ws_ctas); }
  COLS_PPX_CASE(4, 2) COLS_PPX_CASE(4, 4) COLS_PPX_CASE(4, 8)
  COLS_PPX_CASE(8, 2) COLS_PPX_CASE(8, 4) COLS_PPX_CASE(8, 8)
  COLS_PPX_CASE(16, 2) COLS_PPX_CASE(16, 4) COLS_PPX_CASE(16, 8)
  COLS_PPX_CASE(32, 2) COLS_PPX_CASE(32, 4) COLS_PPX_CASE(32, 8)
#undef COLS_PPX_CASE
  AMMSB_REQUIRE(found, "unsupported (K, world) for the column layout");
  if (rc) return rc;
  for (uint32_t i = 0; i < nv; ++i) ranks[i]->ppx_ctas = a.ctas_per_rank;
  if (!h_sums && !h_avg) return 0;
  for (uint32_t i = 0; i < nv; ++i) {
    rc = ammsb_cols_perplexity_result(c, ranks[i], h_sums ? h_sums + 4 * i : nullptr, h_avg ? h_avg + i : nullptr);
    if (rc) return rc;
  }
  return 0;
}

// the four sums / the average of the last ammsb_cols_perplexity launch of this rank (waits for the
// stream).  A caller that drives several devices from one thread launches every device's kernel
// with NULL outputs first -- the kernels wait for each other -- and collects the results here.
extern "C" int ammsb_cols_perplexity_result(ammsb_ctx* c, ammsb_cols* s, double* h_sums, double* h_avg) {
  AMMSB_REQUIRE(s->ppx_ctas > 0, "no perplexity launch to read");
  double sums[4];
  const int rc = ammsb_d2h(c, sums, s->d_ws_d + (size_t)s->ppx_ctas * 4, sizeof sums);
  if (rc) return rc;
  if (h_sums) for (int q = 0; q < 4; ++q) h_sums[q] = sums[q];
  double avg = 0.0;  // perplexity.cc:264-273
  if (sums[2] + sums[3] != 0) avg = (sums[0] + sums[1]) / (sums[2] + sums[3]);
  if (h_avg) *h_avg = -avg;
  return 0;
}
