// oracle/ref_prelude.h -- TEST INFRASTRUCTURE ONLY.
// Host-side stand-in for the reference's kernel dialect (types.cc:407-446 make_base_macros,
// :491-520 GetClTypes) so that the reference's kernel text compiles and runs as host C++:
// one emulated work-item at a time, work-groups with barriers as round-robin fibers.
#ifndef ORACLE_REF_PRELUDE_H_
#define ORACLE_REF_PRELUDE_H_

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace refrt {

struct WorkItem {
  size_t gid, gsize, lid, lsize, grp, ngrp;
};
extern thread_local WorkItem wi;

// MakeCompileFlags (config.cc:66-83) -> run-time globals instead of -D literals
struct KernelConfig {
  int K, N, E, NUM_NEIGHBORS;
  float ALPHA, EPS_A, EPS_B, EPS_C, EPSILON, ETA0, ETA1;
  int disable_noise;
};
extern KernelConfig kc;

void wg_barrier();
// run `fn` for every work-item of `ngroups` groups of `lsize`; with_barriers selects fibers
void launch(size_t ngroups, size_t lsize, bool with_barriers, const std::function<void()>& fn);

}  // namespace refrt

#define Float float
#define Float2 float2
#define Float4 float4
#define FL(X) (X##f)
#define KERNEL static
#define GLOBAL
#define LOCAL
#define LOCAL_DECLARE static thread_local
#define CONSTANT static const
#define GET_GLOBAL_ID() (::refrt::wi.gid)
#define GET_GLOBAL_SIZE() (::refrt::wi.gsize)
#define GET_LOCAL_ID() (::refrt::wi.lid)
#define GET_LOCAL_SIZE() (::refrt::wi.lsize)
#define GET_NUM_GROUPS() (::refrt::wi.ngrp)
#define GET_GROUP_ID() (::refrt::wi.grp)
#define BARRIER_LOCAL ::refrt::wg_barrier()
#define BARRIER_GLOBAL ::refrt::wg_barrier()
#define FABS fabsf
#define EXP expf
#define SQRT sqrtf
#define LOG logf
#define POW powf
#define MAX fmaxf
#ifdef ULONG_MAX
#undef ULONG_MAX
#endif
#define ULONG_MAX 0xffffffffffffffffUL

// the -D flags of MakeCompileFlags (config.cc:66-83)
#define FLOAT_TYPE float
#define VERTEX_TYPE uint
#define EDGE_TYPE ulong
#define K (::refrt::kc.K)
#define N (::refrt::kc.N)
#define E (::refrt::kc.E)
#define NUM_NEIGHBORS (::refrt::kc.NUM_NEIGHBORS)
#define ALPHA (::refrt::kc.ALPHA)
#define EPS_A (::refrt::kc.EPS_A)
#define EPS_B (::refrt::kc.EPS_B)
#define EPS_C (::refrt::kc.EPS_C)
#define EPSILON (::refrt::kc.EPSILON)
#define ETA0 (::refrt::kc.ETA0)
#define ETA1 (::refrt::kc.ETA1)
// phi.cc:673-677
#define PHI_RANDN(X) (::refrt::kc.disable_noise ? 1 : randn(X))
// random.cl.inc:4-8 (C++ string literals in the reference, not part of the raw string)
#define PARAM_R FL(3.44428647676)
#define gsl_rng_get(randomState) rand(randomState)
#define gsl_rng_uniform(randomState) random(randomState)
#define gsl_rng_uniform_int(randomState, n) randint(randomState, 0, n)
// phi.cc:203-212,679-682 / beta.cc:140-141,290-293 / perplexity.cc:88-90 name aliases
#define FloatRowPartitionedMatrix floatRowPartitionedMatrix
#define FloatRowPartitionedMatrix_Row floatRowPartitionedMatrix_Row
#define WG_SUM_Float WG_SUM_float
#define WG_SUM_Float_LOCAL_ WG_SUM_float_LOCAL_
#define WG_SUML_Float WG_SUML_float
#define WG_NORMALIZE_Float WG_NORMALIZE_float

#endif  // ORACLE_REF_PRELUDE_H_
