// K2: fused T-step rollout.  The state of two environments lives in the registers of one thread for the whole
// rollout; the policy (constant escapement, MSY, or a caller-supplied action tensor) is applied in-kernel, the
// Gaussian noise comes from the per-pair Philox stream, and HBM is touched only at entry/exit (plus the optional
// action tensor / traces).  FP64-pipe bound when no trace is requested.  This is the loop of simulate_mdp
// (gym_fishing/envs/shared_env.py:29-54) with escapement.predict / msy.predict (models/policies.py:16-31) and
// BaseFishingEnv.step (envs/base_fishing_env.py:60-81) inlined.
// K4: fixed-shape reduction of per-env returns to per-group statistics.
#include <cuda_runtime.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "dynamics.cuh"
#include "kernel_args.cuh"
#include "philox.cuh"

#ifndef FISHING_K2_UNROLL
#define FISHING_K2_UNROLL 2
#endif
#ifndef FISHING_K2_MINBLOCKS
#define FISHING_K2_MINBLOCKS 0   // resident 128-thread blocks per SM the fused kernel is compiled for (0: no register cap)
#endif

namespace fishing {

constexpr int K2_UNROLL = FISHING_K2_UNROLL;
#ifndef FISHING_K2_FAST_BS
#define FISHING_K2_FAST_BS 128    // threads per block of rollout_fast_kernel (timing builds: 64 / 32 with FISHING_K2_MINBLOCKS 9 / 18)
#endif
constexpr int K2_FAST_BS = FISHING_K2_FAST_BS;
// Timing build (-DFISHING_K2_TIPPING_WARPS5=1, profiles/ab_r02.md r02q): the tipping-point instantiations in 64-thread blocks compiled for
// 9 per SM (92 - 96 registers without a spill, 10 resident blocks = a fifth warp per scheduler).  Plain mode gains 3 % (2.10 -> 2.17e11),
// the freeze-on-done instantiation that cfg 3 runs loses 1 - 2 % (bench.py: 2.01 -> 1.98e11), the logistic kernel loses 8 %: not the default.
#ifndef FISHING_K2_TIPPING_WARPS5
#define FISHING_K2_TIPPING_WARPS5 0
#endif
__host__ __device__ constexpr int fast_bs(int growth) { return (FISHING_K2_TIPPING_WARPS5 && growth == TIPPING) ? 64 : K2_FAST_BS; }
__host__ __device__ constexpr int fast_minblocks(int growth) { return (FISHING_K2_TIPPING_WARPS5 && growth == TIPPING) ? 9 : FISHING_K2_MINBLOCKS; }
#ifndef FISHING_SWEEP_UNROLL
#define FISHING_SWEEP_UNROLL 4
#endif
constexpr int SWEEP_UNROLL = FISHING_SWEEP_UNROLL;   // inner loop of the warp-cooperative sweep kernel   // timing builds may unroll the fused loop (tools/ab_build.py)

// FISHING_GENERIC_ROLLOUT=1 forces the generic K2 kernel (A/B checks of the fast path)
static bool generic_only() {
  static const bool on = [] {
    const char* e = std::getenv("FISHING_GENERIC_ROLLOUT");
    return e && e[0] && e[0] != '0';
  }();
  return on;
}

// A/B checks of the sweep kernels: FISHING_SWEEP_PLAIN=1 keeps sweeps on the one-pair-per-thread kernel, =2 on the
// per-thread amortised kernel (LV levels per thread); default: the warp-cooperative kernel when the level count allows.
static int sweep_mode() {
  static const int mode = [] {
    const char* e = std::getenv("FISHING_SWEEP_PLAIN");
    return (e && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : 0;
  }();
  return mode;
}

// The two environments a thread owns.  Default layout: adjacent ids (2j, 2j+1) -> 16-byte vector accesses.
// Sweep layout (noise_id_div = m > 1, env id = replicate*m + level): thread = (replicate pair, level), i.e. ids
// (2kp*m + j, (2kp+1)*m + j) -> two coalesced 8-byte accesses; both share the level and use the two halves of the
// replicate pair's noise.
struct Lanes {
  int64_t i0, i1;
  bool vec;
};

template <typename T2, typename T>
__device__ __forceinline__ T2 load2(const T* base, const Lanes& L) {
  if (L.vec) return *reinterpret_cast<const T2*>(base + L.i0);
  T2 v;
  v.x = base[L.i0];
  v.y = base[L.i1];
  return v;
}
template <typename T2, typename T>
__device__ __forceinline__ T2 load2_stream(const T* base, const Lanes& L) {
  if (L.vec) return __ldcs(reinterpret_cast<const T2*>(base + L.i0));
  T2 v;
  v.x = __ldcs(base + L.i0);
  v.y = __ldcs(base + L.i1);
  return v;
}
template <typename T2, typename T>
__device__ __forceinline__ void store2(T* base, const Lanes& L, T a, T b) {
  if (L.vec) {
    T2 v;
    v.x = a;
    v.y = b;
    *reinterpret_cast<T2*>(base + L.i0) = v;
  } else {
    base[L.i0] = a;
    base[L.i1] = b;
  }
}
template <typename T2, typename T>
__device__ __forceinline__ void store2_stream(T* base, const Lanes& L, T a, T b) {
  if (L.vec) {
    T2 v;
    v.x = a;
    v.y = b;
    __stcs(reinterpret_cast<T2*>(base + L.i0), v);
  } else {
    __stcs(base + L.i0, a);
    __stcs(base + L.i1, b);
  }
}

struct EnvReg {
  double s, K, r, ret, S, quota_const;
  int32_t years, done_step, model;
  bool frozen;
};

template <int PM>
__device__ __forceinline__ double policy_quota_escapement(const RolloutArgs& a, const EnvReg& e, double& action_out) {
  // escapement.predict, policies.py:27-31, then get_quota of the resulting action (the round trip is not exact)
  double x = __dadd_rn(e.s, 1.0);
  if (PM != P_KONE) x = __dmul_rn(x, e.K);
  const double q = py_max(__dsub_rn(x, e.S), 0.0);
  if (a.discrete) {
    const double act = action_discrete<PM>(q, a.n_actions, e.K, a.kd);
    action_out = act;
    const double t = div_z(act, a.n_actions);
    return PM == P_KONE ? t : __dmul_rn(t, e.K);
  }
  const double act = action_cts<PM>(q, e.K, a.kd);
  action_out = act;
  return quota_f64<PM>(act, e.K);
}

template <int PM>
__device__ __forceinline__ double policy_quota_msy(const RolloutArgs& a, double K, double& action_out) {
  // msy.predict, policies.py:16-19: action = get_action(msy) every step; constant per environment
  if (a.discrete) {
    const double act = action_discrete<PM>(a.msy_q, a.n_actions, K, a.kd);
    action_out = act;
    const double t = div_z(act, a.n_actions);
    return PM == P_KONE ? t : __dmul_rn(t, K);
  }
  const double act = action_cts<PM>(a.msy_q, K, a.kd);
  action_out = act;
  return quota_f64<PM>(act, K);
}

template <int PM>
__device__ __forceinline__ void load_action_pair(const RolloutArgs& a, const Lanes& L, int64_t t, double K0, double K1,
                                                 double& q0, double& q1, double& a0, double& a1) {
  const int64_t off = t * a.n;
  switch (a.action_dtype) {
    case FISHING_ACT_F64: {
      const double2 v = load2_stream<double2>(reinterpret_cast<const double*>(a.actions) + off, L);
      a0 = v.x; a1 = v.y;
      q0 = quota_f64<PM>(v.x, K0); q1 = quota_f64<PM>(v.y, K1);
      break;
    }
    case FISHING_ACT_F32: {
      const float2 v = load2_stream<float2>(reinterpret_cast<const float*>(a.actions) + off, L);
      a0 = v.x; a1 = v.y;
      q0 = quota_f32<PM>(v.x, K0); q1 = quota_f32<PM>(v.y, K1);
      break;
    }
    case FISHING_ACT_I32: {
      const int2 v = load2_stream<int2>(reinterpret_cast<const int32_t*>(a.actions) + off, L);
      a0 = v.x; a1 = v.y;
      q0 = quota_int<PM>(v.x, a.n_actions, a.rcp_n_actions, K0); q1 = quota_int<PM>(v.y, a.n_actions, a.rcp_n_actions, K1);
      break;
    }
    default: {
      const longlong2 v = load2_stream<longlong2>(reinterpret_cast<const long long*>(a.actions) + off, L);
      a0 = (double)v.x; a1 = (double)v.y;
      q0 = quota_int<PM>(v.x, a.n_actions, a.rcp_n_actions, K0); q1 = quota_int<PM>(v.y, a.n_actions, a.rcp_n_actions, K1);
      break;
    }
  }
}

template <int GROWTH, int PM>
__device__ __forceinline__ void env_advance(const RolloutArgs& a, const GrowthP& gp, EnvReg& e, double quota, double z,
                                            bool noisy, int64_t t, uint64_t gid, double& obs, double& reward,
                                            uint8_t& done) {
  const StepOut o = step_core<GROWTH, PM>(e.s, quota, z, e.r, e.K, a.C, a.sigma, noisy, &gp, &a.tab, e.model, a.kd);
  e.years += 1;
  done = (e.years > a.Tmax) || o.dead;
  reward = o.reward;
  // sequential float64 accumulation, same order as a Python loop; without auto-reset only the first episode counts
  if ((a.flags & FISHING_AUTO_RESET) || e.done_step < 0) e.ret = __dadd_rn(e.ret, o.reward);
  e.s = o.s2;
  obs = o.s2;
  if (done) {
    if (e.done_step < 0) e.done_step = (int32_t)t;
    if (a.flags & FISHING_AUTO_RESET) {
      e.years = 0;
      if (PM == P_PERENV) {
        const Normal2 d = normal_pair(a.seed, gid, a.step0 + (uint64_t)t, FISHING_STREAM_AUTORESET);
        e.K = v4_param(a.K_mean, a.sigma_p, d.z0);
        e.r = v4_param(a.r_mean, a.sigma_p, d.z1);
      }
      if (GROWTH == G_MODEL)
        e.model = (int32_t)uniform_index(a.seed, gid, a.step0 + (uint64_t)t, FISHING_STREAM_AUTORESET,
                                         (uint32_t)a.tab.n_models);
      e.s = a.s_reset;
      obs = a.s_reset;
    } else if (a.flags & FISHING_FREEZE_DONE) {
      e.frozen = true;
    }
  }
}

template <int GROWTH, int PM, int POLICY>
__global__ void __launch_bounds__(128) rollout_kernel(const RolloutArgs a) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (a.n >> 1)) return;
  Lanes L;
  uint64_t noise_pair;
  const int64_t m = a.noise_id_div;
  if (m > 1) {
    const int64_t kp = tid / m, j = tid - kp * m;
    L.i0 = (2 * kp) * m + j;
    L.i1 = L.i0 + m;
    L.vec = false;
    noise_pair = (uint64_t)((a.env_id0 / m) >> 1) + (uint64_t)kp;
  } else {
    L.i0 = 2 * tid;
    L.i1 = L.i0 + 1;
    L.vec = true;
    noise_pair = ((uint64_t)a.env_id0 >> 1) + (uint64_t)tid;
  }
  const uint64_t gid0 = (uint64_t)(a.env_id0 + L.i0), gid1 = (uint64_t)(a.env_id0 + L.i1);
  const bool ext_noise = a.noise != nullptr;
  const bool noisy = ext_noise || a.sigma != 0.0;
  const bool tracing = a.trace_obs || a.trace_reward || a.trace_done || a.trace_action;

  EnvReg e0, e1;
  {
    const double2 s = load2<double2>(a.state, L);
    const int2 y = load2<int2>(a.years, L);
    e0.s = s.x; e1.s = s.y;
    e0.years = y.x; e1.years = y.y;
    if (PM == P_PERENV) {
      const double2 K = load2<double2>(a.K_env, L), r = load2<double2>(a.r_env, L);
      e0.K = K.x; e1.K = K.y; e0.r = r.x; e1.r = r.y;
    } else {
      e0.K = e1.K = a.K;
      e0.r = e1.r = a.r;
    }
    e0.S = e1.S = a.S;
    if (POLICY == FISHING_POLICY_ESCAPEMENT && a.S_env) {
      const double2 S = load2<double2>(a.S_env, L);
      e0.S = S.x; e1.S = S.y;
    }
    e0.ret = e1.ret = 0.0;
    e0.done_step = e1.done_step = -1;
    e0.frozen = e1.frozen = false;
    e0.quota_const = e1.quota_const = 0.0;
    e0.model = e1.model = 0;
    if (GROWTH == G_MODEL) {
      const int2 md = load2<int2>(a.model, L);
      e0.model = md.x; e1.model = md.y;
    }
  }
  GrowthP gp = a.gp;   // fishing-v10 lets r drift: r += alpha before every draw (growth_models.py:147-150)
  double act0 = 0.0, act1 = 0.0;
  if (POLICY == FISHING_POLICY_MSY && PM != P_PERENV) {
    e0.quota_const = policy_quota_msy<PM>(a, e0.K, act0);
    e1.quota_const = e0.quota_const;
    act1 = act0;
  }

  for (int64_t t = 0; t < a.T; ++t) {
    if (e0.frozen && e1.frozen && !tracing) break;
    if (a.nonstationary) {
      gp.r = __dadd_rn(gp.r, a.alpha);
      growth_derive(gp, FISHING_FN_BEVERTON_HOLT);
    }
    double z0 = 0.0, z1 = 0.0;
    if (ext_noise) {
      const double2 z = load2_stream<double2>(a.noise + t * a.n, L);
      z0 = z.x; z1 = z.y;
    } else if (noisy) {
      const Normal2 d = normal_pair_keys(a.keys, noise_pair, a.step0 + (uint64_t)t, FISHING_STREAM_STEP);
      z0 = d.z0; z1 = d.z1;
    }
    double q0, q1;
    if (POLICY == FISHING_POLICY_ACTIONS) {
      load_action_pair<PM>(a, L, t, e0.K, e1.K, q0, q1, act0, act1);
    } else if (POLICY == FISHING_POLICY_ESCAPEMENT) {
      q0 = policy_quota_escapement<PM>(a, e0, act0);
      q1 = policy_quota_escapement<PM>(a, e1, act1);
    } else {
      if (PM == P_PERENV) {  // per-env K changes at auto-reset: recompute
        q0 = policy_quota_msy<PM>(a, e0.K, act0);
        q1 = policy_quota_msy<PM>(a, e1.K, act1);
      } else {
        q0 = e0.quota_const;
        q1 = e1.quota_const;
      }
    }
    double o0 = e0.s, o1 = e1.s, r0 = 0.0, r1 = 0.0;
    uint8_t d0 = 1, d1 = 1;
    if (!e0.frozen) env_advance<GROWTH, PM>(a, gp, e0, q0, z0, noisy, t, gid0, o0, r0, d0);
    if (!e1.frozen) env_advance<GROWTH, PM>(a, gp, e1, q1, z1, noisy, t, gid1, o1, r1, d1);
    if (tracing) {
      const int64_t off = t * a.n;
      if (a.trace_obs) store2_stream<double2>(a.trace_obs + off, L, o0, o1);
      if (a.trace_reward) store2_stream<double2>(a.trace_reward + off, L, r0, r1);
      if (a.trace_done) store2_stream<uchar2>(a.trace_done + off, L, d0, d1);
      if (a.trace_action) store2_stream<double2>(a.trace_action + off, L, act0, act1);
    }
  }

  store2<double2>(a.state, L, e0.s, e1.s);
  store2<int2>(a.years, L, e0.years, e1.years);
  if (PM == P_PERENV) {
    store2<double2>(a.K_env, L, e0.K, e1.K);
    store2<double2>(a.r_env, L, e0.r, e1.r);
  }
  if (GROWTH == G_MODEL) store2<int2>(a.model, L, e0.model, e1.model);
  if (a.out_return) store2<double2>(a.out_return, L, e0.ret, e1.ret);
  if (a.out_done_step) store2<int2>(a.out_done_step, L, e0.done_step, e1.done_step);
}

// ------------------------------------------------------------------------------------------------ K2, fast path
// The configuration the big rollouts use (Box action space, no traces, in-kernel noise or none) with everything that
// is loop-invariant resolved at compile time and a handful of comparisons removed that can never fire:
//   * max(x - h, 0.0) (base_fishing_env.py:118) is the identity because h = min(x, quota) <= x (or NaN) always;
//   * "stock <= 0" is the negation of the comparison np.maximum already makes;
//   * the lower clip of the action is dead for the escapement / MSY policies: a = q/K - 1 with q = max(., 0) >= 0;
//   * reward = max(h, 0.0) (base_fishing_env.py:74) is the identity when the stock is never negative: h = min(x, quota)
//     with quota >= 0 always, and x = (s + 1) K >= +-0 from the second step on (s = x2 / K - 1 with x2 >= 0, K >= 0).  Only
//     the state the caller hands in can hold a negative stock: such a thread runs its FIRST step with the comparison
//     (SAFE = true, a rarely taken copy of the loop body) and joins the others afterwards.  Negative K (P_UNIFORM with
//     kd.odd) keeps the comparison; a negative reset stock is routed to the generic kernel by launch_rollout_p.
//   * the return is accumulated unconditionally; when later episodes must not count (plain mode) its value at the first
//     `done` is kept aside, which costs one select less than gating the addition.
// Results are bit-identical to the generic kernel (tests/test_gpu_parity.py runs both against the oracle).
// A/B (tools/ab_build.py imad_select): float64 selects of the fast path done as b + m (a - b) on the integer halves, m = 0 / 1
// -- IADD / IMAD instead of two FSEL.  Measured 6 - 23 % slower (profiles/ab_r02.md): more instructions, and the ALU pipe was
// not the limit.
#ifdef FISHING_AB_IMAD_SELECT
__device__ __forceinline__ double sel64(bool p, double a, double b) {   // p ? a : b
  const unsigned m = p ? 1u : 0u;
  const unsigned alo = (unsigned)__double2loint(a), ahi = (unsigned)__double2hiint(a);
  const unsigned blo = (unsigned)__double2loint(b), bhi = (unsigned)__double2hiint(b);
  unsigned lo, hi;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(m), "r"(alo - blo), "r"(blo));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(m), "r"(ahi - bhi), "r"(bhi));
  return __hiloint2double((int)hi, (int)lo);
}
#else
__device__ __forceinline__ double sel64(bool p, double a, double b) { return p ? a : b; }
#endif

struct FastEnv {
  double s, K, r, S, ret, ret_snap, quota_const;
  int years, done_step;
  bool frozen;
  KDiv kdv;   // fishing-v4: the reciprocal of this environment's K (dynamics.cuh: kdiv_env), renewed when K is redrawn
};

// see "reward" above: does this environment enter with a negative stock?  (Then its first step runs with the comparison.)
template <int PM>
__device__ __forceinline__ bool fast_env_enter(FastEnv& e) {
  double x = __dadd_rn(e.s, 1.0);
  if (PM != P_KONE) x = __dmul_rn(x, e.K);
  e.ret = e.ret_snap = 0.0;
  e.done_step = -1;
  e.frozen = false;
  e.quota_const = 0.0;
  if (PM == P_PERENV && FISHING_V4_RCP) e.kdv = kdiv_env(e.K);
  return x < 0.0;
}

// What happens to an environment whose episode ended: it keeps stepping (the reference's behaviour when a caller ignores
// `done`), it is reset in place (VecEnv semantics), or it is frozen (return statistics of first episodes).  A template
// parameter, so that the plain loop carries no `frozen` tests and the compiler does not have to unswitch the loop.
enum : int { M_PLAIN = 0, M_AUTO_RESET = 1, M_FREEZE = 2 };

// FOLD (auto-reset, K = 1, checked by the host: RolloutArgs::fold_reset): the state reset rides on the np.maximum select -- the
// stock of a finished episode becomes x_reset = s_reset + 1, whose s = x - 1 is s_reset bit for bit -- instead of a second pair of
// selects after the step (two FSEL per environment-step less in the kernel behind bench.py's fused.escapement_in_kernel).
template <int GROWTH, int PM, int POLICY, bool NOISY, int MODE, bool SAFE = false, bool FOLD = false>
__device__ __forceinline__ void fast_env_step(const RolloutArgs& a, FastEnv& e, double action, double z, int t,
                                              uint64_t gid, const SharedTabs& tb, double& reward_out, bool& done_out) {
  constexpr bool auto_reset = MODE == M_AUTO_RESET, freeze = MODE == M_FREEZE;
  static_assert(!FOLD || (MODE == M_AUTO_RESET && PM == P_KONE), "the folded reset is for auto-reset with K = 1");
  double x = __dadd_rn(e.s, 1.0);
  if (PM != P_KONE) x = __dmul_rn(x, e.K);
  double quota;
  if (POLICY == FISHING_POLICY_ESCAPEMENT) {
    const double xs = __dsub_rn(x, e.S);
    const double q = sel64(0.0 > xs, 0.0, xs);                   // max(x - S, 0.0), policies.py:29
    double act = action_cts<PM>(q, e.K, a.kd);                   // get_action, base_fishing_env.py:156
    act = sel64(act > 1.0, 1.0, act);                            // np.clip: the lower bound cannot fire (act >= -1)
    quota = __dadd_rn(act, 1.0);
    if (PM != P_KONE) quota = __dmul_rn(quota, e.K);
  } else if (POLICY == FISHING_POLICY_MSY) {
    quota = e.quota_const;
  } else {
    quota = quota_f64<PM>(action, e.K);
  }
  const double h = sel64(quota < x, quota, x);                   // min(x, quota), harvest_draw
  const double x1 = __dsub_rn(x, h);                             // == max(x - h, 0.0)
  const KDiv& kdq = (PM == P_PERENV && FISHING_V4_RCP) ? e.kdv : a.kd;      // whose reciprocal the step's two quotients use
  const double frac = div_pm<PM>(x1, e.K, kdq);
  const double room = __dsub_rn(1.0, frac);
  double g;
  if (GROWTH == LOGISTIC) {
    g = __dadd_rn(x1, __dmul_rn(__dmul_rn(e.r, x1), room));
    if (NOISY) g = __dadd_rn(g, __dmul_rn(__dmul_rn(x1, a.sigma), z));
    else if (PM == P_PERENV) g = __dadd_rn(g, __dmul_rn(x1, 0.0));   // see grow(): sigma == 0 (odd uniform parameters: generic kernel)
  } else {
    double ex = __dmul_rn(__dmul_rn(e.r, room), __dsub_rn(x1, a.C));
    if (NOISY) ex = __dadd_rn(ex, __dmul_rn(__dmul_rn(x1, a.sigma), z));
    else if (PM == P_PERENV) ex = __dadd_rn(ex, __dmul_rn(x1, 0.0));
    g = __dmul_rn(x1, exp_tab(ex, tb));
  }
  const bool alive = (g > 0.0) || (g != g);                      // np.maximum(g, 0): NaN propagates
  bool done = false;
  double x2;
  if (FOLD) {   // not done implies alive, so (done ? x_reset : g) is np.maximum's result wherever the episode goes on
    e.years += 1;
    done = (e.years > a.Tmax) || !alive;
    x2 = sel64(done, a.x_reset_fold, g);
  } else {
    x2 = sel64(alive, g, 0.0);
  }
  const double s2 = __dsub_rn(div_pm<PM>(x2, e.K, kdq), 1.0);
  const double reward = (SAFE || (PM == P_UNIFORM && a.kd.odd)) ? py_max(h, 0.0) : h;   // see the comment on FastEnv
  if (!FOLD) {  // the source order the other instantiations were tuned with (register allocation: tests/test_cpu_host.py)
    e.years += 1;
    done = (e.years > a.Tmax) || !alive;
  }
  reward_out = reward;
  done_out = done;
  e.ret = __dadd_rn(e.ret, reward);
  e.s = s2;
  if (done) {
    if (e.done_step < 0) {
      e.done_step = t;
      if (!auto_reset && !freeze) e.ret_snap = e.ret;
    }
    if (auto_reset) {
      e.years = 0;
      if (PM == P_PERENV) {
        const Normal2 d = normal_pair(a.seed, gid, a.step0 + (uint64_t)t, FISHING_STREAM_AUTORESET);
        e.K = v4_param(a.K_mean, a.sigma_p, d.z0);
        e.r = v4_param(a.r_mean, a.sigma_p, d.z1);
        if (FISHING_V4_RCP) e.kdv = kdiv_env(e.K);
        if (POLICY == FISHING_POLICY_MSY) {
          double unused;
          e.quota_const = policy_quota_msy<PM>(a, e.K, unused);
        }
      }
      if (!FOLD) e.s = a.s_reset;
    } else if (freeze) {
      e.frozen = true;
    }
  }
}

template <int GROWTH, int PM, int POLICY, bool NOISY, int MODE, bool SAFE = false, bool FOLD = false>
__device__ __forceinline__ void fast_env_step(const RolloutArgs& a, FastEnv& e, double action, double z, int t,
                                              uint64_t gid, const SharedTabs& tb) {
  double reward;
  bool done;
  fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE, FOLD>(a, e, action, z, t, gid, tb, reward, done);
}

// the return to report: all episodes (auto-reset), or the first one (frozen environments stop accumulating by themselves)
template <int MODE>
__device__ __forceinline__ double fast_env_return(const FastEnv& e) {
  return (MODE == M_PLAIN && e.done_step >= 0) ? e.ret_snap : e.ret;
}

// TRACE: per-step observation / reward / done rows [T, n] written with streaming stores (the rollout-buffer fill of a
// whole chunk in one launch).  Not combined with freeze-on-done or an action trace; those go to the generic kernel.
template <int GROWTH, int PM, int POLICY, bool NOISY, int MODE, bool TRACE = false, bool FOLD = false>
__global__ void __launch_bounds__(fast_bs(GROWTH), fast_minblocks(GROWTH)) rollout_fast_kernel(const RolloutArgs a) {
  // the Box-Muller tables, staged once per block for the T draws of every thread (normal.cuh)
  constexpr bool TABLES = NOISY || GROWTH == TIPPING;     // Box-Muller tables, and the exp table of the tipping-point growth
  __shared__ double2 s_tab[TABLES ? NORMAL_TABLE_ENTRIES : 1];
  SharedTabs tb = {};
  if (TABLES) tb = stage_normal_tables<fast_bs(GROWTH)>(s_tab);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (a.n >> 1)) return;
  Lanes L;
  uint64_t noise_pair;
  const int64_t m = a.noise_id_div;
  if (m > 1) {
    const int64_t kp = tid / m, j = tid - kp * m;
    L.i0 = (2 * kp) * m + j;
    L.i1 = L.i0 + m;
    L.vec = false;
    noise_pair = (uint64_t)((a.env_id0 / m) >> 1) + (uint64_t)kp;
  } else {
    L.i0 = 2 * tid;
    L.i1 = L.i0 + 1;
    L.vec = true;
    noise_pair = ((uint64_t)a.env_id0 >> 1) + (uint64_t)tid;
  }
  const uint64_t gid0 = (uint64_t)(a.env_id0 + L.i0), gid1 = (uint64_t)(a.env_id0 + L.i1);

  FastEnv e0, e1;
  bool unsafe;
  {
    const double2 s = load2<double2>(a.state, L);
    const int2 y = load2<int2>(a.years, L);
    e0.s = s.x; e1.s = s.y;
    e0.years = y.x; e1.years = y.y;
    e0.K = e1.K = a.K;
    e0.r = e1.r = a.r;
    if (PM == P_PERENV) {
      const double2 K = load2<double2>(a.K_env, L), r = load2<double2>(a.r_env, L);
      e0.K = K.x; e1.K = K.y; e0.r = r.x; e1.r = r.y;
    }
    e0.S = e1.S = a.S;
    if (POLICY == FISHING_POLICY_ESCAPEMENT && a.S_env) {
      const double2 S = load2<double2>(a.S_env, L);
      e0.S = S.x; e1.S = S.y;
    }
    unsafe = fast_env_enter<PM>(e0);
    unsafe = fast_env_enter<PM>(e1) || unsafe;
    if (POLICY == FISHING_POLICY_MSY) {
      double unused;
      e0.quota_const = policy_quota_msy<PM>(a, e0.K, unused);
      e1.quota_const = policy_quota_msy<PM>(a, e1.K, unused);
    }
  }
  const int T = (int)a.T;
  const double* act_base = reinterpret_cast<const double*>(a.actions);
  double2 act_next = make_double2(0.0, 0.0);
  if (POLICY == FISHING_POLICY_ACTIONS && T > 0) act_next = load2_stream<double2>(act_base, L);
  const int64_t row_f64 = a.n * 8, row_u8 = a.n;   // one trace row, in bytes
  char* tr_obs = TRACE ? reinterpret_cast<char*>(a.trace_obs) + tid * 16 : nullptr;
  char* tr_rew = TRACE ? reinterpret_cast<char*>(a.trace_reward) + tid * 16 : nullptr;
  char* tr_done = TRACE ? reinterpret_cast<char*>(a.trace_done) + tid * 2 : nullptr;

  // one lock-step of the thread's two environments; returns true when the loop can stop (both frozen)
  auto step = [&](int t, auto safe_tag) -> bool {
    constexpr bool SAFE = decltype(safe_tag)::value;
    double2 act = act_next;
    if (POLICY == FISHING_POLICY_ACTIONS && t + 1 < T) act_next = load2_stream<double2>(act_base + (int64_t)(t + 1) * a.n, L);
    double z0 = 0.0, z1 = 0.0;
    if (NOISY) {
      const Normal2 d = normal_pair_keys(a.keys, noise_pair, a.step0 + (uint64_t)t, FISHING_STREAM_STEP, tb);
      z0 = d.z0; z1 = d.z1;
    }
    if (MODE == M_FREEZE) {
      if (!e0.frozen) fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE>(a, e0, act.x, z0, t, gid0, tb);
      if (!e1.frozen) fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE>(a, e1, act.y, z1, t, gid1, tb);
      return e0.frozen && e1.frozen;
    } else if (TRACE) {
      double r0, r1;
      bool d0, d1;
      fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE>(a, e0, act.x, z0, t, gid0, tb, r0, d0);
      fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE>(a, e1, act.y, z1, t, gid1, tb, r1, d1);
      // adjacent-pair layout and all three rows present (launch_rollout_p checks both): one 16-byte / 16-byte / 2-byte
      // streaming store each through pointers that advance by one row per step
      __stcs(reinterpret_cast<double2*>(tr_obs), make_double2(e0.s, e1.s));
      __stcs(reinterpret_cast<double2*>(tr_rew), make_double2(r0, r1));
      __stcs(reinterpret_cast<uchar2*>(tr_done), make_uchar2((uint8_t)d0, (uint8_t)d1));
      tr_obs += row_f64;      // byte strides computed once: plain 64-bit adds, no multiply in the loop
      tr_rew += row_f64;
      tr_done += row_u8;
    } else {
      fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE, FOLD>(a, e0, act.x, z0, t, gid0, tb);
      fast_env_step<GROWTH, PM, POLICY, NOISY, MODE, SAFE, FOLD>(a, e1, act.y, z1, t, gid1, tb);
    }
    return false;
  };
  // the first step is peeled for every thread (so that the loop counter stays warp-uniform); a thread into which a negative
  // stock came runs it with reward = max(h, 0.0)
  bool stop = T <= 0;
  if (!stop) stop = unsafe ? step(0, std::true_type()) : step(0, std::false_type());
  if (!stop) {
#pragma unroll K2_UNROLL
    for (int t = 1; t < T; ++t)
      if (step(t, std::false_type())) break;
  }

  store2<double2>(a.state, L, e0.s, e1.s);
  store2<int2>(a.years, L, e0.years, e1.years);
  if (PM == P_PERENV) {
    store2<double2>(a.K_env, L, e0.K, e1.K);
    store2<double2>(a.r_env, L, e0.r, e1.r);
  }
  if (a.out_return) store2<double2>(a.out_return, L, fast_env_return<MODE>(e0), fast_env_return<MODE>(e1));
  if (a.out_done_step) store2<int2>(a.out_done_step, L, e0.done_step, e1.done_step);
}

// ------------------------------------------------------------------------------------------------ K2, sweep variant
// Escapement grid sweep with common random numbers (noise_id_div = m = number of policy levels, env id =
// replicate * m + level): the noise depends only on (replicate, step), so one thread that owns a replicate PAIR and LV
// levels generates the pair's two normals once per step and applies them to 2*LV environments -- the Philox +
// Box-Muller cost, more than half of the instructions of the plain kernel, is amortised LV times.  The levels of a thread
// are strided by m/LV so that neighbouring threads touch neighbouring addresses.
template <int GROWTH, int PM, bool NOISY, int LV, int MODE>
__global__ void __launch_bounds__(128) rollout_sweep_kernel(const RolloutArgs a) {
  constexpr bool TABLES = NOISY || GROWTH == TIPPING;     // Box-Muller tables, and the exp table of the tipping-point growth
  __shared__ double2 s_tab[TABLES ? NORMAL_TABLE_ENTRIES : 1];
  SharedTabs tb = {};
  if (TABLES) tb = stage_normal_tables<128>(s_tab);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = a.noise_id_div, mq = m / LV;
  if (tid >= (a.n >> 1) / LV) return;
  const int64_t kp = tid / mq, jq = tid - kp * mq;
  const uint64_t noise_pair = (uint64_t)((a.env_id0 / m) >> 1) + (uint64_t)kp;
  FastEnv e[2 * LV];
  int64_t idx[2 * LV];
  bool unsafe = false;
#pragma unroll
  for (int l = 0; l < LV; ++l)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t i = (2 * kp + h) * m + jq + l * mq;
      FastEnv& v = e[2 * l + h];
      idx[2 * l + h] = i;
      v.s = a.state[i];
      v.years = a.years[i];
      v.K = PM == P_PERENV ? a.K_env[i] : a.K;
      v.r = PM == P_PERENV ? a.r_env[i] : a.r;
      v.S = a.S_env ? a.S_env[i] : a.S;
      unsafe = fast_env_enter<PM>(v) || unsafe;
    }
  const int T = (int)a.T;
  auto step = [&](int t, auto safe_tag) -> bool {
    constexpr bool SAFE = decltype(safe_tag)::value;
    double z[2] = {0.0, 0.0};
    if (NOISY) {
      const Normal2 d = normal_pair_keys(a.keys, noise_pair, a.step0 + (uint64_t)t, FISHING_STREAM_STEP, tb);
      z[0] = d.z0;
      z[1] = d.z1;
    }
    bool all_frozen = MODE == M_FREEZE;
#pragma unroll
    for (int k = 0; k < 2 * LV; ++k) {
      if (MODE != M_FREEZE || !e[k].frozen)
        fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE, SAFE>(a, e[k], 0.0, z[k & 1], t,
                                                                                 (uint64_t)(a.env_id0 + idx[k]), tb);
      all_frozen = all_frozen && e[k].frozen;
    }
    return all_frozen;
  };
  bool stop = T <= 0;   // first step peeled, see rollout_fast_kernel
  if (!stop) stop = unsafe ? step(0, std::true_type()) : step(0, std::false_type());
  if (!stop) {
#pragma unroll 1
    for (int t = 1; t < T; ++t)
      if (step(t, std::false_type())) break;
  }
#pragma unroll
  for (int k = 0; k < 2 * LV; ++k) {
    const int64_t i = idx[k];
    a.state[i] = e[k].s;
    a.years[i] = e[k].years;
    if (PM == P_PERENV) {
      a.K_env[i] = e[k].K;
      a.r_env[i] = e[k].r;
    }
    if (a.out_return) a.out_return[i] = fast_env_return<MODE>(e[k]);
    if (a.out_done_step) a.out_done_step[i] = e[k].done_step;
  }
}

// FISHING_K2_FOLD=0 in the environment: the auto-reset fast path overwrites the state after the step (A/B, tests compare both)
static bool fold_enabled() {
  static const bool on = [] {
    const char* e = getenv("FISHING_K2_FOLD");
    return !(e && e[0] == '0');
  }();
  return on;
}
static int done_mode(const RolloutArgs& a) {
  return (a.flags & FISHING_AUTO_RESET) ? M_AUTO_RESET : (a.flags & FISHING_FREEZE_DONE) ? M_FREEZE : M_PLAIN;
}

template <int GROWTH, int PM, int POLICY, bool NOISY>
static void launch_fast_m(const RolloutArgs& a, int grid, int block, cudaStream_t st) {
  if (a.trace_obs || a.trace_reward || a.trace_done) {
    if constexpr (POLICY != FISHING_POLICY_MSY) {   // launch_rollout_p sends MSY + traces and freeze + traces elsewhere
      if (done_mode(a) == M_AUTO_RESET) rollout_fast_kernel<GROWTH, PM, POLICY, NOISY, M_AUTO_RESET, true><<<grid, block, 0, st>>>(a);
      else rollout_fast_kernel<GROWTH, PM, POLICY, NOISY, M_PLAIN, true><<<grid, block, 0, st>>>(a);
    }
    return;
  }
  switch (done_mode(a)) {
    case M_AUTO_RESET:
      if constexpr (PM == P_KONE) {
        if (a.fold_reset && fold_enabled()) {
          rollout_fast_kernel<GROWTH, PM, POLICY, NOISY, M_AUTO_RESET, false, true><<<grid, block, 0, st>>>(a);
          break;
        }
      }
      rollout_fast_kernel<GROWTH, PM, POLICY, NOISY, M_AUTO_RESET><<<grid, block, 0, st>>>(a);
      break;
    case M_FREEZE: rollout_fast_kernel<GROWTH, PM, POLICY, NOISY, M_FREEZE><<<grid, block, 0, st>>>(a); break;
    default: rollout_fast_kernel<GROWTH, PM, POLICY, NOISY, M_PLAIN><<<grid, block, 0, st>>>(a); break;
  }
}

// ------------------------------------------------------------------------------------------------ K2, warp-cooperative sweep
// Same sweep layout, amortised across the WARP instead of across a thread's registers: a warp owns one replicate pair and
// 32 consecutive levels (lane = level), so all 32 lanes need the same two normals at every step.  Lane l generates the
// draw of step t0 + l; the 32 draws are then handed round with shuffles while the warp walks through steps t0 .. t0+31.
// One Philox block + Box-Muller per thread every 32 steps instead of every step (or every step per 4 levels): the loop
// body is the noise-free dynamics plus four SHFL, at the register count -- and occupancy -- of the plain kernel.
template <int GROWTH, int PM, bool NOISY, int MODE>
__global__ void __launch_bounds__(128) rollout_sweep_warp_kernel(const RolloutArgs a) {
  constexpr bool TABLES = NOISY || GROWTH == TIPPING;     // Box-Muller tables, and the exp table of the tipping-point growth
  __shared__ double2 s_tab[TABLES ? NORMAL_TABLE_ENTRIES : 1];
  SharedTabs tb = {};
  if (TABLES) tb = stage_normal_tables<128>(s_tab);
  const unsigned full = 0xffffffffu;
  const int64_t m = a.noise_id_div;                      // levels, a multiple of 32
  const int64_t warps_per_pair = m >> 5;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= ((a.n / m) >> 1) * warps_per_pair) return;  // whole warps only
  const int64_t kp = gw / warps_per_pair;
  const int64_t j = (gw - kp * warps_per_pair) * 32 + lane;
  const int64_t i0 = (2 * kp) * m + j, i1 = i0 + m;
  const uint64_t noise_pair = (uint64_t)((a.env_id0 / m) >> 1) + (uint64_t)kp;
  const uint64_t gid0 = (uint64_t)(a.env_id0 + i0), gid1 = (uint64_t)(a.env_id0 + i1);
  FastEnv e0, e1;
  e0.s = a.state[i0]; e1.s = a.state[i1];
  e0.years = a.years[i0]; e1.years = a.years[i1];
  e0.K = PM == P_PERENV ? a.K_env[i0] : a.K; e1.K = PM == P_PERENV ? a.K_env[i1] : a.K;
  e0.r = PM == P_PERENV ? a.r_env[i0] : a.r; e1.r = PM == P_PERENV ? a.r_env[i1] : a.r;
  e0.S = a.S_env ? a.S_env[i0] : a.S; e1.S = a.S_env ? a.S_env[i1] : a.S;
  bool unsafe = fast_env_enter<PM>(e0);
  unsafe = fast_env_enter<PM>(e1) || unsafe;
  const int T = (int)a.T;
  bool all_done = false;
  if (T > 0) {   // first step peeled (see rollout_fast_kernel); with the comparison when a negative stock came into the warp
    Normal2 d;
    d.z0 = d.z1 = 0.0;
    if (NOISY) d = normal_pair_keys(a.keys, noise_pair, a.step0, FISHING_STREAM_STEP, tb);
    if (__any_sync(full, unsafe)) {
      fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE, true>(a, e0, 0.0, d.z0, 0, gid0, tb);
      fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE, true>(a, e1, 0.0, d.z1, 0, gid1, tb);
    } else {
      fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE>(a, e0, 0.0, d.z0, 0, gid0, tb);
      fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE>(a, e1, 0.0, d.z1, 0, gid1, tb);
    }
    all_done = MODE == M_FREEZE && __all_sync(full, e0.frozen && e1.frozen);
  }
#pragma unroll 1
  for (int t0 = 1; t0 < T && !all_done; t0 += 32) {
    Normal2 d;
    d.z0 = d.z1 = 0.0;
    if (NOISY && t0 + lane < T)
      d = normal_pair_keys(a.keys, noise_pair, a.step0 + (uint64_t)(t0 + lane), FISHING_STREAM_STEP, tb);
    const int nt = (T - t0 < 32) ? T - t0 : 32;
#pragma unroll SWEEP_UNROLL
    for (int dt = 0; dt < nt; ++dt) {
      double z0 = 0.0, z1 = 0.0;
      if (NOISY) {
        z0 = __shfl_sync(full, d.z0, dt);
        z1 = __shfl_sync(full, d.z1, dt);
      }
      const int t = t0 + dt;
      if (MODE != M_FREEZE || !e0.frozen)
        fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE>(a, e0, 0.0, z0, t, gid0, tb);
      if (MODE != M_FREEZE || !e1.frozen)
        fast_env_step<GROWTH, PM, FISHING_POLICY_ESCAPEMENT, NOISY, MODE>(a, e1, 0.0, z1, t, gid1, tb);
    }
    // warp-uniform exit (the shuffles need every lane): looked at once per block of 32 steps -- frozen environments skip
    // their steps anyway, the vote only saves the rest of the walk when the whole warp is extinct
    if (MODE == M_FREEZE && __all_sync(full, e0.frozen && e1.frozen)) all_done = true;
  }
  a.state[i0] = e0.s; a.state[i1] = e1.s;
  a.years[i0] = e0.years; a.years[i1] = e1.years;
  if (PM == P_PERENV) {
    a.K_env[i0] = e0.K; a.K_env[i1] = e1.K;
    a.r_env[i0] = e0.r; a.r_env[i1] = e1.r;
  }
  if (a.out_return) { a.out_return[i0] = fast_env_return<MODE>(e0); a.out_return[i1] = fast_env_return<MODE>(e1); }
  if (a.out_done_step) { a.out_done_step[i0] = e0.done_step; a.out_done_step[i1] = e1.done_step; }
}

template <int GROWTH, int PM, bool NOISY>
static void launch_sweep_warp_m(const RolloutArgs& a, int grid, int block, cudaStream_t st) {
  switch ((a.flags & FISHING_AUTO_RESET) ? M_AUTO_RESET : (a.flags & FISHING_FREEZE_DONE) ? M_FREEZE : M_PLAIN) {
    case M_AUTO_RESET: rollout_sweep_warp_kernel<GROWTH, PM, NOISY, M_AUTO_RESET><<<grid, block, 0, st>>>(a); break;
    case M_FREEZE: rollout_sweep_warp_kernel<GROWTH, PM, NOISY, M_FREEZE><<<grid, block, 0, st>>>(a); break;
    default: rollout_sweep_warp_kernel<GROWTH, PM, NOISY, M_PLAIN><<<grid, block, 0, st>>>(a); break;
  }
}

template <int GROWTH, int PM, int POLICY>
static void launch_fast_n(const RolloutArgs& a, bool noisy, int grid, int block, cudaStream_t st) {
  if (noisy) launch_fast_m<GROWTH, PM, POLICY, true>(a, grid, block, st);
  else launch_fast_m<GROWTH, PM, POLICY, false>(a, grid, block, st);
}

template <int GROWTH, int PM, bool NOISY, int LV>
static void launch_sweep_m(const RolloutArgs& a, int grid, int block, cudaStream_t st) {
  switch (done_mode(a)) {
    case M_AUTO_RESET: rollout_sweep_kernel<GROWTH, PM, NOISY, LV, M_AUTO_RESET><<<grid, block, 0, st>>>(a); break;
    case M_FREEZE: rollout_sweep_kernel<GROWTH, PM, NOISY, LV, M_FREEZE><<<grid, block, 0, st>>>(a); break;
    default: rollout_sweep_kernel<GROWTH, PM, NOISY, LV, M_PLAIN><<<grid, block, 0, st>>>(a); break;
  }
}

template <int GROWTH, int PM>
static cudaError_t launch_rollout_p(const RolloutArgs& a, int policy, cudaStream_t st) {
  const int64_t threads = a.n >> 1;
  const int block = 128;
  const int64_t grid = (threads + block - 1) / block;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  const bool tracing = a.trace_obs || a.trace_reward || a.trace_done || a.trace_action;
  // traces ride on the specialised kernel unless they include the action trace, the done-handling is freeze-on-done, the
  // policy is MSY, or the environments are laid out as a sweep
  const bool fast_trace_ok = !a.trace_action && a.trace_obs && a.trace_reward && a.trace_done && done_mode(a) != M_FREEZE &&
                             policy != FISHING_POLICY_MSY && a.noise_id_div <= 1;
  // noise-free runs with uniform parameters outside (K > 0, r >= 0) need the explicit x * 0 term of grow(): generic kernel
  const bool odd_noise_free = a.kd.odd && a.sigma == 0.0;
  // auto-reset into a negative stock: the specialised kernels assume the reset state holds a stock >= 0 (see FastEnv)
  const double x_reset = PM == P_UNIFORM ? (a.s_reset + 1.0) * a.K : a.s_reset + 1.0;
  const bool neg_reset = (a.flags & FISHING_AUTO_RESET) && x_reset < 0.0;
  const bool fast = !a.discrete && (!tracing || fast_trace_ok) && !a.noise && !odd_noise_free && !neg_reset && a.T <= 0x7fffffffLL &&
                    (policy != FISHING_POLICY_ACTIONS || a.action_dtype == FISHING_ACT_F64) && !generic_only();
  constexpr int SWEEP_LV = 4;
  const bool sweep = fast && policy == FISHING_POLICY_ESCAPEMENT && a.noise_id_div > 1 && sweep_mode() != 1;
  if (sweep && a.noise_id_div % 32 == 0 && sweep_mode() == 0) {
    const int64_t wgrid = (threads + block - 1) / block;          // one thread per (replicate pair, level), whole warps
    if (a.sigma != 0.0) launch_sweep_warp_m<GROWTH, PM, true>(a, (int)wgrid, block, st);
    else launch_sweep_warp_m<GROWTH, PM, false>(a, (int)wgrid, block, st);
  } else if (sweep && a.noise_id_div % SWEEP_LV == 0) {
    const int64_t sgrid = (threads / SWEEP_LV + block - 1) / block;
    if (a.sigma != 0.0) launch_sweep_m<GROWTH, PM, true, SWEEP_LV>(a, (int)sgrid, block, st);
    else launch_sweep_m<GROWTH, PM, false, SWEEP_LV>(a, (int)sgrid, block, st);
  } else if (fast) {
    const bool noisy = a.sigma != 0.0;
    constexpr int FBS = fast_bs(GROWTH);
    const int fgrid = (int)((threads + FBS - 1) / FBS);
    switch (policy) {
      case FISHING_POLICY_ACTIONS: launch_fast_n<GROWTH, PM, FISHING_POLICY_ACTIONS>(a, noisy, fgrid, FBS, st); break;
      case FISHING_POLICY_ESCAPEMENT: launch_fast_n<GROWTH, PM, FISHING_POLICY_ESCAPEMENT>(a, noisy, fgrid, FBS, st); break;
      default: launch_fast_n<GROWTH, PM, FISHING_POLICY_MSY>(a, noisy, fgrid, FBS, st); break;
    }
  } else {
    switch (policy) {
      case FISHING_POLICY_ACTIONS: rollout_kernel<GROWTH, PM, FISHING_POLICY_ACTIONS><<<(int)grid, block, 0, st>>>(a); break;
      case FISHING_POLICY_ESCAPEMENT: rollout_kernel<GROWTH, PM, FISHING_POLICY_ESCAPEMENT><<<(int)grid, block, 0, st>>>(a); break;
      default: rollout_kernel<GROWTH, PM, FISHING_POLICY_MSY><<<(int)grid, block, 0, st>>>(a); break;
    }
  }
  count_launch();
  return cudaGetLastError();
}

template <int GROWTH>
static cudaError_t launch_rollout_ln(const RolloutArgs& a, int policy, cudaStream_t st) {
  const int64_t threads = a.n >> 1;
  const int block = 128;
  const int64_t grid = (threads + block - 1) / block;
  if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  switch (policy) {
    case FISHING_POLICY_ACTIONS: rollout_kernel<GROWTH, P_UNIFORM, FISHING_POLICY_ACTIONS><<<(int)grid, block, 0, st>>>(a); break;
    case FISHING_POLICY_ESCAPEMENT: rollout_kernel<GROWTH, P_UNIFORM, FISHING_POLICY_ESCAPEMENT><<<(int)grid, block, 0, st>>>(a); break;
    default: rollout_kernel<GROWTH, P_UNIFORM, FISHING_POLICY_MSY><<<(int)grid, block, 0, st>>>(a); break;
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rollout(const RolloutArgs& a, int kind, int policy, cudaStream_t st) {
  switch (kind) {
    case FISHING_V5: return launch_rollout_ln<G_ALLEN>(a, policy, st);
    case FISHING_V6: case FISHING_V10: return launch_rollout_ln<G_BEVERTON_HOLT>(a, policy, st);
    case FISHING_V7: return launch_rollout_ln<G_MAY>(a, policy, st);
    case FISHING_V8: return launch_rollout_ln<G_MYERS>(a, policy, st);
    case FISHING_V9: return launch_rollout_ln<G_RICKER>(a, policy, st);
    case FISHING_V11: return launch_rollout_ln<G_MODEL>(a, policy, st);
    default: break;
  }
  if (kind == FISHING_V4) return launch_rollout_p<LOGISTIC, P_PERENV>(a, policy, st);
  const bool kone = a.K == 1.0;
  if (kind == FISHING_V2)
    return kone ? launch_rollout_p<TIPPING, P_KONE>(a, policy, st) : launch_rollout_p<TIPPING, P_UNIFORM>(a, policy, st);
  return kone ? launch_rollout_p<LOGISTIC, P_KONE>(a, policy, st) : launch_rollout_p<LOGISTIC, P_UNIFORM>(a, policy, st);
}

// ------------------------------------------------------------------------------------------------ K4: statistics
// View ret[n] as a matrix [rows = n / groups, groups].  Pass 1: each thread owns one column of a 64-row chunk and adds
// the rows in order -> partial[chunk][group][3].  Pass 2: pairwise (binary-tree) combination of the chunk partials in
// a fixed order.  No atomics, so the result depends only on (n, groups), never on the grid or on timing; contiguous
// power-of-two shards of the row range are subtrees of the same tree (multi-GPU results combine exactly).
constexpr int REDUCE_ROWS = 64;

__global__ void __launch_bounds__(256) reduce_pass1(const double* __restrict__ ret, const int32_t* __restrict__ done_step,
                                                     int64_t rows, int64_t groups, int64_t chunks, double* __restrict__ partial) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  // blockIdx.y strides over the chunks (gridDim.y is capped at 65535; more than 4.19 million rows need several rounds)
  for (int64_t chunk = blockIdx.y; chunk < chunks; chunk += gridDim.y) {
    const int64_t r0 = chunk * REDUCE_ROWS;
    const int64_t r1 = (r0 + REDUCE_ROWS < rows) ? r0 + REDUCE_ROWS : rows;
    double s = 0.0, s2 = 0.0, nd = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
      const double v = ret[r * groups + g];
      s = __dadd_rn(s, v);
      s2 = __dadd_rn(s2, __dmul_rn(v, v));
      if (done_step && done_step[r * groups + g] >= 0) nd += 1.0;
    }
    double* p = partial + (chunk * groups + g) * 3;
    p[0] = s;
    p[1] = s2;
    p[2] = nd;
  }
}

__global__ void __launch_bounds__(256) reduce_pass2(double* __restrict__ partial, int64_t chunks, int64_t groups,
                                                     double* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  // adjacent pairing (2c, 2c+1) -> c, level by level, in place: p[c] is consumed at iteration c/2 <= c before it is
  // overwritten.  An odd leftover is carried up unchanged (x + nothing), so aligned power-of-two sub-ranges are
  // exact subtrees.
  while (chunks > 1) {
    const int64_t next = (chunks + 1) >> 1;
    for (int64_t c = 0; c < next; ++c) {
      double* d = partial + (c * groups + g) * 3;
      const double* p = partial + ((2 * c) * groups + g) * 3;
      if (2 * c + 1 < chunks) {
        const double* q = partial + ((2 * c + 1) * groups + g) * 3;
        const double v0 = __dadd_rn(p[0], q[0]), v1 = __dadd_rn(p[1], q[1]), v2 = __dadd_rn(p[2], q[2]);
        d[0] = v0; d[1] = v1; d[2] = v2;
      } else {
        const double v0 = p[0], v1 = p[1], v2 = p[2];
        d[0] = v0; d[1] = v1; d[2] = v2;
      }
    }
    chunks = next;
  }
  out[g * 3 + 0] = partial[g * 3 + 0];
  out[g * 3 + 1] = partial[g * 3 + 1];
  out[g * 3 + 2] = partial[g * 3 + 2];
}

// groups == 1 convenience: the columns of the internal 256-wide view are folded by one block with the same adjacent
// pairwise tree.
__global__ void __launch_bounds__(256) reduce_fold_columns(const double* __restrict__ cols, int ncols, double* __restrict__ out) {
  __shared__ double sh[256 * 3];
  const int t = threadIdx.x;
  for (int k = 0; k < 3; ++k) sh[t * 3 + k] = t < ncols ? cols[t * 3 + k] : 0.0;
  __syncthreads();
  for (int width = 256; width > 1; width >>= 1) {
    double v[3];
    const bool act = t < width / 2;
    if (act)
      for (int k = 0; k < 3; ++k) v[k] = __dadd_rn(sh[(2 * t) * 3 + k], sh[(2 * t + 1) * 3 + k]);
    __syncthreads();
    if (act)
      for (int k = 0; k < 3; ++k) sh[t * 3 + k] = v[k];
    __syncthreads();
  }
  if (t < 3) out[t] = sh[t];
}

constexpr int VIRTUAL_GROUPS = 256;
static int64_t virtual_groups(int64_t n, int64_t groups) {
  return (groups == 1 && n % VIRTUAL_GROUPS == 0 && n >= VIRTUAL_GROUPS) ? VIRTUAL_GROUPS : groups;
}

int64_t reduce_chunks(int64_t n, int64_t groups) {
  const int64_t rows = n / virtual_groups(n, groups);
  return (rows + REDUCE_ROWS - 1) / REDUCE_ROWS;
}

int64_t reduce_scratch_doubles(int64_t n, int64_t groups) {
  const int64_t vg = virtual_groups(n, groups);
  return (reduce_chunks(n, groups) + 1) * vg * 3;
}

cudaError_t launch_reduce(const double* ret, const int32_t* done_step, int64_t n, int64_t groups, double* scratch,
                          double* out, cudaStream_t st) {
  const int64_t vg = virtual_groups(n, groups);
  const int64_t rows = n / vg;
  const int64_t chunks = reduce_chunks(n, groups);
  dim3 grid((unsigned)((vg + 255) / 256), (unsigned)(chunks < 65535 ? chunks : 65535));
  reduce_pass1<<<grid, 256, 0, st>>>(ret, done_step, rows, vg, chunks, scratch);
  count_launch();
  double* cols = vg == groups ? out : scratch + chunks * vg * 3;
  reduce_pass2<<<(unsigned)((vg + 255) / 256), 256, 0, st>>>(scratch, chunks, vg, cols);
  count_launch();
  if (vg != groups) {
    reduce_fold_columns<<<1, 256, 0, st>>>(cols, (int)vg, out);
    count_launch();
  }
  return cudaGetLastError();
}

}  // namespace fishing
