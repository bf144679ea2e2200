// Per-environment arithmetic of gym_fishing's step(), one IEEE-754 binary64 rounding per reference operation, in
// the reference's evaluation order.  Every add/sub/mul/div goes through the __d*_rn
// intrinsics, which the compiler never contracts into FMA, so the result is bit-identical to numpy/CPython.
#pragma once
#include <stdint.h>

#include "../../include/fishing_b200.h"
#include "philox.cuh"   // GlobalTabs / exp_tab (csrc/normal.cuh)

namespace fishing {

enum Growth { LOGISTIC = 0, TIPPING = 1,
              // log-normal family (gym_fishing/envs/growth_models.py:208-261); G_MODEL picks one of them per environment
              G_ALLEN = 2, G_BEVERTON_HOLT = 3, G_MYERS = 4, G_MAY = 5, G_RICKER = 6, G_MODEL = 7 };

// growth-function parameters of the log-normal family (the reference's `params` dict) and the fishing-v11 model table
struct GrowthP {
  double r, K, C, M, theta, q, b, a, sigma;
  // loop-invariant sub-expressions of mu(x), derived once per launch on the host (or per step on the device for the
  // drifting r of fishing-v10) instead of once per environment-step:
  double logA;  // Beverton-Holt: log(clip(r,0,inf) + 1); Myers: log(r + 1)
  double B;     // Beverton-Holt: clip(K,0,inf) / clip(r,0,inf)
  double omC;   // Allen: 1 - C
  double bq;    // May: b^q
};

// The derived fields, same operations as the reference's expressions (growth_models.py:213-215,222-225,239,248-253).  The
// logarithm is csrc/normal.cuh's log_tab on BOTH sides: explicit fma / mul / add on the same tables, so the host (once per
// launch: K1, also for fishing-v10 whose r drifts from launch to launch) and a kernel (fishing-v10 inside the fused K2, every
// step) get the same bits.  -DFISHING_AB_LIBLOG: the library's log (then fishing-v10 derives on the device in K1 as well).
__host__ __device__ inline double growth_log(double x) {
#if defined(FISHING_AB_LIBLOG)
  return log(x);
#elif defined(__CUDA_ARCH__)
  return log_tab(x, GlobalTabs());
#else
  return log_tab(x, HostTabs());
#endif
}
__host__ __device__ inline void growth_derive(GrowthP& g, int fn) {
  g.logA = 0.0; g.B = 0.0; g.omC = 1.0 - g.C; g.bq = 0.0;
  if (fn == FISHING_FN_BEVERTON_HOLT) {
    const double rc = (g.r < 0.0) ? 0.0 : g.r;
    g.logA = growth_log(rc + 1.0);
    g.B = ((g.K < 0.0) ? 0.0 : g.K) / rc;
  } else if (fn == FISHING_FN_MYERS) {
    g.logA = growth_log(g.r + 1.0);
  } else if (fn == FISHING_FN_MAY) {
    g.bq = pow(g.b, g.q);
  }
}
struct ModelTable {
  int32_t n_models;
  int32_t fn[FISHING_MAX_MODELS];
  GrowthP p[FISHING_MAX_MODELS];
};
// how r and K reach the kernel: shared scalars with K == 1 (the default; x/K and x*K are exact no-ops and are
// skipped), shared scalars with general K, or per-environment arrays (fishing-v4).
enum ParamMode { P_KONE = 0, P_UNIFORM = 1, P_PERENV = 2 };

// Python builtin min(a, b): b if b < a else a (a NaN b returns a).   base_fishing_env.py:117
__device__ __forceinline__ double py_min(double a, double b) { return (b < a) ? b : a; }
// Python builtin max(a, b): b if b > a else a.                        base_fishing_env.py:74,118; policies.py:29
__device__ __forceinline__ double py_max(double a, double b) { return (b > a) ? b : a; }
// np.maximum(a, 0.0): NaN propagates, -0.0 -> +0.0.                   base_fishing_env.py:125-132
__device__ __forceinline__ double np_max0(double a) { return (a > 0.0 || a != a) ? a : 0.0; }
// np.clip(a, lo, hi): NaN propagates.                                 base_fishing_env.py:143-145
__device__ __forceinline__ double np_clip(double a, double lo, double hi) { return (a < lo) ? lo : ((a > hi) ? hi : a); }
__device__ __forceinline__ float np_clipf(float a, float lo, float hi) { return (a < lo) ? lo : ((a > hi) ? hi : a); }

// IEEE a / b for the quotients of this file, whose numerators are very often exactly zero (an extinct stock, a zero
// quota, action 0).  The hardware division sequence sends a zero numerator to its ~55-instruction slow path, and one such
// lane drags its whole warp there; the profile of the fishing-v4 kernel showed 220 of 880 instructions per environment
// pair in that subroutine.  A zero numerator has a known quotient -- zero with the sign of a*b, or NaN when b is 0 or
// NaN -- so it is substituted by 1.0 for the division proper and patched afterwards.  Bit-identical to a / b.
__device__ __forceinline__ double div_z(double a, double b) {
  const bool z = (a == 0.0);
  const double q = __ddiv_rn(z ? 1.0 : a, b);
  if (!z) return q;
  const double signed_zero = __dmul_rn(a, (b < 0.0) ? -1.0 : 1.0);
  return (b == b && b != 0.0) ? signed_zero : __longlong_as_double(0x7ff8000000000000LL);
}

// The same for a divisor that is a constructor parameter and very often exactly 1 (the log-normal family's default K and
// M): a / 1.0 is a for every a, so the division sequence (~25 instructions) is skipped behind a warp-uniform test.
__device__ __forceinline__ double div_u(double a, double b) {
  if (b == 1.0) return a;
  return div_z(a, b);
}

// get_quota, base_fishing_env.py:135-147
template <int PM>
__device__ __forceinline__ double quota_f64(double a, double K) {
  const double t = __dadd_rn(np_clip(a, -1.0, 1.0), 1.0);
  return PM == P_KONE ? t : __dmul_rn(t, K);
}
// float32 Box action under numpy>=2 promotion: clip, +1 and (for a Python-number K) *K are float32 operations; with
// an np.float64 K (fishing-v4) the product is float64.
template <int PM>
__device__ __forceinline__ double quota_f32(float a, double K) {
  const float t = __fadd_rn(np_clipf(a, -1.0f, 1.0f), 1.0f);
  if (PM == P_PERENV) return __dmul_rn((double)t, K);
  if (PM == P_KONE) return (double)t;
  return (double)__fmul_rn(t, (float)K);
}
// a / n for an integer-valued finite a and a count n in [1, 2^31], given rn = RN(1/n) from the host: two residual
// corrections instead of the hardware division sequence (5 FP64 instructions instead of ~14 + a reciprocal seed).  After
// the first correction q is a faithful quotient, and Markstein's theorem (IBM J. Res. Dev. 34, 1990) then makes the second
// one the correctly rounded a / n -- the same bits as __ddiv_rn(a, n); tests/test_gpu_parity.py checks all 6.5 million
// (a, n) pairs with n <= 2048 and |a| around the action range, plus random large ones.
__device__ __forceinline__ double div_count(double a, double n, double rn) {
  double q = __dmul_rn(a, rn);
  q = __fma_rn(__fma_rn(-q, n, a), rn, q);
  return __fma_rn(__fma_rn(-q, n, a), rn, q);
}
// a / K for a carrying capacity K that is a constructor parameter (the same for every environment, but not 1): the host
// supplies rK = RN(1 / K) and says whether K is an ordinary number (2^-500 <= K <= 2^500).  Then, for a zero or
// with 2^-500 <= |a| < 2^501, nothing in the two-correction sequence can overflow or underflow and Markstein's theorem
// applies as in div_count; everything else (NaN, infinities, extreme magnitudes, unusual K) takes the IEEE division.
struct KDiv {
  double rK;
  int fast;
  int odd;   // the uniform parameters are not (K > 0, r >= 0): see "sigma == 0" in grow()
};
inline KDiv make_kdiv(double K, double r) {
  KDiv k;
  k.fast = (K >= 0x1p-500 && K <= 0x1p500) ? 1 : 0;   // positive ordinary numbers only; false for NaN
  k.rK = k.fast ? 1.0 / K : 0.0;
  k.odd = (K > 0.0 && r >= 0.0) ? 0 : 1;
  return k;
}
__device__ __forceinline__ double div_k(double a, double K, const KDiv& kd) {
  const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;   // biased exponent of a
  if (kd.fast && (e - 523u < 1001u || a == 0.0)) {
    const double q0 = __dmul_rn(a, kd.rK);                            // for a zero: the zero of the right sign
    double q = __fma_rn(__fma_rn(-q0, K, a), kd.rK, q0);
    q = __fma_rn(__fma_rn(-q, K, a), kd.rK, q);
    return (a == 0.0) ? q0 : q;
  }
  return div_z(a, K);
}
// fishing-v4: the divisor is the environment's OWN K, and a step divides by it twice (x1 / K in population_draw, x2 / K in
// get_state).  One correctly rounded reciprocal -- __drcp_rn(K), the same bits as the host's 1.0 / K -- serves both quotients
// through div_k (5 FP64 instructions each) instead of two hardware division sequences (~25 instructions each with the
// zero-numerator patch of div_z): the fishing-v4 lock-step kernel executed 389 instructions per warp outside its redraw pass,
// about 100 of them in these four divisions of a pair.  -DFISHING_V4_RCP=0: the round-1 form (A/B, profiles/ab_r02.md).
#ifndef FISHING_V4_RCP
#define FISHING_V4_RCP 1
#endif
__device__ __forceinline__ KDiv kdiv_env(double K, int odd = 0) {
  KDiv k;
  k.fast = (K >= 0x1p-500 && K <= 0x1p500) ? 1 : 0;   // positive ordinary numbers only: false for NaN and for a K clipped to 0 (then div_z)
  k.rK = __drcp_rn(K);                                 // never read when !fast
  k.odd = odd;
  return k;
}
// x / K by parameter mode; for P_PERENV `kd` has to be the ENVIRONMENT's (kdiv_env), never the launch's
template <int PM>
__device__ __forceinline__ double div_pm(double x, double K, const KDiv& kd) {
  if (PM == P_KONE) return x;
  if (PM == P_UNIFORM || FISHING_V4_RCP) return div_k(x, K, kd);
  return div_z(x, K);
}

// (a / n_actions) * K for a Discrete action, base_fishing_env.py:139-141
template <int PM>
__device__ __forceinline__ double quota_int(long long a, double n_actions, double rcp_n_actions, double K) {
  const bool count = n_actions >= 1.0 && n_actions <= 2147483648.0;   // always, for a gym Discrete space
  const double t = count ? div_count((double)a, n_actions, rcp_n_actions) : div_z((double)a, n_actions);
  return PM == P_KONE ? t : __dmul_rn(t, K);
}

// get_action, base_fishing_env.py:149-156
template <int PM>
__device__ __forceinline__ double action_cts(double q, double K, const KDiv& kd = KDiv()) {
  return __dsub_rn(PM == P_KONE ? q : (PM == P_UNIFORM ? div_k(q, K, kd) : div_z(q, K)), 1.0);
}
template <int PM>
__device__ __forceinline__ double action_discrete(double q, double n_actions, double K, const KDiv& kd = KDiv()) {
  const double t = __dmul_rn(q, n_actions);
  return rint(PM == P_KONE ? t : (PM == P_UNIFORM ? div_k(t, K, kd) : div_z(t, K)));  // Python round(): half to even
}

// population_draw: base_fishing_env.py:121-133 (logistic) / fishing_tipping_env.py:24-35 (tipping point)
template <int GROWTH, int PM, class Tabs = GlobalTabs>
__device__ __forceinline__ double grow(double x1, double r, double K, double C, double sigma, double z, bool noisy,
                                       const KDiv& kd = KDiv(), const Tabs& tb = Tabs()) {
  const double frac = div_pm<PM>(x1, K, kd);      // P_PERENV: kd is the environment's own (step_core)
  const double room = __dsub_rn(1.0, frac);
  if (GROWTH == LOGISTIC) {
    double t = __dadd_rn(x1, __dmul_rn(__dmul_rn(r, x1), room));
    // sigma == 0: the reference still adds (x1 * 0) * z.  For a finite stock that is +-0.0 and cannot change the clipped
    // result, so no draw is made.  For an infinite stock it is NaN -- which the growth term already is (inf - inf) whenever
    // K > 0 and r >= 0; only for other parameter signs (or caller-supplied per-env K, r) does the term have to be added
    // (DESIGN.md "noise when sigma = 0").
    if (noisy) t = __dadd_rn(t, __dmul_rn(__dmul_rn(x1, sigma), z));
    else if (PM == P_PERENV || kd.odd) t = __dadd_rn(t, __dmul_rn(x1, 0.0));
    return np_max0(t);
  } else {
    double e = __dmul_rn(__dmul_rn(r, room), __dsub_rn(x1, C));
    if (noisy) e = __dadd_rn(e, __dmul_rn(__dmul_rn(x1, sigma), z));
    else if (PM == P_PERENV || kd.odd) e = __dadd_rn(e, __dmul_rn(x1, 0.0));
    return np_max0(__dmul_rn(x1, exp_tab(e, tb)));
  }
}

// mu(x) of the five growth functions, growth_models.py:208-261, in the reference's evaluation order.  log and exp are the
// table-assisted ones of csrc/normal.cuh (log_tab: <= 2.1 ulp, exp_tab: <= 1.02 ulp; numpy's own are within 1 ulp), the powers
// x^q / x^theta are exp(q log x) (x >= +0 or NaN here; a few ulps from numpy's power for the |q log x| < 10 of these models):
// the family's parity is the 1e-12 tolerance of fishing-v2.  -DFISHING_AB_LIBLOG: libdevice's log / pow (round 1).  Everything
// else is one IEEE operation per reference operation.
#ifdef FISHING_AB_LIBLOG
__device__ __forceinline__ double ln_(double x) { return log(x); }
__device__ __forceinline__ double pow_ln(double x, double lx, double p) { (void)lx; return pow(x, p); }
#else
__device__ __forceinline__ double ln_(double x) { return log_tab(x, GlobalTabs()); }
// x^p given lx = log(x), x >= +0 or NaN.  0 * inf would be NaN where numpy's power(x, 0) is 1 for every x
__device__ __forceinline__ double pow_ln(double x, double lx, double p) {
  (void)x;
  return p == 0.0 ? 1.0 : exp_tab(__dmul_rn(p, lx), GlobalTabs());
}
#endif

template <int FN>
__device__ __forceinline__ double lognormal_mu(double x, const GrowthP& g) {
  if (FN == FISHING_FN_ALLEN) {            // np.log(x) + r * (1 - x / K) * (1 - C) / K
    const double y = div_u(__dmul_rn(__dmul_rn(g.r, __dsub_rn(1.0, div_u(x, g.K))), g.omC), g.K);
    return __dadd_rn(ln_(x), y);
  } else if (FN == FISHING_FN_BEVERTON_HOLT) {   // log(A) + log(x) - log(1 + x / B), A = r + 1, B = K / r
    const double xc = (x < 0.0) ? 0.0 : x;                       // np.clip(x, 0, inf)
    return __dsub_rn(__dadd_rn(g.logA, ln_(xc)), ln_(__dadd_rn(1.0, div_z(xc, g.B))));
  } else if (FN == FISHING_FN_MAY) {       // x + x*r*(1 - x/M) - a*x^q / (x^q + b^q), then log
    const double xq = pow_ln(x, ln_(x), g.q);
    const double e = __dsub_rn(__dadd_rn(x, __dmul_rn(__dmul_rn(x, g.r), __dsub_rn(1.0, div_z(x, g.M)))),
                               div_z(__dmul_rn(g.a, xq), __dadd_rn(xq, g.bq)));
    return ln_(e);
  } else if (FN == FISHING_FN_MYERS) {     // log(r + 1) + theta*log(x) - log(1 + x^theta / M)
    const double lx = ln_(x);
    return __dsub_rn(__dadd_rn(g.logA, __dmul_rn(g.theta, lx)),
                     ln_(__dadd_rn(1.0, div_u(pow_ln(x, lx, g.theta), g.M))));
  } else {                                 // Ricker: log(x) + r * (1 - x / K)
    return __dadd_rn(ln_(x), __dmul_rn(g.r, __dsub_rn(1.0, div_u(x, g.K))));
  }
}

// np.maximum(0, np.random.lognormal(mu, sigma)) with the legacy generator's lognormal = exp(mu + sigma * gauss)
template <int FN>
__device__ __forceinline__ double grow_lognormal(double x1, const GrowthP& g, double z, bool noisy) {
  double e = lognormal_mu<FN>(x1, g);
  if (noisy) e = __dadd_rn(e, __dmul_rn(g.sigma, z));
  const double v = exp_tab(e, GlobalTabs());     // csrc/normal.cuh: table-assisted, <= 1.02 ulp (the family's parity is a tolerance)
  return (0.0 > v) ? 0.0 : v;
}

// fishing-v11: the growth function differs from lane to lane.  A plain switch over the five functions makes every warp
// walk through five branches, each with its own library calls (a pow and a second log for Myers and May, a second log for
// Beverton-Holt): ~10 transcendental evaluations per step at 20 % lane occupancy, 1.9e10 env-steps/s in round 1.  Here ALL
// library calls are shared by the warp -- log(x), one power x^p written as exp(p log x) (p = theta, q, or unused), one second
// logarithm of a per-model argument (1 for the models without one: log(1) = +0 exactly), and the final exp -- and only the
// few operations that prepare their arguments sit in the divergent switch.  Same operations in the same order as
// lognormal_mu<FN> except for the power (pow(x, p) there, exp(p log x) here: a few ulps apart, the family's parity is a
// 1e-12 tolerance).
__device__ __forceinline__ double grow_by_model(int fn, double x1, const GrowthP& g, double z, bool noisy) {
  const double lx = ln_(x1);    // x1 >= +0 or NaN here, so Beverton-Holt's np.clip(x, 0, inf) is the identity
  const bool has_pow = fn == FISHING_FN_MYERS || fn == FISHING_FN_MAY;
  const double pw = pow_ln(x1, lx, fn == FISHING_FN_MYERS ? g.theta : g.q);   // x^theta / x^q (unused otherwise)
  double base = 0.0, arg = 1.0;
  switch (fn) {
    case FISHING_FN_ALLEN:           // np.log(x) + r * (1 - x / K) * (1 - C) / K
      base = __dadd_rn(lx, div_u(__dmul_rn(__dmul_rn(g.r, __dsub_rn(1.0, div_u(x1, g.K))), g.omC), g.K));
      break;
    case FISHING_FN_BEVERTON_HOLT:   // log(A) + log(x) - log(1 + x / B)
      base = __dadd_rn(g.logA, lx);
      arg = __dadd_rn(1.0, div_z(x1, g.B));
      break;
    case FISHING_FN_MYERS:           // log(r + 1) + theta * log(x) - log(1 + x^theta / M)
      base = __dadd_rn(g.logA, __dmul_rn(g.theta, lx));
      arg = __dadd_rn(1.0, div_u(pw, g.M));
      break;
    case FISHING_FN_MAY:             // log(x + x * r * (1 - x / M) - a * x^q / (x^q + b^q))
      arg = __dsub_rn(__dadd_rn(x1, __dmul_rn(__dmul_rn(x1, g.r), __dsub_rn(1.0, div_z(x1, g.M)))),
                      div_z(__dmul_rn(g.a, pw), __dadd_rn(pw, g.bq)));
      break;
    default:                         // Ricker: log(x) + r * (1 - x / K)
      base = __dadd_rn(lx, __dmul_rn(g.r, __dsub_rn(1.0, div_u(x1, g.K))));
      break;
  }
  (void)has_pow;
  const double l2 = ln_(arg);
  double mu = fn == FISHING_FN_MAY ? l2 : __dsub_rn(base, l2);       // base - (+0) = base for Allen and Ricker
  if (noisy) mu = __dadd_rn(mu, __dmul_rn(g.sigma, z));
  const double v = exp_tab(mu, GlobalTabs());
  return (0.0 > v) ? 0.0 : v;
}

template <int GROWTH>
__device__ __forceinline__ double grow_ln(double x1, const GrowthP& g, const ModelTable& tab, int model, double z,
                                          bool noisy_uniform) {
  if (GROWTH == G_ALLEN) return grow_lognormal<FISHING_FN_ALLEN>(x1, g, z, noisy_uniform);
  if (GROWTH == G_BEVERTON_HOLT) return grow_lognormal<FISHING_FN_BEVERTON_HOLT>(x1, g, z, noisy_uniform);
  if (GROWTH == G_MYERS) return grow_lognormal<FISHING_FN_MYERS>(x1, g, z, noisy_uniform);
  if (GROWTH == G_MAY) return grow_lognormal<FISHING_FN_MAY>(x1, g, z, noisy_uniform);
  if (GROWTH == G_RICKER) return grow_lognormal<FISHING_FN_RICKER>(x1, g, z, noisy_uniform);
  // fishing-v11: the model (and its own sigma) belongs to the environment
  return grow_by_model(tab.fn[model], x1, tab.p[model], z, true);
}

struct StepOut {
  double s2;      // next normalised state
  double reward;  // max(harvest, 0)
  bool dead;      // stock <= 0 after growth
  double stock;   // fish_population after growth (x2), before the x/K - 1 round trip
  double harvest; // min(x, quota); differs from reward only for a negative quota
};

// BaseFishingEnv.step body, base_fishing_env.py:60-81, for one environment given its quota.
template <int GROWTH, int PM>
__device__ __forceinline__ StepOut step_core(double s, double quota, double z, double r, double K, double C,
                                             double sigma, bool noisy, const GrowthP* gp = nullptr,
                                             const ModelTable* tab = nullptr, int model = 0, const KDiv& kd = KDiv()) {
  // gp / tab are only dereferenced by the log-normal growth family (fishing-v5..v11)
  double x = __dadd_rn(s, 1.0);                     // get_fish_population :158-160
  if (PM != P_KONE) x = __dmul_rn(x, K);
  const double h = py_min(x, quota);                // harvest_draw :117
  // :118 max(x - h, 0.0) is the identity: h = min(x, quota) <= x (or h = x when the quota is NaN), so x - h is +0, a
  // positive number or NaN for every input, and Python's max(v, 0.0) returns v in all three cases
  const double x1 = __dsub_rn(x, h);
  double x2;
  // fishing-v4: one reciprocal of this environment's K for the step's two quotients (kdiv_env)
  const KDiv kde = (PM == P_PERENV && FISHING_V4_RCP && (GROWTH == LOGISTIC || GROWTH == TIPPING)) ? kdiv_env(K, kd.odd) : kd;
  if constexpr (GROWTH == LOGISTIC || GROWTH == TIPPING) x2 = grow<GROWTH, PM>(x1, r, K, C, sigma, z, noisy, kde);
  else x2 = grow_ln<GROWTH>(x1, *gp, *tab, model, z, noisy);
  StepOut o;
  // get_state :162-164 (the log-normal family always takes the general-K path, with K = 1 by default: div_u)
  if constexpr (GROWTH == LOGISTIC || GROWTH == TIPPING) o.s2 = __dsub_rn(div_pm<PM>(x2, K, kde), 1.0);
  else o.s2 = __dsub_rn(div_u(x2, K), 1.0);
  o.reward = py_max(h, 0.0);                        // :74
  o.dead = x2 <= 0.0;                               // :78
  o.stock = x2;
  o.harvest = h;
  return o;
}

// action element types and their 2-wide vector forms
template <int ACT>
struct ActT;
template <>
struct ActT<FISHING_ACT_F64> { using type = double; using vec2 = double2; };
template <>
struct ActT<FISHING_ACT_F32> { using type = float; using vec2 = float2; };
template <>
struct ActT<FISHING_ACT_I32> { using type = int32_t; using vec2 = int2; };
template <>
struct ActT<FISHING_ACT_I64> { using type = long long; using vec2 = longlong2; };

template <int PM, int ACT>
__device__ __forceinline__ double quota_of(typename ActT<ACT>::type a, double n_actions, double rcp_n_actions, double K) {
  if constexpr (ACT == FISHING_ACT_F64) return quota_f64<PM>(a, K);
  else if constexpr (ACT == FISHING_ACT_F32) return quota_f32<PM>(a, K);
  else return quota_int<PM>((long long)a, n_actions, rcp_n_actions, K);
}

// fishing_model_error.py:37-38 / :42-43: K = clip(K_mean + sigma_p*zK, 0, 1e6); r likewise (K first).
__device__ __forceinline__ double v4_param(double mean, double sigma_p, double z) {
  return np_clip(__dadd_rn(mean, __dmul_rn(sigma_p, z)), 0.0, 1e6);
}

}  // namespace fishing
