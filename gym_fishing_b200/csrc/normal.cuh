// Table-assisted Box-Muller in binary64 (round 2): two standard normals from the four words of one Philox block.
//
//   u1 = (k1 + 0.5) 2^-52,  u2 = (k2 + 0.5) 2^-52          (k1, k2: 52 random bits each, as before)
//   R  = sqrt(-2 ln u1);   z0 = R cos(2 pi u2);   z1 = R sin(2 pi u2)
//
// Same mathematical definition as round 1 (oracle/philox_ref.py is unchanged); what changed is how the three
// transcendental pieces are evaluated.  The fused kernels are bound by instruction issue with the FP64 pipe as the busiest
// unit, and the round-1 evaluation (fdlibm log with a Newton reciprocal, Newton sqrt with a final correction, two 6-term
// sincospi polynomials) cost 58 FP64 instructions per pair of draws.  Here the load/store unit, idle in those kernels,
// does part of the work: 28 FP64 instructions + 3 table loads.
//
//   -2 ln u1:  u1 = 2^e m, m in [1, 2).  i = round(256 (m - 1)) selects c_i = 1 + i/256; the table holds -2 RN(1/c_i) and
//              T_i = 2 ln RN(1/c_i) (+ 2 ln 2 for c_i >= sqrt 2, whose exponent is taken as e + 1, so that nothing cancels
//              as u1 -> 1).  With rr = fma(m, -2 rcp_i, 2) = -2 (m rcp_i - 1), |rr| <= 2^-8:
//                  -2 ln u1 = (e + adj_i)(-2 ln 2) + T_i + rr + rr^2 Q(rr),     Q a cubic (abs. error 2^-58.6).
//   sqrt:      R = t (1 + e/2 + 3 e^2 / 8) with t = a y0, e = 1 - t y0, y0 the hardware rsqrt seed (2^-22): 5 instructions.
//   sin/cos:   u2 = j/256 + 2^-53 + f with j = round(256 frac), |f| <= 2^-9 obtained EXACTLY by one subtraction of two
//              numbers in [1, 2]; the table holds cos / sin of 2 pi (j/256 + 2^-53) (the half step of the open interval lives
//              in the table), and sin(2 pi f) = f PS(f^2), cos(2 pi f) - 1 = f^2 PC(f^2) are two quadratics in f^2
//              (abs. error 2^-61 / 2^-67).  The rotation is applied to (R cos_j, R sin_j).
//
// Accuracy: tools/normal_accuracy.cpp evaluates this file on the host against long double; tests/test_gpu_parity.py
// (test_philox_normals_against_long_double_evaluation) does the same for the device: |error| < 3e-15 for every draw.
// The file compiles for host and device; the host build emulates the rsqrt seed (it is only a seed: the result does
// not depend on its low bits beyond the last ulp).
#pragma once
#include <stdint.h>

#include "normal_tables.cuh"

#include <cmath>
#include <cstring>
#if defined(__CUDACC__)
// host + device: the host side of the library derives loop-invariant logarithms (fishing-v6 / v10: log(r + 1)) with the very
// function the kernels use, so that a value derived on the host and one derived in a kernel agree bit for bit
#define FN_HD __host__ __device__ __forceinline__
#else
using std::exp;
using std::fabs;
#define FN_HD inline
struct double2 { double x, y; };
#endif

namespace fishing {

struct Normal2 {
  double z0, z1;
};

namespace nrm {

FN_HD double hilo(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi, (int)lo);
#else
  const uint64_t b = ((uint64_t)hi << 32) | lo;
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
FN_HD uint32_t hi32(double d) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2hiint(d);
#else
  uint64_t b;
  memcpy(&b, &d, 8);
  return (uint32_t)(b >> 32);
#endif
}
FN_HD uint32_t lo32(double d) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2loint(d);
#else
  uint64_t b;
  memcpy(&b, &d, 8);
  return (uint32_t)b;
#endif
}
// every operation is an explicit round-to-nearest intrinsic: no contraction freedom, identical bits in every kernel
FN_HD double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
FN_HD double add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
FN_HD double sub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
FN_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
FN_HD double rsqrt_seed(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  // the hardware unit reads the high word of x and returns a high word (about 2^-22 relative)
  const double y = 1.0 / std::sqrt(hilo(hi32(x), 0));
  return hilo(hi32(y), 0);
#endif
}
FN_HD double i2d(int e) {
#if defined(__CUDA_ARCH__)
  return __int2double_rn(e);
#else
  return (double)e;
#endif
}

}  // namespace nrm

#if defined(__CUDACC__)
// The tables in global memory (one copy per translation unit, 8 KB; log entries first, then the trig entries): read through
// the read-only path by the kernels that draw once per thread (reset, the noise exporter, scalar paths); the kernels that draw
// in their hot loop stage them in shared memory once per block.
struct NormalTables {
  double2 log_e[FISHING_NLOG + 1];
  double2 trig_e[FISHING_NTRIG + 1];
  double exp_e[64];                    // 2^(j/64), for exp_tab() below
};
static __device__ const NormalTables g_normal_tables = {FISHING_LOG_TABLE_INIT, FISHING_TRIG_TABLE_INIT, FISHING_EXP_TABLE_INIT};
// polynomial coefficients: from the constant bank the compiler loads them in pairs into uniform registers
static __constant__ double FN_C[18] = {FISHING_Q3, FISHING_Q2, FISHING_Q1, FISHING_Q0, FISHING_M2LN2, FISHING_PS2,
                                       FISHING_PS1, FISHING_PS0, FISHING_PC2, FISHING_PC1, FISHING_PC0,
                                       FISHING_EXP_64_LN2, FISHING_EXP_NLN2_64_HI, FISHING_EXP_NLN2_64_LO, FISHING_PE3,
                                       FISHING_PE2, FISHING_PE1, FISHING_PE0};
// The high word of 1.0 as a literal.  (Reading it from the constant bank turns (w & 0xFFFFF) | 0x3FF00000 into ONE LOP3 instead
// of two -- tried in round 2, FISHING_AB_ONE_FROM_CONSTANT: one LOP3 less per pair-step, but 2 % slower fused kernels and a 15 %
// slower trace kernel, the register allocation changes for the worse.)
static __constant__ uint32_t FN_ONE_HI_C = 0x3FF00000u;
#ifdef FISHING_AB_ONE_FROM_CONSTANT
#define FN_ONE_HI FN_ONE_HI_C
#else
#define FN_ONE_HI 0x3FF00000u
#endif

struct GlobalTabs {
  __device__ __forceinline__ double exp_at(uint32_t byte_off) const {
    return __ldg(reinterpret_cast<const double*>(reinterpret_cast<const char*>(g_normal_tables.exp_e) + byte_off));
  }
  __device__ __forceinline__ double2 log_at(uint32_t byte_off) const {
    return __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const char*>(g_normal_tables.log_e) + byte_off));
  }
  __device__ __forceinline__ double2 trig_at(uint32_t byte_off) const {
    return __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const char*>(g_normal_tables.trig_e) + byte_off));
  }
};

// Block-shared copy for kernels that draw in their hot loop.
struct SharedTabs {
  const char* lt;
  const char* tt;
  const char* et;
  __device__ __forceinline__ double exp_at(uint32_t byte_off) const { return *reinterpret_cast<const double*>(et + byte_off); }
  __device__ __forceinline__ double2 log_at(uint32_t byte_off) const { return *reinterpret_cast<const double2*>(lt + byte_off); }
  __device__ __forceinline__ double2 trig_at(uint32_t byte_off) const { return *reinterpret_cast<const double2*>(tt + byte_off); }
};
constexpr int NORMAL_TABLE_ENTRIES = FISHING_NLOG + 1 + FISHING_NTRIG + 1 + 32;   // in double2 units; the last 32 = exp_e[64]
// `buf`: NORMAL_TABLE_ENTRIES double2 of shared memory; called by all BS threads of the block.  Contains a __syncthreads().
template <int BS>
__device__ __forceinline__ SharedTabs stage_normal_tables(double2* buf) {
  const double2* src = g_normal_tables.log_e;   // the struct is contiguous
#pragma unroll
  for (int k0 = 0; k0 < NORMAL_TABLE_ENTRIES; k0 += BS) {
    const int k = k0 + (int)threadIdx.x;
    if (k0 + BS <= NORMAL_TABLE_ENTRIES || k < NORMAL_TABLE_ENTRIES) buf[k] = src[k];
  }
  __syncthreads();
  SharedTabs t;
  t.lt = reinterpret_cast<const char*>(buf);
  t.tt = reinterpret_cast<const char*>(buf + FISHING_NLOG + 1);
  t.et = reinterpret_cast<const char*>(buf + FISHING_NLOG + 1 + FISHING_NTRIG + 1);
  return t;
}
#endif
// host copies (the accuracy harness tools/normal_accuracy.cpp, and the host side of the library)
static const double2 h_normal_log_table[FISHING_NLOG + 1] = FISHING_LOG_TABLE_INIT;
static const double2 h_normal_trig_table[FISHING_NTRIG + 1] = FISHING_TRIG_TABLE_INIT;
static const double h_exp_table[64] = FISHING_EXP_TABLE_INIT;
static const double FN_C_H[18] = {FISHING_Q3, FISHING_Q2, FISHING_Q1, FISHING_Q0, FISHING_M2LN2, FISHING_PS2,
                                  FISHING_PS1, FISHING_PS0, FISHING_PC2, FISHING_PC1, FISHING_PC0,
                                  FISHING_EXP_64_LN2, FISHING_EXP_NLN2_64_HI, FISHING_EXP_NLN2_64_LO, FISHING_PE3,
                                  FISHING_PE2, FISHING_PE1, FISHING_PE0};
struct HostTabs {
  double exp_at(uint32_t byte_off) const { return h_exp_table[byte_off >> 3]; }
  double2 log_at(uint32_t byte_off) const { return h_normal_log_table[byte_off >> 4]; }
  double2 trig_at(uint32_t byte_off) const { return h_normal_trig_table[byte_off >> 4]; }
};
#if defined(__CUDA_ARCH__)
#define FN_COEF(k) FN_C[k]
#else
#define FN_COEF(k) FN_C_H[k]
#endif
#if !defined(__CUDACC__)
#define FN_ONE_HI 0x3FF00000u
#endif

template <class Tabs>
FN_HD Normal2 normal_from_words(const uint32_t (&w)[4], const Tabs& tb) {
  using namespace nrm;
  // ---- a = -2 ln u1
  // 52 random bits -> (k + 0.5) 2^-52 without an integer->double conversion: place the bits in the significand of a number
  // in [1, 2) and subtract 1 - 2^-53; the subtraction is exact and normalises the result.
  const uint32_t one_hi = FN_ONE_HI;
  const double u1 = sub(hilo(one_hi | (w[1] & 0xFFFFFu), w[0]), 0x1.fffffffffffffp-1);
  const uint32_t hi = hi32(u1);
  const uint32_t mant = hi & 0xFFFFFu;
  const double2 le = tb.log_at(((mant + 0x800u) >> 8) & 0x1FF0u);         // entry round(256 (m - 1)), 16 bytes each
  const double m = hilo(one_hi | mant, lo32(u1));
  // e + adj: the addition carries into the exponent field exactly when the entry is one of those with c >= sqrt 2
  const int ep = (int)(hi + (0x00100000u - FISHING_LOG_THR) - 0x3FF00000u) >> 20;
  const double rr = fma_(m, le.x, 2.0);
  const double rr2 = mul(rr, rr);
  double q = fma_(FN_COEF(0), rr, FN_COEF(1));
  q = fma_(q, rr, FN_COEF(2));
  q = fma_(q, rr, FN_COEF(3));
  const double a0 = fma_(i2d(ep), FN_COEF(4), le.y);
  const double a = fma_(rr2, q, add(a0, rr));
  // ---- R = sqrt(a), a in [2^-52, 73.5]
  const double y0 = rsqrt_seed(a);
  const double t = mul(a, y0);
  const double e = fma_(-t, y0, 1.0);
  const double R = fma_(mul(t, e), fma_(e, 0.375, 0.5), t);
  // ---- angle
  const uint32_t mant2 = w[3] & 0xFFFFFu;
  const uint32_t t2 = mant2 + 0x800u;
  const double2 cs = tb.trig_at((t2 >> 8) & 0x1FF0u);                     // entry j = round(256 frac)
  const double f = sub(hilo(one_hi | mant2, w[2]), hilo(0x3FF00000u + (t2 & 0x1FF000u), 0u));   // frac - j/256, exact
  const double f2 = mul(f, f);
  const double ps = fma_(fma_(FN_COEF(5), f2, FN_COEF(6)), f2, FN_COEF(7));
  const double pc = fma_(fma_(FN_COEF(8), f2, FN_COEF(9)), f2, FN_COEF(10));
  const double sth = mul(f, ps);      // sin(2 pi f)
  const double cm1 = mul(f2, pc);     // cos(2 pi f) - 1
  const double RC = mul(R, cs.x), RS = mul(R, cs.y);
  Normal2 out;
  out.z0 = fma_(-RS, sth, fma_(RC, cm1, RC));
  out.z1 = fma_(RC, sth, fma_(RS, cm1, RS));
  return out;
}

// exp(x) for the tipping-point growth of fishing-v2 (fishing_tipping_env.py:31: x1 * exp(...)), table-assisted like the
// pieces above: x = k ln2/64 + r with k = round(64 x / ln 2) read from the low bits of x * 64/ln2 + 1.5 * 2^52,
// exp(x) = 2^(k >> 6) * 2^((k & 63)/64) * (1 + r + r^2 P(r)), |r| <= ln2/128, P a cubic (abs. error 2^-57.7).  10 FP64
// instructions and one 8-byte table load against libdevice's ~20 plus two dozen immediates rebuilt every call; mean 0.33,
// worst 1.02 ulp (tools/normal_accuracy.cpp).  The reference computes this with numpy's exp, itself not correctly rounded: fishing-v2
// parity is a 1e-12 tolerance either way.  Beyond |x| = 700 the scaling is done by two multiplications
// (overflow to inf, gradual underflow to 0); infinities and NaN are patched in a branch that is practically never taken.
template <class Tabs>
FN_HD double exp_tab(double x, const Tabs& tb) {
  using namespace nrm;
#ifdef FISHING_AB_LIBEXP
  return exp(x);      // A/B timing build: the library routine
#endif
  const double kf = fma_(x, FN_COEF(11), 0x1.8p52);
  const int ki = (int)lo32(kf);
  const double kd = sub(kf, 0x1.8p52);
  double r = fma_(kd, FN_COEF(12), x);
  r = fma_(kd, FN_COEF(13), r);
  const double r2 = mul(r, r);
  double p = fma_(FN_COEF(14), r, FN_COEF(15));
  p = fma_(p, r, FN_COEF(16));
  p = fma_(p, r, FN_COEF(17));
  const double q = fma_(r2, p, r);                                  // exp(r) - 1
  const double T = tb.exp_at(((uint32_t)ki & 63u) << 3);
  const double y = fma_(T, q, T);                                   // in (0.99, 2): a normal number
  const int m = ki >> 6;
  if (fabs(x) < 700.0) return hilo(hi32(y) + (uint32_t)(m << 20), lo32(y));   // times 2^m: the result stays normal
  // rare: results that overflow, underflow or are subnormal; infinities and NaN (for which everything above is garbage)
  if (!(fabs(x) <= 1400.0)) return x != x ? add(x, x) : (x > 0.0 ? hilo(0x7FF00000u, 0u) : 0.0);
  const int m1 = m >> 1, m2 = m - m1;                               // two normal powers of two: |m| < 2048
  return mul(mul(y, hilo((uint32_t)(m1 + 1023) << 20, 0u)), hilo((uint32_t)(m2 + 1023) << 20, 0u));
}

// log(x) for the log-normal growth family (fishing-v5..v11, growth_models.py:208-261), from the Box-Muller's own log table:
// x = 2^e m, entry i = round(256 (m - 1)), ln x = -((e + adj_i)(-2 ln 2) + T_i + rr + rr^2 Q(rr)) / 2 -- the expression
// normal_from_words() evaluates for -2 ln u1, for any positive x.  T_0 = T_256 = 0 and e + adj = 0 around x = 1, so nothing
// cancels there.  Mean 0.45 ulp, worst 2.1 ulp over all binades; for |ln x| < 2^-6 the cubic's truncation error (2^-58.6 = 2e-18
// absolute, up to 6 ulp of so small a result) is what is left -- the growth functions pass mu = log(x) + ... through exp, so
// it is the absolute error that counts (tools/normal_accuracy.cpp).  11 FP64 instructions and one 16-byte table load against
// libdevice's ~35.  Zero, subnormal, negative, infinite and NaN arguments
// leave through a rarely taken branch (an extinct stock, x = 0, is the one that occurs: -inf, as numpy).
template <class Tabs>
FN_HD double log_tab(double x, const Tabs& tb) {
  using namespace nrm;
  uint32_t hi = hi32(x), lo = lo32(x);
  int ebias = 0;
  if (hi - 0x00100000u >= 0x7FE00000u) {                   // not a positive normal number
    if (x != x) return add(x, x);
    if (x == 0.0) return -hilo(0x7FF00000u, 0u);          // log(+-0) = -inf
    if (x < 0.0) return hilo(0x7FF80000u, 0u);            // log(negative) = NaN
    if (hi == 0x7FF00000u) return x;                       // +inf
    const double xs = mul(x, 0x1p54);                      // subnormal
    hi = hi32(xs);
    lo = lo32(xs);
    ebias = -54;
  }
  const uint32_t mant = hi & 0xFFFFFu;
  const double2 le = tb.log_at(((mant + 0x800u) >> 8) & 0x1FF0u);
  const double m = hilo(FN_ONE_HI | mant, lo);
  const int ep = ((int)(hi + (0x00100000u - FISHING_LOG_THR) - 0x3FF00000u) >> 20) + ebias;
  const double rr = fma_(m, le.x, 2.0);
  const double rr2 = mul(rr, rr);
  double q = fma_(FN_COEF(0), rr, FN_COEF(1));
  q = fma_(q, rr, FN_COEF(2));
  q = fma_(q, rr, FN_COEF(3));
  const double a0 = fma_(i2d(ep), FN_COEF(4), le.y);
  const double a = fma_(rr2, q, add(a0, rr));
  return fma_(a, -0.5, 0.0);                               // + 0.0: log(1) is +0, not -0
}

}  // namespace fishing
