// Counter-based Gaussian noise for the fishing engine: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller in
// binary64.  This replaces the reference's global numpy MT19937 draw np.random.normal(0, 1)
// (gym_fishing/envs/base_fishing_env.py:130, fishing_tipping_env.py:31, fishing_model_error.py:37-43): the
// stream is a pure function of (seed, noise id, step, stream id), so results do not depend on how environments
// are partitioned over threads, launches or GPUs.  oracle/philox_ref.py restates the same spec in numpy.
//
//   key      = (seed & 0xffffffff, seed >> 32)
//   counter  = (id & 0xffffffff, id >> 32, t & 0xffffffff, (stream << 24) | ((t >> 32) & 0xffffff))
//   (w0..w3) = philox4x32_10(counter, key)
//   u1       = ((w1 & 0xfffff) * 2^32 + w0 + 0.5) * 2^-52          in (0, 1), exact in binary64
//   u2       = ((w3 & 0xfffff) * 2^32 + w2 + 0.5) * 2^-52
//   R        = sqrt(-2 ln u1);   z0 = R cos(2 pi u2);   z1 = R sin(2 pi u2)
//
// Process noise (stream 0) uses id = env_id >> 1 and hands z0 to the even and z1 to the odd environment of the
// pair, so one Philox block + one Box-Muller serves two environments.
//
// The three transcendental pieces are written out here instead of calling libdevice's log / sqrt / sincospi: the
// arguments are known to be normal numbers in (0, 1), (0, 73] and (0, 2), so the denormal / inf / NaN / exact-integer
// paths of the general routines (about a third of their instructions) are dead weight in a kernel that is
// instruction-issue bound.  Same classical algorithms (fdlibm's log and kernel sin/cos polynomials, Newton square root
// from the hardware reciprocal-sqrt seed), < 1.5 ulp each; tests/test_gpu_parity.py checks the draws against an
// 80-bit long-double evaluation of the formulas above.
#pragma once
#include <stdint.h>

namespace fishing {

struct Normal2 {
  double z0, z1;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&w)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  w[0] = c0;
  w[1] = c1;
  w[2] = c2;
  w[3] = c3;
}

// The ten round keys (k + i*W) depend only on the seed: the host precomputes them into the kernel parameters, where the
// xor of each round can take them straight from the constant bank (saves the 18 key additions per Philox block).
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
inline PhiloxKeys philox_round_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    k.k0[i] = a;
    k.k1[i] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}

__device__ __forceinline__ void philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& k,
                                                   uint32_t (&w)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k.k0[i];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k.k1[i];
    c3 = lo0;
  }
  w[0] = c0;
  w[1] = c1;
  w[2] = c2;
  w[3] = c3;
}

// 52 random bits -> (k + 0.5) * 2^-52 without an integer->double conversion: place the bits in the mantissa of a
// number in [1, 2) and subtract (1 - 2^-53); the subtraction is exact.
__device__ __forceinline__ double u52_to_open_unit(uint32_t lo, uint32_t hi) {
  const double d = __hiloint2double(0x3FF00000 | (hi & 0xFFFFF), lo);
  return __dsub_rn(d, 0x1.fffffffffffffp-1);
}

__device__ __forceinline__ double rcp_seed(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// Every operation below is an explicit round-to-nearest intrinsic, so the compiler has no contraction freedom and the
// three kernels that inline this code (K1, K2, noise export) produce identical bits.
#define FD_MUL(a, b) __dmul_rn((a), (b))
#define FD_ADD(a, b) __dadd_rn((a), (b))
#define FD_SUB(a, b) __dsub_rn((a), (b))
#define FD_FMA(a, b, c) __fma_rn((a), (b), (c))

// Polynomial coefficients live in constant memory: literal doubles are materialised with two UMOV instructions each,
// every loop iteration (44 of the 272 instructions of a K2 pair-step); from the constant bank the compiler loads two
// doubles per LDCU.128 and keeps them in uniform registers across iterations.
static __constant__ double FD_C[24] = {
    // [0..6] fdlibm log: Lg7 .. Lg1
    1.479819860511658591e-01, 1.531383769920937332e-01, 1.818357216161805012e-01, 2.222219843214978396e-01,
    2.857142874366239149e-01, 3.999999999940941908e-01, 6.666666666666735130e-01,
    // [7] ln2_hi, [8] ln2_lo
    6.93147180369123816490e-01, 1.90821492927058770002e-10,
    // [9..14] fdlibm kernel sin: S6 .. S1
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
    8.33333333332248946124e-03, -1.66666666666666324348e-01,
    // [15..20] fdlibm kernel cos: C6 .. C1
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02,
    // [21] pi_hi, [22] pi_lo, [23] 2^52 + 2^31
    3.141592653589793116, 1.2246467991473532e-16, 4503601774854144.0};

// ln(u) for a NORMAL double 0 < u < 1 (fdlibm e_log.c algorithm: u = 2^e * m, m in [sqrt(1/2), sqrt(2)),
// s = (m-1)/(m+1), log(m) = f - s*(f - R(s^2)), f = m - 1).
__device__ __forceinline__ double log_unit(double u) {
  int hi = __double2hiint(u);
  const int lo = __double2loint(u);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000FFFFF) | 0x3FF00000;
  if (hi >= 0x3FF6A09F) {  // m >= sqrt(2) (to the precision of the high word): halve it
    hi -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(hi, lo);
  const double f = FD_SUB(m, 1.0);
  const double g = FD_ADD(m, 1.0);
  double rg = rcp_seed(g);                      // ~2^-20 relative
  double t = FD_FMA(-g, rg, 1.0);
  t = FD_FMA(t, t, t);
  rg = FD_FMA(rg, t, rg);                       // ~2^-60
  double s = FD_MUL(f, rg);
  s = FD_FMA(FD_FMA(-s, g, f), rg, s);          // one correction step: s = f/g to < 1 ulp
  const double z = FD_MUL(s, s);
  double R = FD_C[0];
  R = FD_FMA(R, z, FD_C[1]);
  R = FD_FMA(R, z, FD_C[2]);
  R = FD_FMA(R, z, FD_C[3]);
  R = FD_FMA(R, z, FD_C[4]);
  R = FD_FMA(R, z, FD_C[5]);
  R = FD_FMA(R, z, FD_C[6]);
  R = FD_MUL(R, z);
  const double lm = FD_FMA(-s, FD_SUB(f, R), f);  // log(m)
  // e as a double without a conversion instruction: the double with bits 0x43300000:(e xor 2^31) is 2^52 + 2^31 + e
  const double ed = FD_SUB(__hiloint2double(0x43300000, e ^ 0x80000000), FD_C[23]);
  return FD_FMA(ed, FD_C[7], FD_FMA(ed, FD_C[8], lm));
}

// sqrt(a) for a normal double a > 0 (no range checks): Newton on the hardware reciprocal-sqrt seed.
__device__ __forceinline__ double sqrt_pos(double a) {
  double y = rsqrt_seed(a);                     // ~2^-20
  const double e = FD_FMA(FD_MUL(-a, y), y, 1.0);
  y = FD_FMA(FD_MUL(y, FD_FMA(e, 0.375, 0.5)), e, y);  // y ~ a^-1/2 to ~2^-56
  double g = FD_MUL(a, y);
  const double h = FD_MUL(0.5, y);
  g = FD_FMA(FD_FMA(-g, g, a), h, g);           // final correction
  return g;
}

// sin(pi t), cos(pi t) for 0 < t < 2 (fdlibm k_sin.c / k_cos.c polynomials on |pi*y| <= pi/4).
__device__ __forceinline__ void sincospi_02(double t, double& sn, double& cs) {
  const double qd = rint(FD_ADD(t, t));         // nearest quarter-turn index 0..4
  const int q = (int)qd;
  const double y = FD_FMA(qd, -0.5, t);         // exact, |y| <= 1/4
  const double r = FD_FMA(y, FD_C[22], FD_MUL(y, FD_C[21]));  // pi*y, pi = hi + lo
  const double z = FD_MUL(r, r);
  double ps = FD_C[9];
  ps = FD_FMA(ps, z, FD_C[10]);
  ps = FD_FMA(ps, z, FD_C[11]);
  ps = FD_FMA(ps, z, FD_C[12]);
  ps = FD_FMA(ps, z, FD_C[13]);
  ps = FD_FMA(ps, z, FD_C[14]);
  const double s = FD_FMA(FD_MUL(r, z), ps, r);
  double pc = FD_C[15];
  pc = FD_FMA(pc, z, FD_C[16]);
  pc = FD_FMA(pc, z, FD_C[17]);
  pc = FD_FMA(pc, z, FD_C[18]);
  pc = FD_FMA(pc, z, FD_C[19]);
  pc = FD_FMA(pc, z, FD_C[20]);
  const double c = FD_FMA(z, FD_FMA(z, pc, -0.5), 1.0);
  // t = q/2 + y:  q odd -> (sin, cos) = (+-cos, -+sin) of pi*y
  const bool swap = q & 1;
  const double a = swap ? c : s;
  const double b = swap ? s : c;
  // negation = flipping the sign bit of the high word (one shift + one 3-input logic op instead of DADD + 2 selects)
  sn = __hiloint2double(__double2hiint(a) ^ ((q << 30) & 0x80000000), __double2loint(a));
  cs = __hiloint2double(__double2hiint(b) ^ (((q + 1) << 30) & 0x80000000), __double2loint(b));
}

// Uniform integer in [0, n) from the first Philox word (fishing-v11 model draw, np.random.choice(models)):
// floor(w0 * n / 2^32).
__device__ __forceinline__ uint32_t uniform_index(uint64_t seed, uint64_t id, uint64_t t, uint32_t stream, uint32_t n) {
  uint32_t w[4];
  philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu),
                (uint32_t)seed, (uint32_t)(seed >> 32), w);
  return __umulhi(w[0], n);
}

__device__ __forceinline__ Normal2 normal_from_words(const uint32_t (&w)[4]);

// same stream as normal_pair(seed, ...), with the round keys precomputed on the host
__device__ __forceinline__ Normal2 normal_pair_keys(const PhiloxKeys& k, uint64_t id, uint64_t t, uint32_t stream) {
  uint32_t w[4];
  philox4x32_10_keys((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu), k, w);
  return normal_from_words(w);
}

__device__ __forceinline__ Normal2 normal_pair(uint64_t seed, uint64_t id, uint64_t t, uint32_t stream) {
  uint32_t w[4];
  philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu),
                (uint32_t)seed, (uint32_t)(seed >> 32), w);
  return normal_from_words(w);
}

__device__ __forceinline__ Normal2 normal_from_words(const uint32_t (&w)[4]) {
  const double u1 = u52_to_open_unit(w[0], w[1]);
  const double u2 = u52_to_open_unit(w[2], w[3]);
  const double R = sqrt_pos(FD_MUL(-2.0, log_unit(u1)));
  double sn, cs;
  sincospi_02(FD_ADD(u2, u2), sn, cs);
  Normal2 out;
  out.z0 = FD_MUL(R, cs);
  out.z1 = FD_MUL(R, sn);
  return out;
}

}  // namespace fishing
