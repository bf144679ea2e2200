// Round-1 Box-Muller (fdlibm-style log / Newton sqrt / sincospi polynomials, 58 FP64 instructions per pair of draws).
// Kept only as the A/B alternative of csrc/normal.cuh: build with -DFISHING_NORMAL_V1 (tools/ab_build.py) to time it.
// Same mathematical spec, results differ from normal.cuh in the last bits.
#pragma once
// (included by philox.cuh inside namespace fishing)

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

__device__ __forceinline__ Normal2 normal_from_words_v1(const uint32_t (&w)[4]) {
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

