// Host-side accuracy check of gym_fishing_b200/csrc/normal.cuh (the table-assisted Box-Muller) against an 80-bit long
// double evaluation of the same formulas.  Build + run:
//   g++ -O2 -std=c++17 -ffp-contract=off -o /tmp/normal_accuracy tools/normal_accuracy.cpp && /tmp/normal_accuracy
// The device's rsqrt seed is emulated (see normal.cuh); everything else is the code the kernels run.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>

#include "../gym_fishing_b200/csrc/normal.cuh"

static uint64_t sm_state = 0x9E3779B97F4A7C15ull;
static uint64_t splitmix() {
  uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Stats {
  double max_abs = 0, sum_ulp = 0, max_ulp = 0;
  long n = 0, exact = 0;
  void add(double got, long double ref, double floor = 0.0) {     // floor: measure in ulps of max(|ref|, floor)
    const double r = (double)ref;
    const double err = std::fabs((double)((long double)got - ref));
    const double at = std::fabs(r) < floor ? floor : r;
    const double ulp = std::fabs(std::nextafter(at, INFINITY) - at);
    if (err > max_abs) max_abs = err;
    sum_ulp += err / ulp;
    if (err / ulp > max_ulp) max_ulp = err / ulp;
    exact += (got == r);
    ++n;
  }
};

static void one(uint64_t k1, uint64_t k2, Stats& s) {
  const uint32_t w[4] = {(uint32_t)k1, (uint32_t)(k1 >> 32) | 0xABC00000u, (uint32_t)k2, (uint32_t)(k2 >> 32) | 0x12300000u};
  const fishing::Normal2 z = fishing::normal_from_words(w, fishing::HostTabs());
  const long double pi = 3.14159265358979323846264338327950288L;
  const long double u1 = ((long double)(k1 & 0xFFFFFFFFFFFFFull) + 0.5L) * 0x1p-52L;
  const long double u2 = ((long double)(k2 & 0xFFFFFFFFFFFFFull) + 0.5L) * 0x1p-52L;
  const long double R = sqrtl(-2.0L * logl(u1));
  // exact reduction to the nearest quarter turn first: cosl(2 pi u2) alone loses relative accuracy at the zero crossings
  const int q = (int)floorl(4.0L * u2 + 0.5L);
  const long double y = u2 - 0.25L * q;   // exact
  const long double c = cosl(2.0L * pi * y), sn = sinl(2.0L * pi * y);
  const long double C[4] = {c, -sn, -c, sn}, S[4] = {sn, c, -sn, -c};
  s.add(z.z0, R * C[q & 3]);
  s.add(z.z1, R * S[q & 3]);
}

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 4000000;
  Stats rnd, edge;
  for (long i = 0; i < N; ++i) one(splitmix() >> 12, splitmix() >> 12, rnd);
  // edges: both ends of the unit interval, every table-cell boundary of both tables, tiny and huge radii
  const unsigned long long top = (1ull << 52) - 1;
  for (int i = 0; i <= 256; ++i)
    for (int d = -2; d <= 2; ++d)
      for (int e = 0; e < 53; e += 4) {
        // u1 with exponent -1-e and significand at the boundary between entries i-1 and i (+- d ulps)
        long double m = 1.0L + ((long double)i - 0.5L) / 256.0L;
        if (m < 1.0L) m = 1.0L;
        long double u = ldexpl(m, -1 - e);
        long long k1 = (long long)(u * 0x1p52L) + d;
        if (k1 < 0) k1 = 0;
        if ((uint64_t)k1 > top) k1 = top;
        long long k2 = (long long)((((long double)i - 0.5L) / 256.0L) * 0x1p52L) + d;
        if (k2 < 0) k2 = 0;
        if ((uint64_t)k2 > top) k2 = top;
        one((uint64_t)k1, (uint64_t)k2, edge);
        one((uint64_t)k1, splitmix() >> 12, edge);
        one(splitmix() >> 12, (uint64_t)k2, edge);
      }
  for (uint64_t k : {0ull, 1ull, 2ull, top, top - 1, top / 2, top / 2 + 1})
    for (uint64_t j : {0ull, 1ull, top, top - 1, top / 4, top / 4 + 1, top / 2, top / 2 + 1, 3 * (top / 4), 3 * (top / 4) + 1}) one(k, j, edge);
  // exp_tab against expl over the range the tipping-point growth can produce, plus the hand-over to the library at |x| = 700
  Stats ex;
  for (long i = 0; i < N; ++i) {
    const double x = ((double)(splitmix() >> 11) * 0x1p-53 - 0.5) * (i % 4 == 0 ? 1400.0 : (i % 4 == 1 ? 2.0 : 40.0));
    ex.add(fishing::exp_tab(x, fishing::HostTabs()), expl((long double)x));
  }
  for (double x : {0.0, -0.0, 1e-300, -1e-300, 699.999, -699.999, 700.0, -700.0, 709.7, -745.0, -746.0, 710.0, 1399.0, -1399.0, 1401.0, -1401.0, 1e300, -1e300, -720.0, -740.0,
                   (double)INFINITY, -(double)INFINITY})
  {
    const double got = fishing::exp_tab(x, fishing::HostTabs()), want = (double)expl((long double)x);
    const double tol = want > 1e-300 ? 2.5e-16 * want : 5e-324;      // subnormal results: one extra rounding allowed
    if (!(std::fabs(got - want) <= tol || got == want)) { printf("exp(%g) = %g, want %g\n", x, got, want); return 2; }
  }
  if (!std::isnan(fishing::exp_tab(NAN, fishing::HostTabs()))) return 2;
  printf("exp   : n=%ld  mean=%.3f ulp  max=%.2f ulp  correctly rounded %.1f%%\n", ex.n, ex.sum_ulp / ex.n, ex.max_ulp, 100.0 * ex.exact / ex.n);
  if (ex.max_ulp > 1.1) return 3;
  // log_tab against logl: every binade in turn (the exponent term grows with |e|), around 1 (no cancellation allowed), cell
  // boundaries of the table, subnormals, and the special values
  Stats lg, lg1;
  for (long i = 0; i < N; ++i) {
    const double m = 1.0 + (double)(splitmix() >> 12) * 0x1p-52;
    const int e = (int)(splitmix() % 2046) - 1022;
    const double x = ldexp(m, i % 3 == 0 ? e : (int)(splitmix() % 12) - 6);
    lg.add(fishing::log_tab(x, fishing::HostTabs()), logl((long double)x), 0x1p-6);
    const double y = 1.0 + ((double)(splitmix() >> 11) * 0x1p-53 - 0.5) * (i % 2 ? 0x1p-6 : 0x1p-30);
    lg1.add(fishing::log_tab(y, fishing::HostTabs()), logl((long double)y));
  }
  for (int i = 0; i <= 256; ++i)
    for (int d = -2; d <= 2; ++d)
      for (int e : {-1022, -60, -1, 0, 1, 7, 1023}) {
        double x = ldexp(1.0 + (i - 0.5) / 256.0 < 1.0 ? 1.0 : 1.0 + (i - 0.5) / 256.0, e);
        for (int k = 0; k < (d < 0 ? -d : d); ++k) x = std::nextafter(x, d < 0 ? 0.0 : INFINITY);
        lg.add(fishing::log_tab(x, fishing::HostTabs()), logl((long double)x), 0x1p-6);
      }
  for (double x : {5e-324, 1e-310, 2.2250738585072009e-308, 2.2250738585072014e-308, 1.7976931348623157e308})
    lg.add(fishing::log_tab(x, fishing::HostTabs()), logl((long double)x), 0x1p-6);
  {
    const fishing::HostTabs ht;
    if (fishing::log_tab(1.0, ht) != 0.0 || std::signbit(fishing::log_tab(1.0, ht))) return 4;
    if (fishing::log_tab(0.0, ht) != -(double)INFINITY || fishing::log_tab(-0.0, ht) != -(double)INFINITY) return 4;
    if (fishing::log_tab((double)INFINITY, ht) != (double)INFINITY) return 4;
    if (!std::isnan(fishing::log_tab(-1.0, ht)) || !std::isnan(fishing::log_tab(NAN, ht)) || !std::isnan(fishing::log_tab(-(double)INFINITY, ht))) return 4;
  }
  printf("log   : n=%ld  mean=%.3f ulp  max=%.2f ulp  correctly rounded %.1f%%\n", lg.n, lg.sum_ulp / lg.n, lg.max_ulp, 100.0 * lg.exact / lg.n);
  printf("log~1 : n=%ld  max|err|=%.3e  mean=%.3f ulp  max=%.2f ulp  correctly rounded %.1f%%\n", lg1.n, lg1.max_abs, lg1.sum_ulp / lg1.n, lg1.max_ulp,
         100.0 * lg1.exact / lg1.n);
  // what the growth functions need is the ABSOLUTE error of mu = log(x) + ... (it goes through exp): <= 2.5 ulp of max(|result|, 2^-6)
  // (`log`), and a few 1e-18 (the cubic's truncation error at the cell edges, 2^-58.6) where the result is small
  if (lg.max_ulp > 2.5 || lg1.max_abs > 4e-18) return 5;
  for (const auto* s : {&rnd, &edge})
    printf("%s: n=%ld  max|err|=%.3e  mean=%.3f ulp  max=%.2f ulp  correctly rounded %.1f%%\n", s == &rnd ? "random" : "edges ", s->n,
           s->max_abs, s->sum_ulp / s->n, s->max_ulp, 100.0 * s->exact / s->n);
  // random draws: < 3e-15 absolute (the bound tests/test_gpu_parity.py asserts on the device); the hand-picked edge cases
  // include |z| > 8 (probability 1e-15 per draw), where 2 ulp is 3.6e-15
  return (rnd.max_abs < 3e-15 && edge.max_abs < 4e-15) ? 0 : 1;
}
