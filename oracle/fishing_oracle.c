/*
 * C restatement of the gym_fishing step() hot path.  TEST INFRASTRUCTURE ONLY (see oracle/fishing_oracle.py, which is
 * the primary oracle and is pinned to the golden vectors of the unmodified reference; tests/test_oracle_c.py pins this
 * file to the numpy oracle bit for bit).  It exists to give bench.py a "best a CPU can do" figure next to the
 * reference-shaped Python port: the same arithmetic, one IEEE operation per reference operation, compiled with
 * -ffp-contract=off (no FMA), parallelised over environments with OpenMP.
 *
 * Reference lines: gym_fishing/envs/base_fishing_env.py:60-81 (step), :112-119 (harvest_draw), :121-133
 * (population_draw), :135-147 (get_quota), :158-164 (state maps); fishing_tipping_env.py:24-35 (v2 growth).
 */
#include <math.h>
#include <stdint.h>

static inline double py_min(double a, double b) { return (b < a) ? b : a; }
static inline double py_max(double a, double b) { return (b > a) ? b : a; }
static inline double np_max0(double a) { return (a > 0.0 || a != a) ? a : 0.0; }

/* kind: 0 = v0 (discrete, action given as double), 1 = v1, 2 = v2 (tipping), 4 = v4 (per-env r, K) */
void oracle_step(int kind, int64_t n, const double* r_env, const double* K_env, double r, double K, double sigma, double C,
                 double n_actions, int32_t Tmax, int auto_reset, double s_reset, const double* action, const double* z,
                 double* state, int32_t* years, double* reward, uint8_t* done) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const double Ki = (kind == 4) ? K_env[i] : K, ri = (kind == 4) ? r_env[i] : r;
    double quota;
    if (kind == 0) {
      quota = (action[i] / n_actions) * Ki;
    } else {
      double a = action[i];
      a = (a < -1.0) ? -1.0 : ((a > 1.0) ? 1.0 : a);
      quota = (a + 1.0) * Ki;
    }
    const double x = (state[i] + 1.0) * Ki;
    const double h = py_min(x, quota);
    const double x1 = py_max(x - h, 0.0);
    double x2;
    if (kind == 2) {
      const double u8 = (ri * (1.0 - x1 / Ki)) * (x1 - C) + (x1 * sigma) * z[i];
      x2 = np_max0(x1 * exp(u8));
    } else {
      const double t5 = x1 + (ri * x1) * (1.0 - x1 / Ki);
      x2 = np_max0(t5 + (x1 * sigma) * z[i]);
    }
    double s2 = x2 / Ki - 1.0;
    reward[i] = py_max(h, 0.0);
    int32_t y = years[i] + 1;
    const uint8_t d = (y > Tmax) || (x2 <= 0.0);
    if (auto_reset && d && kind != 4) {
      s2 = s_reset;
      y = 0;
    }
    state[i] = s2;
    years[i] = y;
    done[i] = d;
  }
}

/* T lock-steps over the batch with a constant action table [T, n] and noise table [T, n] (auto-reset, not v4) */
void oracle_rollout(int kind, int64_t n, int64_t T, double r, double K, double sigma, double C, double n_actions, int32_t Tmax,
                    double s_reset, const double* actions, const double* z, double* state, int32_t* years, double* reward,
                    uint8_t* done) {
  for (int64_t t = 0; t < T; ++t)
    oracle_step(kind, n, 0, 0, r, K, sigma, C, n_actions, Tmax, 1, s_reset, actions + t * n, z + t * n, state, years, reward, done);
}
