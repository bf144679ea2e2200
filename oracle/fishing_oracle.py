"""CPU oracle for the gym_fishing step/reset/policy hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy, the arithmetic of boettiger-lab/gym_fishing for fishing-v0/v1/v2/v4.
It is the *checker* for the CUDA engine in gym_fishing_b200/: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product package never does, and has no
CPU fallback.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here bit-for-bit (1e-12 for the
`exp` of fishing-v2 across numpy builds) against tests/golden/*, which tests/golden/make_golden.py recorded
by executing the unmodified reference in the authoring container.

Every float64 line below is ONE IEEE-754 operation, in the reference's evaluation order (no FMA; numpy
never fuses).  Citations are reference file:line.

Conventions: `kind` is one of "v0","v1","v2","v4".  All array arguments broadcast; scalars work too.
"""
from __future__ import annotations

import math

import numpy as np

KINDS = ("v0", "v1", "v2", "v4")
F64 = np.float64


# ----------------------------------------------------------------------------------------------------------------
# elementwise pieces (vectorised, bit-identical to the scalar reference)
# ----------------------------------------------------------------------------------------------------------------
def _py_min(a, b):
    """Python builtin min(a, b) == (b if b < a else a); a NaN b therefore returns a.
    gym_fishing/envs/base_fishing_env.py:117"""
    return np.where(b < a, b, a)


def _py_max(a, b):
    """Python builtin max(a, b) == (b if b > a else a).  base_fishing_env.py:74,118; models/policies.py:29"""
    return np.where(b > a, b, a)


def _np_maximum0(a):
    """np.maximum(a, 0.0): a if (a > 0 or isnan(a)) else +0.0.  base_fishing_env.py:125-132"""
    return np.maximum(a, 0.0)


def quota_from_action(kind, action, K, n_actions=100):
    """BaseFishingEnv.get_quota, base_fishing_env.py:135-147.

    v0 (Discrete): (action / n_actions) * K  -- Python int/int true division == IEEE double division.
    v1/v2/v4 (Box): np.clip(action, -1, 1)[0] then (a + 1) * K; NaN propagates through clip.
    Float64 actions only (a float32 action makes the reference compute the quota in float32; see
    quota_from_action_f32).
    """
    K = np.asarray(K, dtype=F64)
    if kind == "v0":
        a = np.asarray(action, dtype=F64)
        return (a / np.asarray(n_actions, dtype=F64)) * K
    a = np.asarray(action, dtype=F64)
    a = np.where(a < -1.0, -1.0, np.where(a > 1.0, 1.0, a))  # np.clip semantics, NaN stays NaN
    return (a + 1.0) * K


def quota_from_action_f32(action32, K, K_is_np_float64=False):
    """get_quota for a float32 Box action under numpy>=2 promotion rules (NEP 50): clip, +1 and *K are all
    float32 operations when K is a Python number; when K is an np.float64 (fishing-v4) the product is
    float64.  base_fishing_env.py:143-146.  Returned as float64 (exact widening)."""
    a = np.asarray(action32, dtype=np.float32)
    a = np.where(a < np.float32(-1), np.float32(-1), np.where(a > np.float32(1), np.float32(1), a)).astype(np.float32)
    t = (a + np.float32(1)).astype(np.float32)
    if K_is_np_float64:
        return t.astype(F64) * np.asarray(K, dtype=F64)
    return (t * np.asarray(K, dtype=np.float32)).astype(np.float32).astype(F64)


def action_from_quota(kind, quota, K, n_actions=100):
    """BaseFishingEnv.get_action, base_fishing_env.py:149-156.
    v0: round(quota * n_actions / K) -- Python round() is round-half-to-even, unbounded above.
    else: quota / K - 1."""
    q = np.asarray(quota, dtype=F64)
    K = np.asarray(K, dtype=F64)
    if kind == "v0":
        return np.rint((q * np.asarray(n_actions, dtype=F64)) / K)
    return q / K - 1.0


def population_from_state(s, K):
    """get_fish_population, base_fishing_env.py:158-160:  x = (s + 1) * K"""
    return (np.asarray(s, dtype=F64) + 1.0) * np.asarray(K, dtype=F64)


def state_from_population(x, K):
    """get_state, base_fishing_env.py:162-164:  s = x / K - 1"""
    return np.asarray(x, dtype=F64) / np.asarray(K, dtype=F64) - 1.0


def harvest(x, quota):
    """harvest_draw, base_fishing_env.py:112-119: h = min(x, quota); x1 = max(x - h, 0.0)"""
    h = _py_min(x, quota)
    x1 = _py_max(x - h, 0.0)
    return h, x1


def grow_logistic(x1, r, K, sigma, z):
    """BaseFishingEnv.population_draw, base_fishing_env.py:121-133 (v0, v1, v4).
    x1 + r*x1*(1 - x1/K) + x1*sigma*z, clipped below at 0; z is drawn even when sigma == 0."""
    t1 = r * x1
    t2 = x1 / K
    t3 = 1.0 - t2
    t4 = t1 * t3
    t5 = x1 + t4
    t6 = x1 * sigma
    t7 = t6 * z
    t8 = t5 + t7
    return _np_maximum0(t8)


def grow_tipping(x1, r, K, C, sigma, z):
    """FishingTippingEnv.population_draw, fishing_tipping_env.py:24-35 (v2).
    x1 * exp(r*(1 - x1/K)*(x1 - C) + x1*sigma*z), clipped below at 0 (noise is INSIDE the exponent)."""
    u1 = x1 / K
    u2 = 1.0 - u1
    u3 = r * u2
    u4 = x1 - C
    u5 = u3 * u4
    u6 = x1 * sigma
    u7 = u6 * z
    u8 = u5 + u7
    return _np_maximum0(x1 * np.exp(u8))


def step_vec(kind, s, action, z, years, *, r, K, sigma, C=0.5, n_actions=100, Tmax=100, quota=None):
    """One BaseFishingEnv.step for a batch (base_fishing_env.py:60-81).

    s: normalised state(s); action: float64 (Box) or integer-valued (Discrete); z: the N(0,1) draw this
    step consumes; years: years_passed BEFORE the step.  If `quota` is given it overrides get_quota
    (used for the float32-action path).  Returns (s', reward, done, years')."""
    s = np.asarray(s, dtype=F64)
    z = np.asarray(z, dtype=F64)
    r = np.asarray(r, dtype=F64)
    K = np.asarray(K, dtype=F64)
    sigma = np.asarray(sigma, dtype=F64)
    with np.errstate(all="ignore"):
        q = quota_from_action(kind, action, K, n_actions) if quota is None else np.asarray(quota, dtype=F64)  # :63
        x = population_from_state(s, K)                                  # :64
        h, x1 = harvest(x, q)                                            # :67
        if kind == "v2":
            x2 = grow_tipping(x1, r, K, np.asarray(C, dtype=F64), sigma, z)   # :68
        else:
            x2 = grow_logistic(x1, r, K, sigma, z)                       # :68
        s2 = state_from_population(x2, K)                                # :71
        reward = _py_max(h, 0.0)                                         # :74
        years2 = np.asarray(years) + 1                                   # :75
        done = (years2 > np.asarray(Tmax)) | (x2 <= 0.0)                 # :76-79
    return s2, reward, done, years2


def reset_state(kind, x0, K):
    """BaseFishingEnv.reset (base_fishing_env.py:83-91): s = x0 / K - 1.
    FishingModelError.reset (fishing_model_error.py:41-48): s = x0, UN-normalised (an upstream quirk)."""
    if kind == "v4":
        return np.asarray(x0, dtype=F64) + 0.0
    return np.asarray(x0, dtype=F64) / np.asarray(K, dtype=F64) - 1.0


def v4_draw(K_mean, r_mean, sigma_p, zK, zr):
    """fishing_model_error.py:37-38 / :42-43.  np.random.normal(m, s) == m + s*z; K is drawn first.
    np.clip(v, 0, 1e6): v<0 -> 0 ; v>1e6 -> 1e6 ; NaN stays."""
    def clip(v):
        return np.where(v < 0.0, 0.0, np.where(v > 1e6, 1e6, v))
    K = clip(np.asarray(K_mean, dtype=F64) + np.asarray(sigma_p, dtype=F64) * np.asarray(zK, dtype=F64))
    r = clip(np.asarray(r_mean, dtype=F64) + np.asarray(sigma_p, dtype=F64) * np.asarray(zr, dtype=F64))
    return K, r


# ----------------------------------------------------------------------------------------------------------------
# policies (gym_fishing/models/policies.py)
# ----------------------------------------------------------------------------------------------------------------
def bmsy(kind, *, r, K, C=0.5, K_is_np_float64=False, z=0.0):
    """BMSY, policies.py:51-67.  Grid = linspace(-1, 1, 10001, dtype=float32) (observation_space bounds and
    dtype); x_0 = (grid + 1) * K elementwise through get_fish_population; one *vectorised* population_draw with
    sigma forced to 0 (it still consumes one normal draw, multiplied by 0); S = x_0[argmax(growth - x_0)].

    For v0/v1/v2 K and r are Python numbers, so every operation stays in float32 (the noise term is
    float32_array * 0 * python_float).  For v4 K, r are np.float64 scalars and the grid is promoted to
    float64.  Returns a numpy scalar of the dtype the reference returns."""
    grid = np.linspace(np.float32(-1), np.float32(1), num=10001, dtype=np.float32)
    if kind == "v4" or K_is_np_float64:
        x0 = np.asarray([(g + 1) * F64(K) for g in grid])                 # np.float32 + int -> f32; * f64 -> f64
        rr = F64(r)
        KK = F64(K)
        one = 1.0
    else:
        x0 = np.asarray([(g + 1) * K for g in grid], dtype=np.float32)    # float32 throughout
        rr, KK, one = r, K, 1.0
    with np.errstate(all="ignore"):
        if kind == "v2":
            g = np.maximum(x0 * np.exp(rr * (1 - x0 / KK) * (x0 - C) + x0 * 0 * float(z)), 0)
        else:
            g = np.maximum(x0 + rr * x0 * (one - x0 / KK) + x0 * 0 * float(z), 0.0)
        growth = g - x0
    return x0[np.argmax(growth)]


def msy_quota(kind, S, *, r, K, C=0.5, z=0.0):
    """msy.__init__, policies.py:4-14: one scalar population_draw at x = S with sigma = 0, minus S.
    S keeps the dtype BMSY returned (np.float32 for v0/v1/v2), so the result is float32 there."""
    with np.errstate(all="ignore"):
        if kind == "v2":
            g = np.maximum(S * np.exp(r * (1 - S / K) * (S - C) + S * 0 * float(z)), 0)
        else:
            g = np.maximum(S + r * S * (1.0 - S / K) + S * 0 * float(z), 0.0)
    return g - S


def escapement_action(kind, s, S, K, n_actions=100):
    """escapement.predict, policies.py:27-31: x = (s+1)*K; quota = max(x - S, 0.0); action = get_action(quota).
    S enters as float64(S) (exact widening of the float32 BMSY value)."""
    x = population_from_state(s, K)
    q = _py_max(x - F64(S), 0.0)
    return action_from_quota(kind, q, K, n_actions)


# ----------------------------------------------------------------------------------------------------------------
# batched rollouts (what the fused CUDA kernel is checked against)
# ----------------------------------------------------------------------------------------------------------------
def rollout_vec(kind, T, N, *, r, K, sigma, C=0.5, x0=0.75, n_actions=100, Tmax=100, policy="actions",
                actions=None, S=None, msy_q=None, noise=None, auto_reset=False, reset_noise=None,
                K_mean=None, r_mean=None, sigma_p=None, s0=None, record=True):
    """Lock-step rollout of N envs for T steps.

    policy: "actions" (actions[T, N] supplied), "escapement" (needs S) or "msy" (needs msy_q = constant quota).
    noise: [T, N] standard normals (zeros if None).  auto_reset: SB3 DummyVecEnv semantics -- on done the
    terminal observation is kept aside and the env is reset (v4 redraws K, r from reset_noise[t+1, 0/1, :]).
    Returns a dict with per-step obs/reward/done(/term/action) when record=True plus per-env `ret` (sum of
    rewards while the env had not yet been done, i.e. the first episode), `done_step` (first t with done, or
    -1) and the final state."""
    r = np.broadcast_to(np.asarray(r, dtype=F64), (N,)).copy()
    K = np.broadcast_to(np.asarray(K, dtype=F64), (N,)).copy()
    s = reset_state(kind, x0, K) if s0 is None else np.asarray(s0, dtype=F64).copy()
    s = np.broadcast_to(s, (N,)).copy()
    years = np.zeros(N, dtype=np.int64)
    ret = np.zeros(N)
    total = np.zeros(N)
    done_step = np.full(N, -1, dtype=np.int64)
    alive = np.ones(N, dtype=bool)
    out = {}
    if record:
        out = dict(obs=np.zeros((T, N)), reward=np.zeros((T, N)), done=np.zeros((T, N), dtype=np.uint8),
                   term=np.full((T, N), np.nan), action=np.zeros((T, N)))
    for t in range(T):
        z = np.zeros(N) if noise is None else noise[t]
        if policy == "actions":
            a = actions[t]
        elif policy == "escapement":
            a = escapement_action(kind, s, S, K, n_actions)
        elif policy == "msy":
            a = action_from_quota(kind, np.broadcast_to(F64(msy_q), (N,)), K, n_actions)
        else:
            raise ValueError(policy)
        s2, rew, done, years = step_vec(kind, s, a, z, years, r=r, K=K, sigma=sigma, C=C, n_actions=n_actions,
                                        Tmax=Tmax)
        total = total + rew
        ret = np.where(alive, ret + rew, ret)
        first = alive & done
        done_step[first] = t
        alive &= ~done
        if record:
            out["reward"][t] = rew
            out["done"][t] = done
            out["action"][t] = a
        if auto_reset and done.any():
            if record:
                out["term"][t] = np.where(done, s2, np.nan)
            if kind == "v4":
                Kn, rn = v4_draw(K_mean, r_mean, sigma_p, reset_noise[t + 1, 0], reset_noise[t + 1, 1])
                K = np.where(done, Kn, K)
                r = np.where(done, rn, r)
            s2 = np.where(done, reset_state(kind, x0, K), s2)
            years = np.where(done, 0, years)
        s = s2
        if record:
            out["obs"][t] = s
    out.update(ret=ret, total=total, done_step=done_step, final_state=s, years=years, K=K, r=r)
    return out


# ----------------------------------------------------------------------------------------------------------------
# scalar "reference-shaped" port: same per-step numpy-scalar call pattern as the reference, used ONLY as the
# CPU baseline that bench.py times (cpu_baseline.kind == "port") and as a second, independent oracle in tests.
# ----------------------------------------------------------------------------------------------------------------
class ScalarPortEnv:
    """A single environment stepping one Python call at a time, with the same numpy-scalar operations per step
    as BaseFishingEnv.step (np.clip on a 1-element array, Python min/max, np.maximum, np.array allocation and
    one np.random.normal call), so that its speed is representative of the reference's CPU path
    (base_fishing_env.py:60-81, :112-164; fishing_tipping_env.py:24-35; fishing_model_error.py:35-48)."""

    def __init__(self, kind="v1", r=0.3, K=1, sigma=0.0, C=0.5, n_actions=100, init_state=0.75, Tmax=100,
                 sigma_p=0.1, normal=None):
        self.kind, self.sigma, self.C, self.n_actions = kind, sigma, C, n_actions
        self.init_state, self.Tmax, self.sigma_p = init_state, Tmax, sigma_p
        self.K_mean, self.r_mean = K, r
        self.normal = normal if normal is not None else np.random.normal
        self.lo = np.array([-1], dtype=np.float32)
        self.hi = np.array([1], dtype=np.float32)
        self.K, self.r = K, r
        if kind == "v4":
            self._draw()
        self.years_passed = 0
        self.state = np.array([init_state / self.K - 1])

    def _draw(self):
        self.K = np.clip(self.normal(self.K_mean, self.sigma_p), 0, 1e6)
        self.r = np.clip(self.normal(self.r_mean, self.sigma_p), 0, 1e6)

    def reset(self):
        if self.kind == "v4":
            self._draw()
            self.state = np.array([self.init_state])
        else:
            self.state = np.array([self.init_state / self.K - 1])
        self.years_passed = 0
        return self.state

    def step(self, action):
        K = self.K
        if self.kind == "v0":
            quota = (action / self.n_actions) * K
        else:
            quota = (np.clip(action, self.lo, self.hi)[0] + 1) * K
        x = (self.state[0] + 1) * K
        h = min(x, quota)
        x = max(x - h, 0.0)
        if self.kind == "v2":
            x = np.maximum(x * np.exp(self.r * (1 - x / K) * (x - self.C) + x * self.sigma * self.normal(0, 1)), 0)
        else:
            x = np.maximum(x + self.r * x * (1.0 - x / K) + x * self.sigma * self.normal(0, 1), 0.0)
        self.state = np.array([x / K - 1])
        reward = max(h, 0.0)
        self.years_passed += 1
        done = bool(self.years_passed > self.Tmax)
        if x <= 0.0:
            done = True
        return self.state, reward, done, {}


def simulate_table(kind, policy, *, reps, r, K, sigma, C=0.5, x0=0.75, n_actions=100, Tmax=100, S=None, msy_q=None,
                   noise=None, f32_actions=False):
    """simulate_mdp, gym_fishing/envs/shared_env.py:29-54, as rows [t, x, quota, reward, rep].
    Row t records the state BEFORE step t and the quota/reward of step t-1 (one-step lag, :37-38,49);
    the loop runs at most Tmax steps and breaks after a done step (:51).  `noise` is consumed one per step."""
    rows = []
    zi = 0
    for rep in range(reps):
        s = float(reset_state(kind, x0, K))
        years = 0
        quota = 0.0
        reward = 0.0
        for t in range(Tmax):
            x = float(population_from_state(s, K))
            rows.append([t, x, quota, reward, rep])
            if policy == "escapement":
                a = escapement_action(kind, s, S, K, n_actions)
            else:
                a = action_from_quota(kind, msy_q, K, n_actions)
            z = 0.0 if noise is None else noise[zi]
            zi += 1
            if f32_actions and kind != "v0":
                a32 = np.float32(a)
                q = quota_from_action_f32(a32, K)
                s2, rew, done, years = step_vec(kind, s, a32, z, years, r=r, K=K, sigma=sigma, C=C,
                                                n_actions=n_actions, Tmax=Tmax, quota=q)
            else:
                q = quota_from_action(kind, a, K, n_actions)
                s2, rew, done, years = step_vec(kind, s, a, z, years, r=r, K=K, sigma=sigma, C=C,
                                                n_actions=n_actions, Tmax=Tmax)
            s, reward, quota = float(s2), float(rew), float(q)
            if bool(done):
                break
    return rows


def hexf(v):
    return float(v).hex()


def unhex(h):
    return float.fromhex(h) if isinstance(h, str) else float(h)


__all__ = [n for n in dir() if not n.startswith("_")] + ["math"]


# ----------------------------------------------------------------------------------------------------------------
# "next row" f1: the log-normal growth family fishing-v5..v9, v10 (non-stationary r) and v11 (model uncertainty),
# gym_fishing/envs/growth_models.py:6-270.  Same step skeleton (base_fishing_env.py:60-81) with population_draw
# replaced by x2 = max(0, lognormal(mu(x1), sigma)) = max(0, exp(mu + sigma*z)).  These functions read the env's
# `params` DICT (growth_models.py:24,39,69,105,120), not the env attributes, so the attribute `sigma` that the policies
# zero (policies.py:9-13,60-65) has no effect on them.
# ----------------------------------------------------------------------------------------------------------------
LN_KINDS = ("v5", "v6", "v7", "v8", "v9", "v10", "v11")
MODEL_NAMES = ("allen", "beverton_holt", "myers", "may", "ricker")     # growth_models.py:156
MODEL_OF_KIND = {"v5": "allen", "v6": "beverton_holt", "v7": "may", "v8": "myers", "v9": "ricker", "v10": "beverton_holt"}
# constructor defaults, growth_models.py:7-16,29-31,44-53,76-87,110-112,125-134,157-179
LN_DEFAULTS = {
    "v5": dict(r=0.3, K=1, C=0.5, sigma=0.0, x0=0.75),
    "v6": dict(r=0.3, K=1, sigma=0.0, x0=0.75),
    "v7": dict(r=0.7, K=1.5, M=1.5, q=3, b=0.15, sigma=0.0, a=0.2, x0=0.75),
    "v8": dict(r=1.0, K=1.0, M=1.0, theta=3.0, sigma=0.0, x0=1.5),
    "v9": dict(r=0.3, K=1, sigma=0.0, x0=0.75),
    "v10": dict(r=0.8, K=1, sigma=0.0, alpha=-0.007, x0=0.75),
}
V11_PARAMS = {
    "allen": dict(r=0.3, K=1.0, sigma=0.0, C=0.5, x0=0.75),
    "beverton_holt": dict(r=0.3, K=1, sigma=0.0, x0=0.75),
    "myers": dict(r=1.0, K=1.0, M=1.0, theta=3.0, sigma=0.0, x0=1.5),
    "may": dict(r=0.7, K=1.5, M=1.5, q=3, b=0.15, sigma=0.0, a=0.2, x0=0.75),
    "ricker": dict(r=0.3, K=1, sigma=0.0, x0=0.75),
}


def lognormal_mu(model, x, p):
    """mu(x) of the five growth functions, growth_models.py:208-261, one IEEE operation per line of the reference's
    expression tree (np.log / np.power are the only non-exact operations)."""
    x = np.asarray(x, dtype=F64)
    with np.errstate(all="ignore"):
        if model == "allen":                                   # :208-217
            return np.log(x) + F64(p["r"]) * (1 - x / F64(p["K"])) * F64(1 - p["C"]) / F64(p["K"])
        if model == "beverton_holt":                           # :220-226
            xc = np.where(x < 0.0, 0.0, x)                     # np.clip(x, 0, inf); NaN stays
            rc = F64(p["r"]) if p["r"] >= 0 else F64(0.0)
            A = rc + 1
            B = (F64(p["K"]) if p["K"] >= 0 else F64(0.0)) / rc
            return np.log(A) + np.log(xc) - np.log(1 + xc / B)
        if model == "may":                                     # :229-242
            q = F64(p["q"])
            xq = np.power(x, q)
            e = x + x * F64(p["r"]) * (1 - x / F64(p["M"])) - F64(p["a"]) * xq / (xq + np.power(F64(p["b"]), q))
            return np.log(e)
        if model == "myers":                                   # :247-255
            A = F64(p["r"]) + 1
            th = F64(p["theta"])
            return np.log(A) + th * np.log(x) - np.log(1 + np.power(x, th) / F64(p["M"]))
        if model == "ricker":                                  # :258-261
            return np.log(x) + F64(p["r"]) * (1 - x / F64(p["K"]))
    raise ValueError(model)


def grow_lognormal(model, x1, p, z):
    """np.maximum(0, np.random.lognormal(mu, sigma)); the legacy generator computes exp(mu + sigma * gauss)."""
    with np.errstate(all="ignore"):
        mu = lognormal_mu(model, x1, p)
        return np.maximum(0, np.exp(mu + F64(p["sigma"]) * np.asarray(z, dtype=F64)))


def step_ln(model, s, action, z, years, p, *, K_state, Tmax=100):
    """BaseFishingEnv.step with a log-normal growth model.  K_state is the env attribute K used by get_quota /
    get_fish_population / get_state (for fishing-v11 it is the base default 1 whatever the model, growth_models.py:183-189)."""
    s = np.asarray(s, dtype=F64)
    K = F64(K_state)
    with np.errstate(all="ignore"):
        q = quota_from_action("v1", action, K)
        x = population_from_state(s, K)
        h, x1 = harvest(x, q)
        x2 = grow_lognormal(model, x1, p, z)
        s2 = state_from_population(x2, K)
        reward = _py_max(h, 0.0)
        years2 = np.asarray(years) + 1
        done = (years2 > np.asarray(Tmax)) | (x2 <= 0.0)
    return s2, reward, done, years2
