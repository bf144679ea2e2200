#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

Run in the authoring container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (boettiger-lab/gym_fishing) is pure Python but imports `gym` and `matplotlib`, neither of
which is installed here, so the throw-away stand-ins under oracle/refstub/ are put on sys.path first
(SURVEY.md section 8c).  Nothing from the reference is copied: only its *outputs* are recorded.

Noise protocol: the reference looks `np.random.normal` up at call time
(gym_fishing/envs/base_fishing_env.py:130, fishing_tipping_env.py:31, fishing_model_error.py:37-43), so it
is replaced by an injector that returns PYTHON floats `loc + scale*z` for z taken in order from a fixed
array.  (A np.float64 return would change BMSY's float32 promotion under NEP 50.)

Outputs:
  kat.json         -- known-answer vectors (hex floats), Appendix B of SURVEY.md re-derived here
  traj_<kind>.npz  -- random trajectories: per-trajectory params, actions, noise, and the reference's
                      obs / reward / done for every step (no auto-reset; stepping continues past done)
  autoreset.npz    -- SB3-DummyVecEnv-style auto-reset trajectories (reset() called on done), incl. v4 redraws
  policy.json      -- BMSY / msy constants for a grid of parameter sets
  simulate.json    -- env.simulate(escapement) tables (simulate_mdp, shared_env.py:29-54)
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refstub"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

import gym  # noqa: E402  (the stub)
import gym_fishing  # noqa: E402,F401  (the unmodified reference; registers the env ids)
from gym_fishing.models.policies import BMSY, escapement, msy  # noqa: E402

_REAL_NORMAL = np.random.normal
_REAL_LOGNORMAL = np.random.lognormal
_REAL_CHOICE = np.random.choice


class Injector:
    """Replaces numpy.random.normal; hands out loc + scale*Z[i] as Python floats, counting the calls."""

    def __init__(self, z, choices=None):
        self.z = np.asarray(z, dtype=np.float64)
        self.i = 0
        self.choices = choices
        self.ci = 0

    def __call__(self, loc=0.0, scale=1.0, size=None):
        assert size is None
        v = float(self.z[self.i])
        self.i += 1
        return float(loc + scale * v)

    def lognormal(self, mean=0.0, sigma=1.0, size=None):
        """np.random.lognormal of the legacy generator is exp(mean + sigma * gauss) (one gauss per output element)."""
        assert size is None
        m = np.asarray(mean, dtype=np.float64)
        if m.ndim == 0:
            v = float(self.z[self.i])
            self.i += 1
            return float(np.exp(np.float64(float(m) + sigma * v)))
        z = self.z[self.i:self.i + m.size].reshape(m.shape)
        self.i += m.size
        return np.exp(m + sigma * z)

    def choice(self, a, size=None):
        """np.random.choice(models) for fishing-v11: consumes the next entry of `choices`."""
        assert size is None
        v = a[self.choices[self.ci]]
        self.ci += 1
        return v

    def __enter__(self):
        np.random.normal = self
        np.random.lognormal = self.lognormal
        if getattr(self, "choices", None) is not None:
            np.random.choice = self.choice
        return self

    def __exit__(self, *a):
        np.random.normal = _REAL_NORMAL
        np.random.lognormal = _REAL_LOGNORMAL
        np.random.choice = _REAL_CHOICE


def hx(v):
    return float(v).hex()


def wrap(kind, a):
    """Action as the reference's callers pass it: int for Discrete, 1-element float64 array for Box."""
    if kind == "v0":
        return int(a)
    return np.array([a], dtype=np.float64)


Z = np.random.default_rng(2024).standard_normal(4096)


# --------------------------------------------------------------------------------------------- KATs
def kat():
    out = {"Z_head": [hx(z) for z in Z[:3]]}

    # KAT-1: v0 Tmax=1000 escapement
    with Injector(np.zeros(8)):
        env = gym.make("fishing-v0", Tmax=1000)
        pol = escapement(env)
    with Injector(np.zeros(2000)):
        obs = env.reset()
        acts, tot = [], 0.0
        for t in range(1000):
            a, _ = pol.predict(obs)
            obs, r, d, _ = env.step(a)
            acts.append(int(a))
            tot += r
        k1 = dict(S=hx(pol.S), actions_head=acts[:6], actions_tail=acts[-6:],
                  wsum=int(sum(a * (t + 1) for t, a in enumerate(acts))), final_s=hx(obs[0]),
                  final_x=hx(env.fish_population), sumR=hx(tot), done=bool(d), years=int(env.years_passed),
                  actions=acts)
        a, _ = pol.predict(obs)
        _, _, d2, _ = env.step(a)
        k1["done_1001"] = bool(d2)
    out["kat1"] = k1

    # KAT-2: v1 defaults, action -0.85 x 101
    with Injector(np.zeros(200)):
        env = gym.make("fishing-v1")
        env.reset()
        ss, rr, dd = [], [], []
        for t in range(101):
            obs, r, d, _ = env.step(np.array([-0.85]))
            ss.append(hx(obs[0])); rr.append(hx(r)); dd.append(bool(d))
    out["kat2"] = dict(s=ss, r=rr, done=dd, sumR=hx(sum(float.fromhex(v) for v in rr)))

    # KAT-2b
    with Injector(np.zeros(10)):
        env = gym.make("fishing-v1")
        env.reset()
        o1 = env.step(np.array([0.1])); o1 = (hx(o1[0][0]), hx(o1[1]), bool(o1[2]))
        o2 = env.step(np.array([-1.0])); o2 = (hx(o2[0][0]), hx(o2[1]), bool(o2[2]))
    out["kat2b"] = [o1, o2]

    # KAT-3: v2 sigma=0 (the only numeric assertions upstream, tests/test-envs.py:93-106)
    with Injector(np.zeros(10)):
        env = gym.make("fishing-v2", sigma=0, init_state=0.75)
        env.reset()
        obs, *_ = env.step(np.array([env.get_action(0)]))
        xa = env.get_fish_population(obs)
        env.init_state = 0.3
        env.reset()
        obs, *_ = env.step(np.array([env.get_action(0)]))
        xb = env.get_fish_population(obs)
    out["kat3"] = dict(x_075=hx(xa), x_030=hx(xb))

    # KAT-3b: v2 sigma=0 Tmax=1000 escapement
    with Injector(np.zeros(2000)):
        env = gym.make("fishing-v2", sigma=0, Tmax=1000)
        pol = escapement(env)
        obs = env.reset()
        tot = 0.0
        for t in range(1000):
            a, _ = pol.predict(obs)
            obs, r, d, _ = env.step(np.array([a]))
            tot += r
    out["kat3b"] = dict(S=hx(pol.S), final_s=hx(obs[0]), sumR=hx(tot), done=bool(d))

    # KAT-5: v1 sigma=0.05 constant action, noise Z[0:1000]
    with Injector(Z):
        env = gym.make("fishing-v1", sigma=0.05, Tmax=1000)
        obs = env.reset()
        tot = 0.0
        for t in range(1000):
            obs, r, d, _ = env.step(np.array([-0.95]))
            tot += r
    out["kat5"] = dict(final_s=hx(obs[0]), sumR=hx(tot), done=bool(d))

    # KAT-6 / KAT-7: injector installed before escapement(env) -> BMSY eats Z[0]
    for name, envid in (("kat6", "fishing-v0"), ("kat7", "fishing-v2")):
        with Injector(Z):
            env = gym.make(envid, sigma=0.05, Tmax=1000)
            pol = escapement(env)
            obs = env.reset()
            tot, acts = 0.0, []
            for t in range(1000):
                a, _ = pol.predict(obs)
                acts.append(float(a))
                obs, r, d, _ = env.step(wrap("v0" if envid == "fishing-v0" else "v2", a))
                tot += r
        rec = dict(S=hx(pol.S), final_s=hx(obs[0]), sumR=hx(tot), done=bool(d))
        if envid == "fishing-v0":
            rec["wsum"] = int(sum(int(a) * (t + 1) for t, a in enumerate(acts)))
        out[name] = rec

    # KAT-8: v4, injector before make
    with Injector(Z):
        env = gym.make("fishing-v4", sigma=0.05, Tmax=1000)
        rec = dict(ctor_K=hx(env.K), ctor_r=hx(env.r))
        obs = env.reset()
        rec.update(reset_K=hx(env.K), reset_r=hx(env.r), reset_obs=hx(obs[0]))
        tot = 0.0
        for t in range(1000):
            obs, r, d, _ = env.step(np.array([-0.95]))
            tot += r
            if t == 0:
                rec.update(s0=hx(obs[0]), r0=hx(r))
            if d:
                rec.update(done_t=t)
                break
        rec.update(sumR=hx(tot))
    out["kat8"] = rec

    # Edge cases: NaN / out-of-range actions, stepping after done, negative/oversized discrete actions
    edge = []
    cases = [
        ("fishing-v1", {}, [float("nan"), -0.9, -0.9]),
        ("fishing-v1", {}, [5.0, -0.5, 0.0]),
        ("fishing-v1", {}, [-7.0, -7.0, 1.0, -1.0, -1.0]),
        ("fishing-v1", {"K": 2.5, "init_state": 1.0, "r": 0.9}, [-0.8, -0.7, -0.99, 0.3, -1.0]),
        ("fishing-v1", {"Tmax": 3}, [-0.95] * 6),
        ("fishing-v0", {}, [-10, 250, 3]),
        ("fishing-v0", {"n_actions": 7, "K": 1.7}, [1, 2, 0, 6, 9]),
        ("fishing-v2", {"C": 0.2, "sigma": 0.0}, [-0.97, float("nan"), -0.97]),
        ("fishing-v2", {"K": 3.0, "init_state": 0.1, "C": 0.5}, [-1.0] * 8),
    ]
    for envid, kw, actions in cases:
        kind = "v" + envid.split("-v")[1]
        with Injector(Z[100:200]):
            env = gym.make(envid, **kw)
            obs0 = env.reset()
            steps = []
            for a in actions:
                obs, r, d, _ = env.step(wrap(kind, a))
                steps.append((hx(obs[0]), hx(r), bool(d)))
        edge.append(dict(env=envid, kw=kw, actions=[hx(a) if kind != "v0" else int(a) for a in actions],
                         obs0=hx(obs0[0]), steps=steps, noise_offset=100))
    out["edge"] = edge
    return out


# --------------------------------------------------------------------------- random trajectories
def trajectories(kind, M, T, seed, auto_reset=False):
    """M independent reference envs x T steps with per-trajectory parameters; records everything."""
    rng = np.random.default_rng(seed)
    envid = {"v0": "fishing-v0", "v1": "fishing-v1", "v2": "fishing-v2", "v4": "fishing-v4"}[kind]
    P = dict(r=np.zeros(M), K=np.zeros(M), sigma=np.zeros(M), C=np.zeros(M), x0=np.zeros(M),
             Tmax=np.zeros(M, dtype=np.int32), n_actions=np.zeros(M, dtype=np.int32), sigma_p=np.zeros(M))
    A = np.zeros((T, M)) if kind != "v0" else np.zeros((T, M), dtype=np.int32)
    NZ = np.zeros((T, M))          # z consumed by step t
    RZ = np.zeros((T + 1, 2, M))   # (z_K, z_r) consumed by the reset that follows step t-1 (v4; index 0 = initial)
    OBS0 = np.zeros(M)
    OBS = np.zeros((T, M)); REW = np.zeros((T, M)); DONE = np.zeros((T, M), dtype=np.uint8)
    TERM = np.full((T, M), np.nan)  # terminal observation when auto-reset fired
    KE = np.zeros((T + 1, M)); RE = np.zeros((T + 1, M))  # per-episode K, r in force for step t (index t)
    for m in range(M):
        default = m % 4 == 0
        r = 0.3 if default else float(rng.uniform(0.05, 1.2))
        K = 1.0 if (default or m % 4 == 1) else float(rng.uniform(0.5, 4.0))
        sigma = float(rng.choice([0.0, 0.01, 0.05, 0.2]))
        C = 0.5 if default else float(rng.uniform(0.05, 0.6)) * K
        x0 = 0.75 * K if default else float(rng.uniform(0.05, 1.6)) * K
        Tmax = int(rng.choice([5, 17, 100, 1000]))
        n_actions = 100 if default else int(rng.choice([3, 10, 100, 255]))
        sigma_p = 0.1 if default else float(rng.uniform(0.0, 0.4))
        for k, v in dict(r=r, K=K, sigma=sigma, C=C, x0=x0, Tmax=Tmax, n_actions=n_actions,
                         sigma_p=sigma_p).items():
            P[k][m] = v
        if kind == "v0":
            style = m % 3
            if style == 0:
                acts = rng.integers(0, max(2, n_actions // 8), size=T)
            elif style == 1:
                acts = rng.integers(0, n_actions, size=T)
            else:
                acts = rng.integers(-3, 2 * n_actions + 5, size=T)
            A[:, m] = acts
        else:
            style = m % 3
            if style == 0:
                acts = rng.uniform(-1.0, -0.8, size=T)
            elif style == 1:
                acts = rng.uniform(-1.0, 1.0, size=T)
            else:
                acts = rng.uniform(-1.3, -0.6, size=T)
            A[:, m] = acts
        z_step = rng.standard_normal(T)
        z_reset = rng.standard_normal((T + 1, 2))
        NZ[:, m] = z_step
        RZ[:, :, m] = z_reset

        # construction (v4 eats 2 draws we do not care about), then the initial reset
        if kind == "v0":
            kw = dict(r=r, K=K, sigma=sigma, n_actions=n_actions, init_state=x0, Tmax=Tmax)
        elif kind == "v1":
            kw = dict(r=r, K=K, sigma=sigma, init_state=x0, Tmax=Tmax)
        elif kind == "v2":
            kw = dict(r=r, K=K, C=C, sigma=sigma, init_state=x0, Tmax=Tmax)
        else:
            kw = dict(K_mean=K, r_mean=r, sigma=sigma, sigma_p=sigma_p, init_state=x0, Tmax=Tmax)
        with Injector(np.zeros(2)):
            env = gym.make(envid, **kw)
        with Injector(z_reset[0]):
            obs = env.reset()
        OBS0[m] = obs[0]
        KE[0, m] = env.K; RE[0, m] = env.r
        for t in range(T):
            with Injector(z_step[t:t + 1]) as inj:
                obs, rew, done, _ = env.step(wrap(kind, A[t, m]))
                assert inj.i == 1
            DONE[t, m] = done
            REW[t, m] = rew
            if auto_reset and done:
                TERM[t, m] = obs[0]
                with Injector(z_reset[t + 1]):
                    obs = env.reset()
            OBS[t, m] = obs[0]
            KE[t + 1, m] = env.K; RE[t + 1, m] = env.r
    return dict(actions=A, noise=NZ, reset_noise=RZ, obs0=OBS0, obs=OBS, reward=REW, done=DONE, term=TERM,
                K_env=KE, r_env=RE, **{"p_" + k: v for k, v in P.items()})


# ------------------------------------------------------------------------------- policy constants
def policy_table():
    rows = []
    sets = [
        ("fishing-v0", {}), ("fishing-v1", {}), ("fishing-v2", {}),
        ("fishing-v0", {"r": 0.8, "K": 2.0}), ("fishing-v1", {"r": 0.1, "K": 0.7}),
        ("fishing-v1", {"r": 1.1, "K": 3.3, "sigma": 0.1}),
        ("fishing-v2", {"r": 0.9, "K": 2.0, "C": 0.4}), ("fishing-v2", {"r": 0.3, "K": 1.0, "C": 0.1, "sigma": 0.05}),
        ("fishing-v4", {}), ("fishing-v4", {"K_mean": 2.0, "r_mean": 0.5, "sigma_p": 0.2}),
    ]
    for envid, kw in sets:
        z = Z[300:320]
        with Injector(z) as inj:
            env = gym.make(envid, **kw)
            n_ctor = inj.i
            K_ctor, r_ctor = float(env.K), float(env.r)
            S = BMSY(env)
            n_bmsy = inj.i - n_ctor
            K_after, r_after = float(env.K), float(env.r)
        with Injector(z) as inj:
            env = gym.make(envid, **kw)
            pol = msy(env)
            n_msy = inj.i - n_ctor
        rows.append(dict(env=envid, kw=kw, S=hx(S), S_dtype=str(np.asarray(S).dtype), msy=hx(pol.msy),
                         msy_dtype=str(np.asarray(pol.msy).dtype), draws_ctor=n_ctor, draws_bmsy=n_bmsy,
                         draws_msy=n_msy, K_ctor=hx(K_ctor), r_ctor=hx(r_ctor), K_bmsy=hx(K_after),
                         r_bmsy=hx(r_after), noise_offset=300,
                         msy_action=hx(pol.predict(np.array([0.0]))[0])))
    return rows


# ----------------------------------------------------------------------------------- simulate_mdp
def simulate_tables():
    out = []
    for envid, kw, which, reps in [("fishing-v0", {}, "escapement", 2), ("fishing-v1", {"sigma": 0.05}, "escapement", 2),
                                   ("fishing-v2", {"sigma": 0.02}, "msy", 1), ("fishing-v1", {}, "msy", 1)]:
        with Injector(Z[500:1500]):
            env = gym.make(envid, **kw)
            pol = {"escapement": escapement, "msy": msy}[which](env)
            df = env.simulate(pol, reps=reps)
        out.append(dict(env=envid, kw=kw, policy=which, reps=reps, noise_offset=500,
                        columns=list(df.columns),
                        rows=[[int(r[0]), hx(r[1]), hx(r[2]), hx(r[3]), int(r[4])] for r in df.values.tolist()]))
    return out


# ------------------------------------------------------------------- next row f1: fishing-v5..v11 trajectories
def trajectories_ln(kind, M, T, seed):
    """Per-trajectory parameters for the log-normal growth family (growth_models.py).  v10 records the drifting r,
    v11 the model drawn at construction and at the (single) reset."""
    import copy

    rng = np.random.default_rng(seed)
    envid = "fishing-" + kind
    names = ["r", "K", "C", "M", "theta", "q", "b", "a", "alpha", "sigma", "x0"]
    P = {k: np.zeros(M) for k in names}
    P["Tmax"] = np.zeros(M, dtype=np.int32)
    P["model"] = np.zeros(M, dtype=np.int32)
    A, NZ = np.zeros((T, M)), np.zeros((T, M))
    OBS0, OBS, REW, DONE = np.zeros(M), np.zeros((T, M)), np.zeros((T, M)), np.zeros((T, M), dtype=np.uint8)
    models = ["allen", "beverton_holt", "myers", "may", "ricker"]
    for m in range(M):
        default = m % 4 == 0
        sigma = float(rng.choice([0.0, 0.02, 0.1]))
        Tmax = int(rng.choice([7, 40, 1000]))
        kw = dict(sigma=sigma, Tmax=Tmax)
        if not default:
            if kind in ("v5", "v6", "v9", "v10"):
                kw.update(r=float(rng.uniform(0.1, 1.0)), K=float(rng.uniform(0.6, 3.0)))
                kw["init_state"] = float(rng.uniform(0.1, 1.5)) * kw["K"]
            if kind == "v5":
                kw["C"] = float(rng.uniform(0.05, 0.9))
            if kind == "v7":
                kw.update(r=float(rng.uniform(0.3, 1.0)), K=float(rng.uniform(1.0, 2.0)), M=float(rng.uniform(1.0, 2.0)),
                          q=int(rng.choice([2, 3, 4])), b=float(rng.uniform(0.1, 0.3)), a=float(rng.uniform(0.1, 0.3)),
                          init_state=float(rng.uniform(0.2, 1.5)))
            if kind == "v8":
                kw.update(r=float(rng.uniform(0.5, 1.5)), K=float(rng.uniform(0.8, 1.5)), M=float(rng.uniform(0.5, 2.0)),
                          theta=float(rng.choice([2.0, 2.5, 3.0])), init_state=float(rng.uniform(0.5, 2.0)))
            if kind == "v10":
                kw["alpha"] = float(rng.uniform(-0.01, 0.005))
        if kind == "v11":
            kw = dict(Tmax=Tmax, params=copy.deepcopy({k: dict(v) for k, v in {
                "allen": {"r": 0.3, "K": 1.0, "sigma": sigma, "C": 0.5, "x0": 0.75},
                "beverton_holt": {"r": 0.3, "K": 1, "sigma": sigma, "x0": 0.75},
                "myers": {"r": 1.0, "K": 1.0, "M": 1.0, "theta": 3.0, "sigma": sigma, "x0": 1.5},
                "may": {"r": 0.7, "K": 1.5, "M": 1.5, "q": 3, "b": 0.15, "sigma": sigma, "a": 0.2, "x0": 0.75},
                "ricker": {"r": 0.3, "K": 1, "sigma": sigma, "x0": 0.75}}.items()}))
        style = m % 3
        acts = (rng.uniform(-1.0, -0.8, size=T) if style == 0 else rng.uniform(-1.0, 1.0, size=T) if style == 1
                else rng.uniform(-1.3, -0.6, size=T))
        z = rng.standard_normal(T)
        choices = [int(rng.integers(0, 5)), int(rng.integers(0, 5))]
        with Injector(np.zeros(4), choices=choices if kind == "v11" else None):
            env = gym.make(envid, **kw)
            if kind == "v10":
                env.params = dict(env.params)   # undo the shared mutable default dict (growth_models.py:136-142 quirk)
            obs = env.reset()
        if kind == "v11":
            P["model"][m] = models.index(env.model)
            pp = env.params[env.model]
        else:
            pp = env.params
        for k in names:
            if k in pp:
                P[k][m] = pp[k]
        P["sigma"][m] = pp["sigma"]
        P["Tmax"][m] = Tmax
        P["K"][m] = pp["K"]
        P["x0"][m] = env.init_state
        OBS0[m] = obs[0]
        A[:, m], NZ[:, m] = acts, z
        for t in range(T):
            with Injector(z[t:t + 1]) as inj:
                obs, rew, done, _ = env.step(np.array([acts[t]]))
                assert inj.i == 1
            OBS[t, m], REW[t, m], DONE[t, m] = obs[0], rew, done
    out = dict(actions=A, noise=NZ, obs0=OBS0, obs=OBS, reward=REW, done=DONE)
    out.update({"p_" + k: v for k, v in P.items()})
    return out


def policy_table_ln():
    """BMSY / msy of the reference for the log-normal family at sigma = 0 (what upstream's own tests use)."""
    rows = []
    for envid in ("fishing-v5", "fishing-v6", "fishing-v7", "fishing-v8", "fishing-v9"):
        with Injector(np.zeros(40000)):
            env = gym.make(envid, sigma=0)
            pol = msy(env)
        rows.append(dict(env=envid, S=hx(pol.S), msy=hx(pol.msy), S_dtype=str(np.asarray(pol.S).dtype)))
    return rows


def main():
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat(), f, indent=1)
    for i, kind in enumerate(["v0", "v1", "v2", "v4"]):
        np.savez_compressed(os.path.join(HERE, f"traj_{kind}.npz"), **trajectories(kind, 48, 160, 1000 + i))
        np.savez_compressed(os.path.join(HERE, f"autoreset_{kind}.npz"),
                            **trajectories(kind, 24, 160, 2000 + i, auto_reset=True))
    for i, kind in enumerate(["v5", "v6", "v7", "v8", "v9", "v10", "v11"]):
        np.savez_compressed(os.path.join(HERE, f"traj_{kind}.npz"), **trajectories_ln(kind, 24, 120, 3000 + i))
    with open(os.path.join(HERE, "policy.json"), "w") as f:
        json.dump(policy_table(), f, indent=1)
    with open(os.path.join(HERE, "policy_ln.json"), "w") as f:
        json.dump(policy_table_ln(), f, indent=1)
    with open(os.path.join(HERE, "simulate.json"), "w") as f:
        json.dump(simulate_tables(), f, indent=1)
    np.save(os.path.join(HERE, "Z2024.npy"), Z)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
