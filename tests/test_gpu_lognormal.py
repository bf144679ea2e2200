""""Next row" f1: fishing-v5..v11, the log-normal growth family of
gym_fishing/envs/growth_models.py, through the same C-ABI step / rollout entry points.  log / exp are the engine's table-assisted ones
(csrc/normal.cuh, powers as exp(q log x)) and numpy's in the reference (a few ulp apart), so parity is |a-b| <= 1e-12*max(|b|, K) like fishing-v2;
done flags exact."""
import os

import numpy as np
import pytest

from oracle import fishing_oracle as fo
from oracle import philox_ref as pr

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
KIND = {f"v{i}": i for i in (5, 6, 7, 8, 9, 10, 11)}
TOL = 1e-12


def close(a, b, scale):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=TOL, atol=TOL * scale, equal_nan=True)


def dev(x, dtype=None):
    import torch

    t = torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    return t if dtype is None else t.to(dtype)


def batch_for(kind, n, d, m, **kw):
    from gym_fishing_b200.engine import EnvBatch
    from gym_fishing_b200.vec_env import V11_MODELS

    g = {k: float(d["p_" + k][m]) for k in ("M", "theta", "q", "b", "a", "alpha")}
    if kind == "v11":
        sig = float(d["p_sigma"][m])
        models = [(name, dict(fo.V11_PARAMS[name], sigma=sig)) for name in V11_MODELS]
        return EnvBatch(11, n, r=0.3, K=1.0, sigma=sig, x0=0.75, Tmax=int(d["p_Tmax"][m]), models=models, **kw)
    return EnvBatch(KIND[kind], n, r=float(d["p_r"][m]), K=float(d["p_K"][m]), sigma=float(d["p_sigma"][m]),
                    C=float(d["p_C"][m]), x0=float(d["p_x0"][m]), Tmax=int(d["p_Tmax"][m]), growth_K=float(d["p_K"][m]), **g, **kw)


@pytest.mark.parametrize("kind", ["v5", "v6", "v7", "v8", "v9", "v10", "v11"])
def test_golden_trajectories(kind):
    import torch

    d = np.load(os.path.join(G, f"traj_{kind}.npz"))
    T, M = d["actions"].shape
    for m in range(M):
        b = batch_for(kind, 1, d, m)
        b.reset(reset_model=dev(d["p_model"][m:m + 1].astype(np.int32)) if kind == "v11" else None)
        assert b.state.item() == d["obs0"][m]
        acts, noise = dev(d["actions"][:, m:m + 1]), dev(d["noise"][:, m:m + 1])
        obs = torch.empty(T, 1, dtype=torch.float64, device="cuda")
        rew = torch.empty_like(obs)
        done = torch.empty(T, 1, dtype=torch.uint8, device="cuda")
        for t in range(T):
            b.step(acts[t], noise=noise[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t])
        K = 1.0 if kind == "v11" else float(d["p_K"][m])
        assert np.array_equal(done.cpu().numpy()[:, 0], d["done"][:, m]), (kind, m)
        assert close(obs.cpu().numpy()[:, 0], d["obs"][:, m], 1.0), (kind, m)
        assert close(rew.cpu().numpy()[:, 0], d["reward"][:, m], K), (kind, m)
        if kind == "v10":   # the drifted r is tracked on the host exactly like params["r"] in the reference
            r = float(d["p_r"][m])
            for _ in range(T):
                r = r + float(d["p_alpha"][m])
            assert b.cfg["r"] == r


@pytest.mark.parametrize("kind", ["v5", "v6", "v7", "v8", "v9", "v10", "v11"])
def test_batched_step_rollout_and_autoreset_against_oracle(kind):
    """Vector path with the engine's own Philox noise exported and injected into the oracle; K1 stepping == fused K2."""
    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.engine import EnvBatch
    from gym_fishing_b200.vec_env import V11_MODELS, growth_kwargs, normalise_kwargs

    n, T = 2048, 40
    kid = KIND[kind]
    kw = dict(sigma=0.1, Tmax=15) if kind != "v11" else dict(Tmax=15)
    cfg = normalise_kwargs(kid, kw)
    gk = growth_kwargs(kid, kw)
    if kind == "v11":
        gk["models"] = [(name, dict(p, sigma=0.1)) for name, p in gk["models"]]
    mk = lambda auto: EnvBatch(kid, n, seed=13, auto_reset=auto, r=cfg["r"], K=cfg["K"], sigma=cfg["sigma"], C=cfg["C"],  # noqa: E731
                               x0=cfg["x0"], Tmax=cfg["Tmax"], **gk)
    rng = np.random.default_rng(2)
    acts = rng.uniform(-1.0, -0.5, size=(T, n))
    b = mk(True)
    b.reset()
    z = b.noise(T, step0=0).cpu().numpy()
    obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
    rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
    b.step_many(dev(acts), obs, rew, done)
    # oracle replay
    ids = np.arange(n, dtype=np.uint64)
    if kind == "v11":
        w0 = pr.philox4x32_10((ids & 0xFFFFFFFF).astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32), np.uint32(0),
                              np.uint32(nat.STREAM_RESET << 24), np.uint32(13), np.uint32(0))[0]
        model = ((w0.astype(np.uint64) * np.uint64(5)) >> np.uint64(32)).astype(np.int64)
        assert np.array_equal(b.model.cpu().numpy()[:0], model[:0])
    s = np.full(n, cfg["x0"] / cfg["K"] - 1)
    years = np.zeros(n, dtype=np.int64)
    p = dict(r=cfg["r"], K=cfg["K"], C=cfg["C"], sigma=cfg["sigma"], **{k: gk.get(k, 0.0) for k in ("M", "theta", "q", "b", "a")})
    r_drift = cfg["r"]
    e_obs, e_rew, e_done = np.zeros((T, n)), np.zeros((T, n)), np.zeros((T, n), dtype=np.uint8)
    for t in range(T):
        if kind == "v10":
            r_drift = r_drift + gk["alpha"]
            p["r"] = r_drift
        if kind == "v11":
            s2, rw, dn = np.zeros(n), np.zeros(n), np.zeros(n, dtype=bool)
            for k, name in enumerate(V11_MODELS):
                sel = model == k
                a_, b_, c_, _ = fo.step_ln(name, s[sel], acts[t, sel], z[t, sel], years[sel], dict(fo.V11_PARAMS[name], sigma=0.1),
                                           K_state=1.0, Tmax=15)
                s2[sel], rw[sel], dn[sel] = a_, b_, c_
            years = years + 1
        else:
            s2, rw, dn, years = fo.step_ln(fo.MODEL_OF_KIND[kind], s, acts[t], z[t], years, p, K_state=cfg["K"], Tmax=15)
        e_rew[t], e_done[t] = rw, dn
        if kind == "v11" and dn.any():   # auto-reset redraws the model from the AUTORESET stream at step t
            w0 = pr.philox4x32_10((ids & 0xFFFFFFFF).astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32),
                                  np.uint32(t), np.uint32(nat.STREAM_AUTORESET << 24), np.uint32(13), np.uint32(0))[0]
            new = ((w0.astype(np.uint64) * np.uint64(5)) >> np.uint64(32)).astype(np.int64)
            model = np.where(dn, new, model)
        s = np.where(dn, cfg["x0"] / cfg["K"] - 1, s2)
        years = np.where(dn, 0, years)
        e_obs[t] = s
    assert np.array_equal(done.cpu().numpy(), e_done) and e_done.sum() > n
    assert close(obs.cpu().numpy(), e_obs, 1.0) and close(rew.cpu().numpy(), e_rew, float(cfg["K"]))
    # fused rollout == T single steps, bit for bit (same device functions, same Philox stream)
    b1, b2 = mk(True), mk(True)
    b1.reset(), b2.reset()
    o1 = torch.empty(T, n, dtype=torch.float64, device="cuda")
    b1.step_many(dev(acts), o1, torch.empty_like(o1), torch.empty(T, n, dtype=torch.uint8, device="cuda"))
    out = b2.rollout(T, policy="actions", actions=dev(acts), trace=True)
    assert torch.equal(out["obs"], o1) and torch.equal(b1.state, b2.state) and torch.equal(b1.years, b2.years)
    if kind == "v11":
        assert torch.equal(b1.model, b2.model)
    if kind == "v10":
        assert b1.cfg["r"] == b2.cfg["r"] == r_drift


def test_scalar_api_lognormal_family_reproduces_reference_trajectories():
    """make("fishing-v5".."v11") with the reference's constructor kwargs and the numpy-noise injection protocol."""
    import gym_fishing_b200 as gf
    from gym_fishing_b200.vec_env import V11_MODELS

    class Inj:
        def __init__(self, z, choices=None):
            self.z, self.i, self.choices, self.ci = np.asarray(z, dtype=np.float64), 0, choices, 0

        def __call__(self, loc=0.0, scale=1.0, size=None):
            v = float(self.z[self.i])
            self.i += 1
            return float(loc + scale * v)

        def choice(self, a, size=None):
            v = a[self.choices[self.ci]]
            self.ci += 1
            return v

        def __enter__(self):
            self._n, self._c = np.random.normal, np.random.choice
            np.random.normal = self
            if self.choices is not None:
                np.random.choice = self.choice
            return self

        def __exit__(self, *a):
            np.random.normal, np.random.choice = self._n, self._c

    for kind in ("v5", "v7", "v8", "v10", "v11"):
        d = np.load(os.path.join(G, f"traj_{kind}.npz"))
        T, M = d["actions"].shape
        for m in (0, 1, 5):
            g = lambda k: float(d["p_" + k][m])  # noqa: E731
            kw = dict(sigma=g("sigma"), Tmax=int(d["p_Tmax"][m]))
            if kind in ("v5", "v10"):
                kw.update(r=g("r"), K=g("K"), init_state=g("x0"))
            if kind == "v5":
                kw["C"] = g("C")
            if kind == "v10":
                kw["alpha"] = g("alpha")
            if kind == "v7":
                kw.update(r=g("r"), K=g("K"), M=g("M"), q=g("q"), b=g("b"), a=g("a"), init_state=g("x0"))
            if kind == "v8":
                kw.update(r=g("r"), K=g("K"), M=g("M"), theta=g("theta"), init_state=g("x0"))
            if kind == "v11":
                kw = dict(Tmax=int(d["p_Tmax"][m]), params={k: dict(v, sigma=g("sigma")) for k, v in fo.V11_PARAMS.items()})
            mi = int(d["p_model"][m])
            with Inj(d["noise"][:, m], choices=[0, mi] if kind == "v11" else None):
                env = gf.make("fishing-" + kind, **kw)
                obs = env.reset()
                assert obs[0] == d["obs0"][m]
                if kind == "v11":
                    assert env.model == V11_MODELS[mi]
                for t in range(T):
                    obs, r, dn, _ = env.step(np.array([d["actions"][t, m]]))
                    assert close(obs[0], d["obs"][t, m], 1.0) and close(r, d["reward"][t, m], 4.0) and dn == bool(d["done"][t, m]), (kind, m, t)


def test_policies_on_the_lognormal_family():
    """BMSY / msy through the env's device population_draw.  The reference evaluates mu in float32 with numpy's float32
    log, this engine in float64, so S may land on a neighbouring grid point where the growth curve is flat: within two
    grid steps, msy to 1e-5 relative."""
    import json

    import gym_fishing_b200 as gf

    rows = json.load(open(os.path.join(G, "policy_ln.json")))
    for row in rows:
        env = gf.make(row["env"], sigma=0)
        pol = gf.msy(env)
        S_ref, q_ref = fo.unhex(row["S"]), fo.unhex(row["msy"])
        assert isinstance(pol.S, np.float32)
        assert abs(float(pol.S) - S_ref) <= 2 * (2.0 * env.K / 10000) + 1e-7, (row["env"], float(pol.S), S_ref)
        assert abs(float(pol.msy) - q_ref) <= 1e-5 * abs(q_ref), (row["env"], float(pol.msy), q_ref)
        df = env.simulate(gf.escapement(env))
        assert len(df) == env.Tmax and df.reward.sum() > 0


def _ulps(got, want, floor=0.0):
    at = np.maximum(np.abs(want), floor)
    return np.abs(got - want) / np.spacing(at)


def test_table_assisted_log_and_exp_on_the_device_against_numpy():
    """csrc/normal.cuh:log_tab / exp_tab as the kernels evaluate them (fishing_probe_math) against numpy's log / exp, which are
    within 1 ulp of the truth: every binade, the neighbourhood of 1, table-cell boundaries, subnormals and the special values.
    tools/normal_accuracy.cpp (run by tests/test_cpu_host.py) checks the host build of the same header against long double."""
    import torch

    from gym_fishing_b200 import _native as nat

    def probe(fn, x):
        xd = torch.as_tensor(x, device="cuda")
        out = torch.empty_like(xd)
        nat.check(nat.lib().fishing_probe_math(fn, xd.numel(), xd.data_ptr(), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "fishing_probe_math")
        return out.cpu().numpy()

    rng = np.random.default_rng(7)
    m = 1.0 + rng.random(1 << 20)
    x = np.concatenate([
        np.ldexp(m, rng.integers(-1022, 1024, m.size)),                 # every binade
        np.ldexp(m, rng.integers(-6, 6, m.size)),                       # the stock sizes of the growth models
        1.0 + (rng.random(1 << 18) - 0.5) * 2.0 ** -6,                  # around 1
        np.ldexp(1.0 + (np.arange(257) - 0.5).clip(0) / 256.0, 0), np.ldexp(1.0 + (np.arange(257) - 0.5).clip(0) / 256.0, -1),
        rng.random(1 << 16) * 2.2250738585072014e-308,                  # subnormals
        [5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 1.0, 0.5, 2.0]])
    with np.errstate(all="ignore"):
        want = np.log(x)
        got = probe(0, x)
        assert _ulps(got, want, 2.0 ** -6).max() < 3.2                  # 2.1 ulp of the truth + numpy's own error
        # around 1 it is the absolute error that counts (see csrc/normal.cuh); numpy's own error is of the same size there, so
        # the truth comes from mpmath
        import mpmath

        near = np.flatnonzero(np.abs(want) < 2.0 ** -6)[::61]
        mpmath.mp.dps = 40
        err = [abs(float(mpmath.mpf(float(got[i])) - mpmath.log(mpmath.mpf(float(x[i]))))) for i in near]
        assert len(err) > 2000 and max(err) < 2.5e-18
        assert probe(0, np.array([1.0]))[0] == 0.0 and not np.signbit(probe(0, np.array([1.0]))[0])
        special = np.array([0.0, -0.0, np.inf, -np.inf, -1.0, np.nan, -5e-324])
        assert np.array_equal(probe(0, special), np.log(special), equal_nan=True)
        # exp: the range of mu of the growth models, the whole finite range, and the special values
        y = np.concatenate([(rng.random(1 << 20) - 0.5) * 40.0, (rng.random(1 << 18) - 0.5) * 1500.0, (rng.random(1 << 16) - 0.5) * 2e-3,
                            [0.0, -0.0, 709.78, 709.79, -745.13, -745.14, -708.4, 700.0, -700.0]])
        want = np.exp(y)
        got = probe(1, y)
        normal = (want > 2.3e-308) & np.isfinite(want)
        assert np.array_equal(np.isinf(got), np.isinf(want))            # overflow at the same arguments
        assert _ulps(got[normal], want[normal]).max() < 2.1              # 1.02 ulp of the truth + numpy's own error
        tiny = want <= 2.3e-308                                          # subnormal or zero results
        assert tiny.sum() > 100 and np.abs(got[tiny] - want[tiny]).max() <= 2 * 5e-324
        special = np.array([np.inf, -np.inf, np.nan, 1e308, -1e308])
        assert np.array_equal(probe(1, special), np.exp(special), equal_nan=True)
