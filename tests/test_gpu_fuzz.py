"""Randomised-configuration parity: constructor parameters, batch shapes, action distributions and done-handling modes
drawn from a seeded generator, every draw run through K1 (lock-step kernel, vector or scalar path) and K2 (fused
rollout, generic and specialised kernels) via the C ABI and compared with the oracle on the same inputs.  Same bar as
tests/test_gpu_parity.py: done flags / years / done steps exact, fishing-v0/v1/v4 bit-exact, fishing-v2 within
1e-12 * max(|b|, K).  The fixed cases of test_gpu_parity.py pin known configurations; this file walks the parameter space
between them (general K through div_z, tiny and huge Tmax, odd batch sizes and offsets, n_actions from 2 to 300,
out-of-range / NaN actions)."""
import numpy as np
import pytest

from oracle import fishing_oracle as fo

pytestmark = pytest.mark.gpu
KIND = {"v0": 0, "v1": 1, "v2": 2, "v4": 4}
RTOL_V2 = 1e-12


def _torch():
    import torch

    return torch


def dev(x, dtype=None):
    t = _torch().as_tensor(np.ascontiguousarray(x), device="cuda")
    return t if dtype is None else t.to(dtype)


def same(a, b, kind, what, scale):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if kind == "v2":
        ok = np.allclose(a, b, rtol=RTOL_V2, atol=RTOL_V2 * scale, equal_nan=True)
    else:
        ok = np.array_equal(a, b, equal_nan=True)
    if not ok:
        bad = np.flatnonzero(~np.isclose(a.ravel(), b.ravel(), rtol=0 if kind != "v2" else RTOL_V2,
                                         atol=0 if kind != "v2" else RTOL_V2 * scale, equal_nan=True))
        raise AssertionError(f"{what}: {bad.size} mismatches, first at {bad[:5]}: {a.ravel()[bad[:5]]} vs {b.ravel()[bad[:5]]}")


def draw_config(rng):
    kind = ["v0", "v1", "v2", "v4"][rng.integers(4)]
    K = 1.0 if rng.random() < 0.35 else float(rng.uniform(0.3, 3.0))
    cfg = dict(r=float(rng.uniform(0.05, 1.2)), K=K, sigma=0.0 if rng.random() < 0.25 else float(rng.uniform(0.0, 0.2)),
               x0=float(rng.uniform(0.1, 1.2) * K), Tmax=int(rng.choice([1, 2, 5, 17, 40, 1000])))
    if kind == "v0":
        cfg["n_actions"] = int(rng.choice([2, 3, 7, 100, 300]))
    if kind == "v2":
        cfg["C"] = float(rng.uniform(0.1, 0.9) * K)
    if kind == "v4":
        cfg["sigma_p"] = float(rng.uniform(0.0, 0.4))
    return kind, cfg


def draw_actions(rng, kind, T, n, n_actions, style):
    if kind == "v0":
        if style == "gentle":
            return rng.integers(0, max(2, n_actions // 5), size=(T, n)).astype(np.int64)
        return rng.integers(-3, n_actions + 5, size=(T, n)).astype(np.int64)
    if style == "gentle":
        return rng.uniform(-1.0, -0.75, size=(T, n))
    a = rng.uniform(-1.3, 1.3, size=(T, n))
    a[rng.random((T, n)) < 0.003] = np.nan
    a[rng.random((T, n)) < 0.003] = np.inf
    a[rng.random((T, n)) < 0.003] = -0.0
    return a


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_k1_against_oracle(seed):
    torch = _torch()
    from gym_fishing_b200.engine import EnvBatch

    rng = np.random.default_rng(1000 + seed)
    kind, cfg = draw_config(rng)
    n = int(rng.choice([1, 2, 63, 64, 777, 2048, 4099]))
    env_id0 = int(rng.choice([0, 0, 1, 6, 1 << 33]))
    auto_reset = bool(rng.random() < 0.6)
    T = 30
    acts = draw_actions(rng, kind, T, n, cfg.get("n_actions", 100), "gentle" if rng.random() < 0.5 else "wild")
    noise = rng.standard_normal((T, n))
    rnoise = rng.standard_normal((T + 1, 2, n))
    b = EnvBatch(KIND[kind], n, auto_reset=auto_reset, env_id0=env_id0, **cfg)
    b.reset(reset_noise=dev(rnoise[0]) if kind == "v4" else None)
    K0, r0 = cfg["K"], cfg["r"]
    if kind == "v4":
        K0, r0 = fo.v4_draw(cfg["K"], cfg["r"], cfg["sigma_p"], rnoise[0, 0], rnoise[0, 1])
    exp = fo.rollout_vec(kind, T, n, r=r0, K=K0, sigma=cfg["sigma"], C=cfg.get("C", 0.5), x0=cfg["x0"],
                         n_actions=cfg.get("n_actions", 100), Tmax=cfg["Tmax"], policy="actions", actions=acts,
                         noise=noise, auto_reset=auto_reset, reset_noise=rnoise, K_mean=cfg["K"], r_mean=cfg["r"],
                         sigma_p=cfg.get("sigma_p", 0.1))
    d_act = dev(acts, torch.int64 if kind == "v0" else torch.float64)
    d_noise, d_rn = dev(noise), dev(rnoise)
    obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
    rew = torch.empty_like(obs)
    term = torch.full((T, n), float("nan"), dtype=torch.float64, device="cuda")
    done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
    for t in range(T):
        b.step(d_act[t], noise=d_noise[t], reset_noise=d_rn[t + 1] if kind == "v4" else None, out_obs=obs[t],
               out_reward=rew[t], out_done=done[t], out_terminal=term[t] if auto_reset else None)
    what = f"seed {seed} {kind} n={n} id0={env_id0} auto_reset={auto_reset} {cfg}"
    assert np.array_equal(done.cpu().numpy(), exp["done"]), what
    same(obs.cpu().numpy(), exp["obs"], kind, what + " obs", cfg["K"])
    same(rew.cpu().numpy(), exp["reward"], kind, what + " reward", cfg["K"])
    if auto_reset:
        same(term.cpu().numpy(), exp["term"], kind, what + " terminal", cfg["K"])
    assert np.array_equal(b.years.cpu().numpy(), exp["years"]), what
    if kind == "v4":
        assert np.array_equal(b.K_env.cpu().numpy(), exp["K"]) and np.array_equal(b.r_env.cpu().numpy(), exp["r"]), what


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_k2_against_oracle(seed):
    """Fused rollout with injected noise (generic kernel) against the oracle, then the specialised kernel (in-kernel
    Philox, all three done-handling modes) against the generic one on the engine's own stream."""
    torch = _torch()
    from gym_fishing_b200.engine import EnvBatch

    rng = np.random.default_rng(5000 + seed)
    kind, cfg = draw_config(rng)
    n = int(rng.choice([2, 64, 1000, 4096]))
    T = int(rng.choice([1, 7, 45]))
    policy = ["escapement", "msy", "actions"][rng.integers(3)]
    S = float(np.float32(rng.uniform(0.05, 0.9) * cfg["K"]))
    msy_q = float(np.float32(rng.uniform(0.0, 0.2) * cfg["K"]))
    acts = draw_actions(rng, kind, T, n, cfg.get("n_actions", 100), "gentle" if rng.random() < 0.7 else "wild")
    noise = rng.standard_normal((T, n))
    rn0 = rng.standard_normal((2, n))
    d_acts = dev(acts, torch.int64 if kind == "v0" else torch.float64) if policy == "actions" else None
    b = EnvBatch(KIND[kind], n, **cfg)
    b.reset(reset_noise=dev(rn0) if kind == "v4" else None)
    K0, r0 = cfg["K"], cfg["r"]
    if kind == "v4":
        K0, r0 = fo.v4_draw(cfg["K"], cfg["r"], cfg["sigma_p"], rn0[0], rn0[1])
    out = b.rollout(T, policy=policy, S=S, msy_q=msy_q, actions=d_acts, noise=dev(noise), trace=True)
    exp = fo.rollout_vec(kind, T, n, r=r0, K=K0, sigma=cfg["sigma"], C=cfg.get("C", 0.5), x0=cfg["x0"],
                         n_actions=cfg.get("n_actions", 100), Tmax=cfg["Tmax"], policy=policy, actions=acts, S=S,
                         msy_q=msy_q, noise=noise)
    what = f"seed {seed} {kind} n={n} T={T} {policy} {cfg}"
    assert np.array_equal(out["done"].cpu().numpy(), exp["done"]), what
    same(out["obs"].cpu().numpy(), exp["obs"], kind, what + " obs", cfg["K"])
    same(out["reward"].cpu().numpy(), exp["reward"], kind, what + " reward", cfg["K"])
    same(out["ret"].cpu().numpy(), exp["ret"], kind, what + " return", cfg["K"] * T)
    assert np.array_equal(out["done_step"].cpu().numpy(), exp["done_step"]), what
    same(b.state.cpu().numpy(), exp["final_state"], kind, what + " final state", cfg["K"])
    if kind == "v0":
        return   # Discrete action spaces always take the generic kernel
    for mode in ("plain", "auto_reset", "freeze"):
        res = []
        for traced in (True, False):
            bb = EnvBatch(KIND[kind], n, auto_reset=(mode == "auto_reset"), seed=seed, **cfg)
            bb.reset()
            o = bb.rollout(T, policy=policy, S=S, msy_q=msy_q, actions=d_acts, freeze_done=(mode == "freeze"), trace=traced,
                           trace_action=traced)   # the action trace keeps the traced leg on the generic kernel
            res.append((o["ret"].clone(), o["done_step"].clone(), bb.state.clone(), bb.years.clone(),
                        None if bb.K_env is None else bb.K_env.clone()))
        for x, y in zip(*res):
            if x is not None:   # bit patterns, so that NaN states (NaN actions are part of the draw) compare equal
                xb = x.view(torch.int64) if x.dtype == torch.float64 else x
                yb = y.view(torch.int64) if y.dtype == torch.float64 else y
                assert torch.equal(xb, yb), (what, mode)
