"""GPU parity tests: the CUDA engine, called through the C ABI, against the oracle (oracle/fishing_oracle.py) and the
golden vectors recorded from the unmodified reference (tests/golden).  Bit-exact for v0/v1/v4 (all float64
arithmetic is + - * / and comparisons); fishing-v2 within 1e-12 relative (north_star tolerance; CUDA `exp` vs numpy
`exp` may differ by an ulp); done flags and discrete actions always exact."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import fishing_oracle as fo
from oracle import philox_ref as pr

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
KIND = {"v0": 0, "v1": 1, "v2": 2, "v4": 4}
RTOL_V2 = 1e-12


def _torch():
    import torch

    return torch


def same(a, b, kind, what="", scale=1.0):
    """Bit-exact, except fishing-v2: |a - b| <= 1e-12 * max(|b|, scale) with scale = K, the size of the stock.  (A purely
    relative bound cannot hold for v2: one ulp of `exp` in a stock x becomes a large RELATIVE error in x - quota when
    the quota takes almost everything -- the same happens between two numpy builds of the reference itself.)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if kind == "v2":
        ok = np.allclose(a, b, rtol=RTOL_V2, atol=RTOL_V2 * scale, equal_nan=True)
    else:
        ok = np.array_equal(a, b, equal_nan=True)
    if not ok:
        bad = np.flatnonzero(~np.isclose(a.ravel(), b.ravel(), rtol=0 if kind != "v2" else RTOL_V2,
                                         atol=0 if kind != "v2" else RTOL_V2 * scale, equal_nan=True))
        raise AssertionError(f"{what}: {bad.size} mismatches, first at {bad[:5]}: {a.ravel()[bad[:5]]} vs {b.ravel()[bad[:5]]}")


def dev(x, dtype=None):
    torch = _torch()
    t = torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    return t if dtype is None else t.to(dtype)


def make_batch(kind, n, **kw):
    from gym_fishing_b200.engine import EnvBatch

    return EnvBatch(KIND[kind], n, **kw)


# --------------------------------------------------------------------------------------------- golden trajectories
@pytest.mark.parametrize("kind", ["v0", "v1", "v2", "v4"])
def test_golden_trajectories_one_env_per_launch(kind):
    """Each golden trajectory has its own constructor parameters, so each runs as its own 1-env batch (scalar kernel
    path) with the reference's recorded actions and noise; every step of every trajectory must match."""
    torch = _torch()
    d = np.load(os.path.join(G, f"traj_{kind}.npz"))
    T, M = d["actions"].shape
    for m in range(0, M, 2):
        b = make_batch(kind, 1, r=d["p_r"][m], K=d["p_K"][m], sigma=d["p_sigma"][m], C=d["p_C"][m], x0=d["p_x0"][m],
                       Tmax=int(d["p_Tmax"][m]), n_actions=int(d["p_n_actions"][m]), sigma_p=d["p_sigma_p"][m])
        b.reset(reset_noise=dev(d["reset_noise"][0, :, m]) if kind == "v4" else None)
        assert b.state.item() == d["obs0"][m]
        if kind == "v4":
            assert b.K_env.item() == d["K_env"][0, m] and b.r_env.item() == d["r_env"][0, m]
        acts = dev(d["actions"][:, m:m + 1], torch.int32 if kind == "v0" else torch.float64)
        noise = dev(d["noise"][:, m:m + 1])
        obs = torch.empty(T, 1, dtype=torch.float64, device="cuda")
        rew = torch.empty_like(obs)
        done = torch.empty(T, 1, dtype=torch.uint8, device="cuda")
        for t in range(T):
            b.step(acts[t], noise=noise[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t])
        same(obs.cpu().numpy()[:, 0], d["obs"][:, m], kind, f"{kind} traj {m} obs", 4.0)
        same(rew.cpu().numpy()[:, 0], d["reward"][:, m], kind, f"{kind} traj {m} reward", 4.0)
        assert np.array_equal(done.cpu().numpy()[:, 0], d["done"][:, m]), (kind, m)


@pytest.mark.parametrize("kind", ["v0", "v1", "v2", "v4"])
def test_golden_autoreset_trajectories(kind):
    torch = _torch()
    d = np.load(os.path.join(G, f"autoreset_{kind}.npz"))
    T, M = d["actions"].shape
    for m in range(0, M, 2):
        b = make_batch(kind, 1, auto_reset=True, r=d["p_r"][m], K=d["p_K"][m], sigma=d["p_sigma"][m], C=d["p_C"][m],
                       x0=d["p_x0"][m], Tmax=int(d["p_Tmax"][m]), n_actions=int(d["p_n_actions"][m]),
                       sigma_p=d["p_sigma_p"][m])
        b.reset(reset_noise=dev(d["reset_noise"][0, :, m]) if kind == "v4" else None)
        acts = dev(d["actions"][:, m:m + 1], torch.int32 if kind == "v0" else torch.float64)
        noise = dev(d["noise"][:, m:m + 1])
        rnoise = dev(d["reset_noise"][:, :, m])
        obs = torch.empty(T, 1, dtype=torch.float64, device="cuda")
        rew = torch.empty_like(obs)
        term = torch.full((T, 1), float("nan"), dtype=torch.float64, device="cuda")
        done = torch.empty(T, 1, dtype=torch.uint8, device="cuda")
        for t in range(T):
            b.step(acts[t], noise=noise[t], reset_noise=rnoise[t + 1] if kind == "v4" else None, out_obs=obs[t],
                   out_reward=rew[t], out_done=done[t], out_terminal=term[t])
        same(obs.cpu().numpy()[:, 0], d["obs"][:, m], kind, f"{kind} autoreset {m} obs", 4.0)
        same(rew.cpu().numpy()[:, 0], d["reward"][:, m], kind, f"{kind} autoreset {m} reward", 4.0)
        same(term.cpu().numpy()[:, 0], d["term"][:, m], kind, f"{kind} autoreset {m} terminal", 4.0)
        assert np.array_equal(done.cpu().numpy()[:, 0], d["done"][:, m])
        if kind == "v4":
            assert b.K_env.item() == d["K_env"][T, m] and b.r_env.item() == d["r_env"][T, m]


# --------------------------------------------------------------------------------------------- batched K1 vs oracle
CASES = [
    ("v1", dict(r=0.3, K=1.0, sigma=0.01, x0=0.75, Tmax=100)),
    ("v1", dict(r=0.7, K=2.5, sigma=0.1, x0=1.1, Tmax=30)),
    ("v0", dict(r=0.3, K=1.0, sigma=0.05, x0=0.75, Tmax=100, n_actions=100)),
    ("v0", dict(r=0.9, K=1.7, sigma=0.02, x0=0.5, Tmax=25, n_actions=7)),
    ("v2", dict(r=0.3, K=1.0, sigma=0.05, C=0.5, x0=0.75, Tmax=100)),
    ("v2", dict(r=0.6, K=3.0, sigma=0.02, C=0.9, x0=2.0, Tmax=40)),
    ("v4", dict(r=0.3, K=1.0, sigma=0.03, x0=0.75, Tmax=50, sigma_p=0.1)),
]


def random_actions(rng, kind, T, n, n_actions=100, gentle=False):
    if kind == "v0":
        hi = max(2, n_actions // 6) if gentle else n_actions + 4
        return rng.integers(-2 if not gentle else 0, hi, size=(T, n)).astype(np.int64)
    if gentle:
        return rng.uniform(-1.0, -0.8, size=(T, n))
    a = rng.uniform(-1.2, 1.2, size=(T, n))
    a[rng.random((T, n)) < 0.002] = np.nan
    return a


@pytest.mark.parametrize("kind,cfg", CASES)
@pytest.mark.parametrize("n", [4096, 1001])
@pytest.mark.parametrize("auto_reset", [False, True])
def test_batched_step_matches_oracle(kind, cfg, n, auto_reset):
    """Vector path (even n) and scalar path (odd n), injected noise, random + gentle actions, NaNs included."""
    torch = _torch()
    rng = np.random.default_rng(zlib.crc32(repr((kind, n, auto_reset, sorted(cfg.items()))).encode()))
    T = 40
    gentle = rng.random() < 0.5
    acts = random_actions(rng, kind, T, n, cfg.get("n_actions", 100), gentle)
    noise = rng.standard_normal((T, n))
    rnoise = rng.standard_normal((T + 1, 2, n))
    b = make_batch(kind, n, auto_reset=auto_reset, **cfg)
    b.reset(reset_noise=dev(rnoise[0]) if kind == "v4" else None)
    K0, r0 = cfg["K"], cfg["r"]
    if kind == "v4":
        K0, r0 = fo.v4_draw(cfg["K"], cfg["r"], cfg["sigma_p"], rnoise[0, 0], rnoise[0, 1])
        assert np.array_equal(b.K_env.cpu().numpy(), K0) and np.array_equal(b.r_env.cpu().numpy(), r0)
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
    assert np.array_equal(done.cpu().numpy(), exp["done"]), "done flags must be bit-exact"
    same(obs.cpu().numpy(), exp["obs"], kind, "obs", cfg["K"])
    same(rew.cpu().numpy(), exp["reward"], kind, "reward", cfg["K"])
    if auto_reset:
        same(term.cpu().numpy(), exp["term"], kind, "terminal obs", cfg["K"])
        assert exp["done"].sum() > 0
    assert np.array_equal(b.years.cpu().numpy(), exp["years"])


def test_action_dtypes_i32_i64_f32():
    torch = _torch()
    rng = np.random.default_rng(3)
    n, T = 2048, 10
    a = rng.integers(0, 30, size=(T, n))
    outs = []
    for dt in (torch.int32, torch.int64):
        b = make_batch("v0", n, r=0.3, K=1.3, sigma=0.0, x0=0.9, n_actions=50)
        b.reset()
        for t in range(T):
            o, r_, d_ = b.step(dev(a[t], dt))
        outs.append((o.cpu().numpy().copy(), r_.cpu().numpy().copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    # float32 Box actions: quota in float32 arithmetic (numpy >= 2 semantics), both K == 1 and K != 1
    for K in (1.0, 2.5):
        a32 = rng.uniform(-1.1, -0.5, size=n).astype(np.float32)
        b = make_batch("v1", n, r=0.3, K=K, sigma=0.0, x0=0.75 * K)
        b.reset()
        o, r_, d_ = b.step(dev(a32))
        q = fo.quota_from_action_f32(a32, K)
        s2, rew, done, _ = fo.step_vec("v1", fo.reset_state("v1", 0.75 * K, K), a32, 0.0, 0, r=0.3, K=K, sigma=0.0, quota=q)
        assert np.array_equal(o.cpu().numpy(), np.broadcast_to(s2, (n,))) and np.array_equal(r_.cpu().numpy(), rew)
    with pytest.raises(Exception, match="not valid for this environment"):
        b.step(dev(a[0], torch.int64))


def test_unaligned_slices_and_odd_offsets_take_the_scalar_path():
    torch = _torch()
    rng = np.random.default_rng(5)
    n = 1000
    big = make_batch("v1", n + 3, r=0.3, K=1.0, sigma=0.2, x0=0.75, seed=11)
    big.reset()
    acts = dev(rng.uniform(-1, -0.7, size=n + 3))
    o_big, r_big, d_big = big.step(acts)
    # same envs 3..n+3 as their own batch with env_id0 = 3 and pointers offset by 3 elements (8-byte aligned only)
    from gym_fishing_b200.engine import EnvBatch

    sub = EnvBatch(1, n, env_id0=3, r=0.3, K=1.0, sigma=0.2, x0=0.75, seed=11)
    buf = torch.empty(n + 3, dtype=torch.float64, device="cuda")
    sub.state = buf[3:]
    sub.years = torch.zeros(n + 3, dtype=torch.int32, device="cuda")[3:]
    sub.reset()
    o_sub, r_sub, d_sub = sub.step(acts[3:])
    assert torch.equal(o_sub, o_big[3:]) and torch.equal(r_sub, r_big[3:]) and torch.equal(d_sub, d_big[3:])


# --------------------------------------------------------------------------------------------- noise
def test_philox_noise_matches_numpy_restatement_and_is_partition_invariant():
    b = make_batch("v1", 4096, env_id0=2 ** 33 + 10, seed=0xDEADBEEF12345, sigma=0.1)
    z = b.noise(5, step0=2 ** 32 - 2).cpu().numpy()
    ids = np.arange(4096, dtype=np.uint64) + np.uint64(2 ** 33 + 10)
    for t in range(5):
        ref = pr.step_noise(0xDEADBEEF12345, ids, np.uint64(2 ** 32 - 2 + t))
        assert np.max(np.abs(z[t] - ref)) < 1e-13
    from gym_fishing_b200 import _native as nat

    for stream in (nat.STREAM_AUTORESET, nat.STREAM_RESET):
        zz = b.noise(2, stream=stream, step0=7).cpu().numpy()
        zK, zr = pr.reset_noise(0xDEADBEEF12345, ids, np.uint64(8), stream)
        assert np.max(np.abs(zz[1, 0] - zK)) < 1e-13 and np.max(np.abs(zz[1, 1] - zr)) < 1e-13
    sub = make_batch("v1", 100, env_id0=2 ** 33 + 10 + 501, seed=0xDEADBEEF12345, sigma=0.1)
    assert np.array_equal(sub.noise(1, step0=2 ** 32 - 2).cpu().numpy()[0], z[0, 501:601])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03


def test_philox_normals_against_long_double_evaluation():
    """The hand-written log / sqrt / sincospi of csrc/philox.cuh against an 80-bit evaluation of the same formulas."""
    n = 1 << 16
    b = make_batch("v1", n, seed=7, sigma=0.1)
    z = b.noise(4, step0=123).cpu().numpy()
    ld = np.longdouble
    pi = ld("3.14159265358979323846264338327950288")
    for t in range(4):
        pairs = np.arange(n // 2, dtype=np.uint64)
        u1, u2 = pr.uniforms(7, pairs, np.uint64(123 + t), 0)
        R = np.sqrt(ld(-2) * np.log(u1.astype(ld)))
        z0 = (R * np.cos(ld(2) * pi * u2.astype(ld))).astype(np.float64)
        z1 = (R * np.sin(ld(2) * pi * u2.astype(ld))).astype(np.float64)
        ref = np.empty(n)
        ref[0::2], ref[1::2] = z0, z1
        err = np.abs(z[t] - ref)
        assert err.max() < 3e-15, err.max()
        assert np.mean(err == 0) > 0.5   # most draws are correctly rounded


@pytest.mark.parametrize("kind", ["v1", "v2", "v4"])
def test_in_kernel_noise_equals_exported_noise(kind):
    """The stochastic-parity protocol: export the engine's own draws, inject them into the oracle."""
    torch = _torch()
    n, T = 4096, 25
    cfg = dict(r=0.3, K=1.0, sigma=0.05, C=0.5, x0=0.75, Tmax=10, sigma_p=0.1, seed=99)
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, -0.6, size=(T, n))
    b = make_batch(kind, n, auto_reset=True, **cfg)
    b.reset()
    K0 = b.K_env.cpu().numpy() if kind == "v4" else 1.0
    r0 = b.r_env.cpu().numpy() if kind == "v4" else 0.3
    from gym_fishing_b200 import _native as nat

    if kind == "v4":
        zK, zr = pr.reset_noise(99, np.arange(n), np.uint64(0), nat.STREAM_RESET)
        Kx, rx = fo.v4_draw(1.0, 0.3, 0.1, zK, zr)
        assert np.allclose(K0, Kx, rtol=1e-13) and np.allclose(r0, rx, rtol=1e-12)
    z = b.noise(T, step0=0)
    rz = b.noise(T, stream=nat.STREAM_AUTORESET, step0=0).cpu().numpy()
    rnoise = np.concatenate([np.zeros((1, 2, n)), rz])
    obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
    rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
    d_act = dev(acts)
    for t in range(T):
        b.step(d_act[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t])
    exp = fo.rollout_vec(kind, T, n, r=r0, K=K0, sigma=0.05, C=0.5, x0=0.75, Tmax=10, policy="actions", actions=acts,
                         noise=z.cpu().numpy(), auto_reset=True, reset_noise=rnoise, K_mean=1.0, r_mean=0.3, sigma_p=0.1)
    assert np.array_equal(done.cpu().numpy(), exp["done"])
    same(obs.cpu().numpy(), exp["obs"], kind, "obs")
    same(rew.cpu().numpy(), exp["reward"], kind, "reward")


# --------------------------------------------------------------------------------------------- fused rollout (K2)
@pytest.mark.parametrize("kind,cfg", CASES)
@pytest.mark.parametrize("policy", ["escapement", "msy", "actions"])
def test_fused_rollout_matches_oracle_and_single_steps(kind, cfg, policy):
    torch = _torch()
    n, T = 2048, 60
    rng = np.random.default_rng(zlib.crc32(repr((kind, policy, sorted(cfg.items()))).encode()))
    S = float(np.float32(0.4 * cfg["K"]))
    msy_q = float(np.float32(0.06 * cfg["K"]))
    acts = random_actions(rng, kind, T, n, cfg.get("n_actions", 100), gentle=True)
    noise = rng.standard_normal((T, n))
    rn0 = rng.standard_normal((2, n))
    b = make_batch(kind, n, **cfg)
    b.reset(reset_noise=dev(rn0) if kind == "v4" else None)
    K0, r0 = cfg["K"], cfg["r"]
    if kind == "v4":
        K0, r0 = fo.v4_draw(cfg["K"], cfg["r"], cfg["sigma_p"], rn0[0], rn0[1])
    out = b.rollout(T, policy=policy, S=S, msy_q=msy_q,
                    actions=dev(acts, torch.int64 if kind == "v0" else torch.float64) if policy == "actions" else None,
                    noise=dev(noise), trace=True, trace_action=True)
    exp = fo.rollout_vec(kind, T, n, r=r0, K=K0, sigma=cfg["sigma"], C=cfg.get("C", 0.5), x0=cfg["x0"],
                         n_actions=cfg.get("n_actions", 100), Tmax=cfg["Tmax"], policy=policy, actions=acts, S=S,
                         msy_q=msy_q, noise=noise)
    assert np.array_equal(out["done"].cpu().numpy(), exp["done"])
    same(out["obs"].cpu().numpy(), exp["obs"], kind, "trace obs", cfg["K"])
    same(out["reward"].cpu().numpy(), exp["reward"], kind, "trace reward", cfg["K"])
    if kind == "v0" or policy == "actions":
        assert np.array_equal(out["action"].cpu().numpy(), exp["action"].astype(np.float64), equal_nan=True)
    else:
        same(out["action"].cpu().numpy(), exp["action"], kind, "trace action", cfg["K"])
    same(out["ret"].cpu().numpy(), exp["ret"], kind, "return", cfg["K"] * T)
    assert np.array_equal(out["done_step"].cpu().numpy(), exp["done_step"])
    same(b.state.cpu().numpy(), exp["final_state"], kind, "final state", cfg["K"])
    # the specialised no-trace kernel must agree with the traced (generic) one bit for bit
    b2 = make_batch(kind, n, **cfg)
    b2.reset(reset_noise=dev(rn0) if kind == "v4" else None)
    sig = cfg["sigma"]
    for batch in (b, b2):
        batch.cfg["sigma"] = 0.0   # the fast path has no injected-noise input: compare both kernels at sigma = 0
    b.reset(reset_noise=dev(rn0) if kind == "v4" else None)
    kw = dict(policy=policy, S=S, msy_q=msy_q,
              actions=dev(acts, torch.int64 if kind == "v0" else torch.float64) if policy == "actions" else None)
    slow = b.rollout(T, trace=True, trace_action=True, **kw)   # the action trace keeps this leg on the generic kernel
    fast = b2.rollout(T, **kw)
    assert torch.equal(slow["ret"], fast["ret"]) and torch.equal(slow["done_step"], fast["done_step"])
    assert torch.equal(b.state, b2.state) and torch.equal(b.years, b2.years)
    b.cfg["sigma"] = sig


def test_fused_rollout_with_philox_equals_T_single_steps_bitwise():
    """Property that holds at any size: K2 (state in registers) == K1 launched T times, same Philox stream."""
    torch = _torch()
    n, T = 1 << 16, 50
    cfg = dict(r=0.3, K=1.0, sigma=0.05, x0=0.75, Tmax=20, seed=5)
    rng = np.random.default_rng(0)
    acts = dev(rng.uniform(-1, -0.7, size=(T, n)))
    for auto_reset in (False, True):
        b1 = make_batch("v1", n, auto_reset=auto_reset, **cfg)
        b1.reset()
        rsum = torch.zeros(n, dtype=torch.float64, device="cuda")
        for t in range(T):
            o, r_, d_ = b1.step(acts[t])
            rsum += r_ if auto_reset else 0
        b2 = make_batch("v1", n, auto_reset=auto_reset, **cfg)
        b2.reset()
        out = b2.rollout(T, policy="actions", actions=acts)
        assert torch.equal(b1.state, b2.state) and torch.equal(b1.years, b2.years)
        if auto_reset:
            assert torch.equal(out["ret"], rsum)


def test_freeze_done_and_sweep_layout():
    torch = _torch()
    # escapement sweep: env id = replicate * levels + level; noise shared across levels (common random numbers)
    levels, reps, T = 64, 32, 80
    n = levels * reps
    S_env = np.tile(np.arange(levels) / (levels - 1.0), reps)
    b = make_batch("v1", n, r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=1000, seed=21)
    b.reset()
    out = b.rollout(T, policy="escapement", S_env=dev(S_env), noise_id_div=levels, freeze_done=True)
    # oracle with the engine's own noise: replicate k uses z(noise id k, t)
    z = np.stack([pr.step_noise(21, np.arange(reps, dtype=np.uint64), np.uint64(t)) for t in range(T)])  # [T, reps]
    # device-exported values of the same stream for an exact comparison
    zb = make_batch("v1", reps, seed=21, sigma=0.1)
    z_dev = zb.noise(T, step0=0).cpu().numpy()
    assert np.max(np.abs(z_dev - z)) < 1e-13
    # per-env S: run the oracle level by level
    ret = np.zeros(n)
    dstep = np.zeros(n, dtype=np.int64)
    for j in range(levels):
        idx = np.arange(reps) * levels + j
        e = fo.rollout_vec("v1", T, reps, r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=1000, policy="escapement",
                           S=S_env[j], noise=z_dev, record=False)
        ret[idx] = e["ret"]
        dstep[idx] = e["done_step"]
    assert np.array_equal(out["ret"].cpu().numpy(), ret)
    assert np.array_equal(out["done_step"].cpu().numpy(), dstep)
    assert (dstep >= 0).any() and (dstep < 0).any()


def test_reduce_stats_is_deterministic_and_partition_invariant():
    torch = _torch()
    rng = np.random.default_rng(2)
    groups, rows = 1024, 512
    ret = rng.standard_normal(groups * rows) * 10.0 ** rng.integers(-3, 3, size=groups * rows)
    dstep = rng.integers(-1, 3, size=groups * rows).astype(np.int32)
    b = make_batch("v1", 2)
    d_ret, d_ds = dev(ret), dev(dstep)
    full = b.reduce_stats(d_ret, d_ds, groups=groups)
    again = b.reduce_stats(d_ret, d_ds, groups=groups)
    assert torch.equal(full, again)
    m = ret.reshape(rows, groups)
    ref = np.stack([m.sum(0), (m * m).sum(0), (dstep.reshape(rows, groups) >= 0).sum(0)], axis=1)
    assert np.allclose(full.cpu().numpy(), ref, rtol=1e-12, atol=1e-9)
    assert np.array_equal(full.cpu().numpy()[:, 2], ref[:, 2])
    # two half shards (aligned power-of-two row blocks) combine to exactly the same bits
    half = groups * rows // 2
    a = b.reduce_stats(d_ret[:half], d_ds[:half], groups=groups)
    c = b.reduce_stats(d_ret[half:], d_ds[half:], groups=groups)
    assert torch.equal(a + c, full)
    one = b.reduce_stats(d_ret, d_ds, groups=1).cpu().numpy()
    assert np.allclose(one[0], [ret.sum(), (ret * ret).sum(), (dstep >= 0).sum()], rtol=1e-11)


# --------------------------------------------------------------------------------------------- BMSY + scalar API KATs
def test_bmsy_kernel_against_reference_table():
    import gym_fishing_b200 as gf

    rows = json.load(open(os.path.join(G, "policy.json")))
    for row in rows:
        kind = "v" + row["env"].split("-v")[1]
        kw = row["kw"]
        if kind == "v4":
            K, r = fo.unhex(row["K_ctor"]), fo.unhex(row["r_ctor"])
            S, _ = gf.bmsy(4, r=r, K=K, K_is_f64=True)
        else:
            S, q = gf.bmsy(KIND[kind], r=kw.get("r", 0.3), K=kw.get("K", 1), C=kw.get("C", 0.5))
            assert float(q).hex() == row["msy"], row   # float32 arithmetic incl. numpy's float32 exp: exact
        assert float(S).hex() == row["S"], row


class Injector:
    def __init__(self, z):
        self.z, self.i = np.asarray(z, dtype=np.float64), 0

    def __call__(self, loc=0.0, scale=1.0, size=None):
        v = float(self.z[self.i])
        self.i += 1
        return float(loc + scale * v)

    def __enter__(self):
        self._real = np.random.normal
        np.random.normal = self
        return self

    def __exit__(self, *a):
        np.random.normal = self._real


def test_scalar_api_reproduces_reference_kats():
    """cfg 1 of BASELINE.json and Appendix B of SURVEY.md through make()/step()/policies, with the same numpy-noise
    injection protocol that generated the golden vectors from the reference."""
    import gym_fishing_b200 as gf

    KAT = json.load(open(os.path.join(G, "kat.json")))
    Z = np.load(os.path.join(G, "Z2024.npy"))
    uh = fo.unhex

    # KAT-1
    with Injector(np.zeros(2100)):
        env = gf.make("fishing-v0", Tmax=1000)
        pol = gf.escapement(env)
        assert isinstance(pol.S, np.float32) and float(pol.S).hex() == KAT["kat1"]["S"]
        obs = env.reset()
        acts, tot = [], 0.0
        for t in range(1000):
            a, _ = pol.predict(obs)
            obs, r, d, info = env.step(a)
            acts.append(int(a))
            tot += r
        k = KAT["kat1"]
        assert acts == k["actions"] and float(obs[0]).hex() == k["final_s"] and tot.hex() == k["sumR"]
        assert float(env.fish_population).hex() == k["final_x"] and d is False and env.years_passed == 1000
        a, _ = pol.predict(obs)
        assert env.step(a)[2] is True
    # KAT-2 / 2b
    with Injector(np.zeros(300)):
        env = gf.make("fishing-v1")
        env.reset()
        for t in range(101):
            obs, r, d, _ = env.step(np.array([-0.85]))
            assert (float(obs[0]).hex(), float(r).hex(), d) == (KAT["kat2"]["s"][t], KAT["kat2"]["r"][t], KAT["kat2"]["done"][t])
        env.reset()
        o1 = env.step(np.array([0.1]))
        o2 = env.step(np.array([-1.0]))
        assert [float(o1[0][0]).hex(), float(o1[1]).hex(), o1[2]] == KAT["kat2b"][0]
        assert [float(o2[0][0]).hex(), float(o2[1]).hex(), o2[2]] == KAT["kat2b"][1]
    # KAT-3 (the upstream test_tipping assertions, tests/test-envs.py:93-106)
    with Injector(np.zeros(20)):
        env = gf.make("fishing-v2", sigma=0, init_state=0.75)
        env.reset()
        obs, *_ = env.step(np.array([env.get_action(0)]))
        x = env.get_fish_population(obs)
        assert x >= 0.75 and abs(x - uh(KAT["kat3"]["x_075"])) <= 1e-12 * x
        env.init_state = 0.3
        env.reset()
        obs, *_ = env.step(np.array([env.get_action(0)]))
        x = env.get_fish_population(obs)
        assert x <= 0.3 and abs(x - uh(KAT["kat3"]["x_030"])) <= 1e-12 * x
    # KAT-5
    with Injector(Z):
        env = gf.make("fishing-v1", sigma=0.05, Tmax=1000)
        obs = env.reset()
        tot = 0.0
        for t in range(1000):
            obs, r, d, _ = env.step(np.array([-0.95]))
            tot += r
        assert float(obs[0]).hex() == KAT["kat5"]["final_s"] and tot.hex() == KAT["kat5"]["sumR"] and d is False
    # KAT-6 (v0) and KAT-7 (v2): BMSY eats Z[0]
    for name, envid in (("kat6", "fishing-v0"), ("kat7", "fishing-v2")):
        with Injector(Z):
            env = gf.make(envid, sigma=0.05, Tmax=1000)
            pol = gf.escapement(env)
            assert float(pol.S).hex() == KAT[name]["S"]
            obs = env.reset()
            tot, wsum = 0.0, 0
            for t in range(1000):
                a, _ = pol.predict(obs)
                obs, r, d, _ = env.step(a if envid == "fishing-v0" else np.array([a]))
                tot += r
                wsum += int(a) * (t + 1) if envid == "fishing-v0" else 0
        if envid == "fishing-v0":
            assert float(obs[0]).hex() == KAT[name]["final_s"] and tot.hex() == KAT[name]["sumR"] and wsum == KAT[name]["wsum"]
        else:
            assert abs(obs[0] - uh(KAT[name]["final_s"])) <= 1e-12 * abs(obs[0]) and abs(tot - uh(KAT[name]["sumR"])) <= 1e-12 * tot
    # KAT-8: v4 draws at construction and reset
    with Injector(Z):
        env = gf.make("fishing-v4", sigma=0.05, Tmax=1000)
        k = KAT["kat8"]
        assert (float(env.K).hex(), float(env.r).hex()) == (k["ctor_K"], k["ctor_r"])
        obs = env.reset()
        assert (float(env.K).hex(), float(env.r).hex(), float(obs[0]).hex()) == (k["reset_K"], k["reset_r"], k["reset_obs"])
        tot = 0.0
        for t in range(1000):
            obs, r, d, _ = env.step(np.array([-0.95]))
            tot += r
            if t == 0:
                assert (float(obs[0]).hex(), float(r).hex()) == (k["s0"], k["r0"])
            if d:
                break
        assert t == k["done_t"] and tot.hex() == k["sumR"]


def test_scalar_policies_and_simulate_tables():
    import gym_fishing_b200 as gf

    Z = np.load(os.path.join(G, "Z2024.npy"))
    rows = json.load(open(os.path.join(G, "policy.json")))
    for row in rows:
        kind = "v" + row["env"].split("-v")[1]
        with Injector(Z[row["noise_offset"]:]) as inj:
            env = gf.make(row["env"], **row["kw"])
            assert inj.i == row["draws_ctor"]
            pol = gf.msy(env)
            assert inj.i - row["draws_ctor"] == row["draws_msy"]
        assert float(pol.S).hex() == row["S"] and str(np.asarray(pol.S).dtype) == row["S_dtype"]
        assert float(pol.msy).hex() == row["msy"], row
        assert float(pol.predict(np.array([0.0]))[0]).hex() == row["msy_action"]
    tabs = json.load(open(os.path.join(G, "simulate.json")))
    for tab in tabs:
        kind = "v" + tab["env"].split("-v")[1]
        with Injector(Z[tab["noise_offset"]:]):
            env = gf.make(tab["env"], **tab["kw"])
            pol = {"escapement": gf.escapement, "msy": gf.msy}[tab["policy"]](env)
            df = env.simulate(pol, reps=tab["reps"])
        assert list(df.columns) == tab["columns"] and len(df) == len(tab["rows"])
        tol = 0.0 if kind != "v2" else 1e-12
        for mine, ref in zip(df.values.tolist(), tab["rows"]):
            assert int(mine[0]) == ref[0] and int(mine[4]) == ref[4]
            for j in (1, 2, 3):
                rv = fo.unhex(ref[j])
                assert abs(float(mine[j]) - rv) <= tol * abs(rv), (tab["env"], tab["policy"], mine, ref)


def test_edge_cases_through_scalar_api():
    import gym_fishing_b200 as gf

    KAT = json.load(open(os.path.join(G, "kat.json")))
    Z = np.load(os.path.join(G, "Z2024.npy"))
    for case in KAT["edge"]:
        kind = "v" + case["env"].split("-v")[1]
        with Injector(Z[case["noise_offset"]:]):
            env = gf.make(case["env"], **case["kw"])
            obs0 = env.reset()
            assert float(obs0[0]).hex() == case["obs0"]
            for a, (es, er, ed) in zip(case["actions"], case["steps"]):
                act = int(a) if kind == "v0" else np.array([fo.unhex(a)])
                obs, r, d, _ = env.step(act)
                same(obs[0], fo.unhex(es), kind, str(case))
                same(r, fo.unhex(er), kind, str(case))
                assert d is ed


@pytest.mark.parametrize("kind,cfg", [c for c in CASES if c[0] != "v0"])
@pytest.mark.parametrize("mode", ["plain", "auto_reset", "freeze"])
def test_fast_rollout_kernel_equals_generic_kernel_with_philox(kind, cfg, mode, monkeypatch):
    """K2's specialised kernel (in-kernel Philox, no traces) against the generic one, which is checked against the
    oracle above.  FISHING_GENERIC_ROLLOUT is read once per process, so the generic result comes from the traced call."""
    torch = _torch()
    n, T = 4096, 120
    rng = np.random.default_rng(11)
    acts = dev(rng.uniform(-1.0, -0.7, size=(T, n)))
    S = float(np.float32(0.45 * cfg["K"]))
    for policy in ("escapement", "msy", "actions"):
        res = []
        for traced in (True, False):
            b = make_batch(kind, n, auto_reset=(mode == "auto_reset"), seed=3, **cfg)
            b.reset()
            out = b.rollout(T, policy=policy, S=S, msy_q=float(np.float32(0.05 * cfg["K"])),
                            actions=acts if policy == "actions" else None, freeze_done=(mode == "freeze"), trace=traced,
                            trace_action=traced)   # the action trace keeps the traced leg on the generic kernel
            res.append((out["ret"].clone(), out["done_step"].clone(), b.state.clone(), b.years.clone(),
                        None if b.K_env is None else b.K_env.clone()))
        for x, y in zip(*res):
            if x is not None:
                assert torch.equal(x, y), (kind, mode, policy)
        assert (res[0][1] >= 0).any()


@pytest.mark.parametrize("kind,cfg", [c for c in CASES if c[0] != "v0"])
@pytest.mark.parametrize("mode", ["plain", "auto_reset", "freeze"])
@pytest.mark.parametrize("levels", [32, 96, 8])
def test_sweep_kernels_equal_generic_kernel(kind, cfg, mode, levels):
    """Escapement sweeps with common random numbers: the warp-cooperative kernel (levels % 32 == 0: one replicate pair x 32
    levels per warp, draws handed round with shuffles) and the per-thread amortised kernel (levels % 4 == 0) against the
    generic traced kernel on the same Philox stream, every done-handling mode, T not a multiple of 32."""
    torch = _torch()
    reps, T = 6, 75
    n = levels * reps
    S_env = dev(np.tile(np.linspace(0.0, 1.1 * cfg["K"], levels), reps))
    res = []
    for traced in (True, False):
        b = make_batch(kind, n, auto_reset=(mode == "auto_reset"), seed=17, **cfg)
        b.reset()
        out = b.rollout(T, policy="escapement", S_env=S_env, noise_id_div=levels, freeze_done=(mode == "freeze"), trace=traced)
        res.append((out["ret"].clone(), out["done_step"].clone(), b.state.clone(), b.years.clone(),
                    None if b.K_env is None else b.K_env.clone()))
    for x, y in zip(*res):
        if x is not None:
            assert torch.equal(x, y), (kind, mode, levels)


@pytest.mark.parametrize("kind,cfg", [c for c in CASES if c[0] != "v0"])
@pytest.mark.parametrize("auto_reset", [False, True])
def test_traced_fast_rollout_equals_generic_traces_and_single_steps(kind, cfg, auto_reset):
    """The specialised kernel with per-step observation / reward / done rows (the rollout-buffer fill of a whole chunk in
    one launch) against the generic kernel's traces (forced by also asking for the action trace) and against T launches of
    K1 on the same Philox stream: every row bit for bit."""
    torch = _torch()
    n, T = 4096, 130
    rng = np.random.default_rng(23)
    acts = dev(rng.uniform(-1.0, -0.5, size=(T, n)))
    S = float(np.float32(0.45 * cfg["K"]))
    for policy in ("actions", "escapement"):
        outs = []
        for generic in (True, False):
            b = make_batch(kind, n, auto_reset=auto_reset, seed=9, **cfg)
            b.reset()
            o = b.rollout(T, policy=policy, S=S, actions=acts if policy == "actions" else None, trace=True, trace_action=generic)
            outs.append((o, b.state.clone(), b.years.clone()))
        (g, gs, gy), (f, fs, fy) = outs
        for k in ("obs", "reward", "done", "ret", "done_step"):
            a_, b_ = g[k], f[k]
            if a_.dtype == torch.float64:
                a_, b_ = a_.view(torch.int64), b_.view(torch.int64)
            assert torch.equal(a_, b_), (kind, policy, auto_reset, k)
        assert torch.equal(gs.view(torch.int64), fs.view(torch.int64)) and torch.equal(gy, fy)
        assert bool(f["done"].any())
        if policy == "actions":
            b = make_batch(kind, n, auto_reset=auto_reset, seed=9, **cfg)
            b.reset()
            obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
            rew = torch.empty_like(obs)
            done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
            for t in range(T):
                b.step(acts[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t])
            assert torch.equal(obs.view(torch.int64), f["obs"].view(torch.int64))
            assert torch.equal(rew.view(torch.int64), f["reward"].view(torch.int64)) and torch.equal(done, f["done"])


def test_discrete_quota_division_is_the_ieee_quotient():
    """quota = (action / n_actions) * K (base_fishing_env.py:139-141) is computed with the reciprocal of n_actions and two
    residual corrections (div_count, csrc/dynamics.cuh) instead of a hardware division; Markstein's theorem says the result
    is the correctly rounded quotient.  Checked here against numpy's division for every n_actions in 1..2048 with every
    action in -64..3263 (6.8 million pairs), and for random large values; both the lean kernel (reward) and the full one
    (harvest output, which also shows negative quotas)."""
    torch = _torch()
    from gym_fishing_b200.engine import EnvBatch

    acts_np = np.arange(-64, 3264, dtype=np.int64)
    m = acts_np.size
    acts = dev(acts_np)
    state = torch.full((m,), 1e12, dtype=torch.float64, device="cuda")      # a stock no quota here exceeds
    obs = torch.empty_like(state)
    rew = torch.empty_like(state)
    harv = torch.empty_like(state)
    done = torch.empty(m, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(0)
    counts = list(range(1, 2049)) + [int(v) for v in rng.integers(2049, 2 ** 31 - 1, size=64)] + [2 ** 31 - 1]
    for n_actions in counts:
        b = EnvBatch(0, m, r=0.3, K=1.0, sigma=0.0, x0=0.75, Tmax=10 ** 6, n_actions=n_actions)
        want = acts_np.astype(np.float64) / np.float64(n_actions)
        b.step(acts, state_in=state, out_obs=obs, out_reward=rew, out_done=done, out_harvest=harv)
        got = harv.cpu().numpy()
        assert np.array_equal(got, want), (n_actions, np.flatnonzero(got != want)[:5])
        b.step(acts, state_in=state, out_obs=obs, out_reward=rew, out_done=done)          # lean kernel
        assert np.array_equal(rew.cpu().numpy(), np.maximum(want, 0.0)), n_actions
    big = rng.integers(-2 ** 62, 2 ** 62, size=m, dtype=np.int64)
    state.fill_(1e300)
    for n_actions in (1, 3, 7, 100, 12345, 2 ** 31 - 1):
        b = EnvBatch(0, m, r=0.3, K=1.0, sigma=0.0, x0=0.75, Tmax=10 ** 6, n_actions=n_actions)
        b.step(dev(big), state_in=state, out_obs=obs, out_reward=rew, out_done=done, out_harvest=harv)
        assert np.array_equal(harv.cpu().numpy(), big.astype(np.float64) / np.float64(n_actions)), n_actions


@pytest.mark.parametrize("K", [2.5, 0.3, -1.5, 2.0 ** -500, 2.0 ** 500, 2.0 ** -501, 2.0 ** 501, 1e-160, 7e150, float("inf")])
def test_uniform_K_division_is_the_ieee_quotient(K):
    """x / K for a uniform K != 1 goes through the host's reciprocal and two residual corrections when K and x are ordinary
    numbers (div_k, csrc/dynamics.cuh) and through the IEEE division otherwise.  One lock-step and a 3-step fused rollout
    of fishing-v1 / v2 with states whose magnitudes straddle every guard (zeros, denormals, 2^-502..2^-498, 2^498..2^502,
    infinities, NaN, ordinary values) must match the oracle (numpy's true division) bit for bit."""
    torch = _torch()
    rng = np.random.default_rng(5)
    mags = np.concatenate([2.0 ** rng.uniform(-504, -496, 2000), 2.0 ** rng.uniform(496, 504, 2000), rng.uniform(0, 3, 4000),
                           2.0 ** rng.uniform(-1074, -1000, 500), 2.0 ** rng.uniform(-40, 40, 1500),
                           [0.0, np.inf, np.nan, 5e-324, 1.7e308, 1.0, 2.0 ** -500, 2.0 ** 501]])
    x = mags * np.abs(K) if np.isfinite(K) else mags
    with np.errstate(all="ignore"):
        s0 = x / K - 1.0                               # states whose stock (s + 1) * K lands on the magnitudes above
        s0 = np.concatenate([s0, mags - 1.0, -mags - 1.0])
    n = s0.size + (s0.size & 1)
    s0 = np.resize(s0, n)
    acts = rng.uniform(-1.0, -0.2, size=(3, n))
    for kind in ("v1", "v2"):
        cfg = dict(r=0.3, K=K, sigma=0.0, C=0.5, x0=0.75, Tmax=5)
        b = make_batch(kind, n, **cfg)
        b.reset()
        b.state.copy_(dev(s0))
        obs = torch.empty(3, n, dtype=torch.float64, device="cuda")
        rew = torch.empty_like(obs)
        done = torch.empty(3, n, dtype=torch.uint8, device="cuda")
        for t in range(3):
            b.step(dev(acts[t]), out_obs=obs[t], out_reward=rew[t], out_done=done[t])
        with np.errstate(all="ignore"):
            exp = fo.rollout_vec(kind, 3, n, r=0.3, K=K, sigma=0.0, C=0.5, x0=0.75, Tmax=5, policy="actions", actions=acts, s0=s0)
        if kind == "v1":
            assert np.array_equal(obs.cpu().numpy(), exp["obs"], equal_nan=True), K
            assert np.array_equal(rew.cpu().numpy(), exp["reward"], equal_nan=True)
            assert np.array_equal(done.cpu().numpy(), exp["done"])
        else:   # one exp evaluation: after that an ulp of difference is amplified without bound on these explosive states
            same(obs[0].cpu().numpy(), exp["obs"][0], kind, "obs", max(abs(K), 1.0) if np.isfinite(K) else 1.0)
            assert np.array_equal(done[0].cpu().numpy(), exp["done"][0])
        b2 = make_batch(kind, n, **cfg)
        b2.reset()
        b2.state.copy_(dev(s0))
        out = b2.rollout(3, policy="actions", actions=dev(acts))
        assert torch.equal(b2.state.view(torch.int64), obs[2].view(torch.int64)), (kind, K)
        if kind == "v1":
            assert np.array_equal(out["ret"].cpu().numpy(), exp["ret"], equal_nan=True)


def test_special_values_bitwise_against_oracle():
    """NaN / inf / negative / signed-zero states, actions and noise through K1 (both paths) and K2: the comparisons
    removed from the kernels as provably redundant (max(x - h, 0), ...) must not change any bit on any input."""
    torch = _torch()
    from gym_fishing_b200.engine import EnvBatch

    specials = np.array([np.nan, np.inf, -np.inf, -5.0, -1.0, -1.0 + 2 ** -52, -0.0, 0.0, 2 ** -1074, 1e-310, 0.25, 1.0, 7.0,
                         1e300, -1e300])
    acts = np.array([np.nan, np.inf, -np.inf, -2.0, -1.0, -0.999, -0.0, 0.0, 0.5, 1.0, 1.5])
    zs = np.array([0.0, -0.0, 1.0, -3.0, 40.0, np.nan, np.inf])
    S, A, Z = np.meshgrid(specials, acts, zs, indexing="ij")
    s0, a0, z0 = S.ravel(), A.ravel(), Z.ravel()
    if s0.size % 2:
        s0, a0, z0 = s0[:-1], a0[:-1], z0[:-1]
    n = s0.size
    for kind, code, cfg in (("v1", 1, dict(r=0.3, K=1.0, sigma=0.2)), ("v1", 1, dict(r=0.7, K=2.5, sigma=0.2)),
                            ("v2", 2, dict(r=0.3, K=1.0, sigma=0.2, C=0.5)), ("v4", 4, dict(r=0.3, K=1.0, sigma=0.2))):
        for sub in (slice(0, n), slice(1, n)):          # even (vector path) and odd offset (scalar path)
            m = s0[sub].size
            b = EnvBatch(code, m, Tmax=3, x0=0.75, **cfg)
            b.reset()
            b.state.copy_(dev(s0[sub]))
            K = b.K_env.cpu().numpy() if kind == "v4" else cfg["K"]
            r = b.r_env.cpu().numpy() if kind == "v4" else cfg["r"]
            with np.errstate(all="ignore"):
                e_s, e_r, e_d, _ = fo.step_vec(kind, s0[sub], a0[sub], z0[sub], np.zeros(m, dtype=np.int64), r=r, K=K,
                                               sigma=0.2, C=cfg.get("C", 0.5), Tmax=3)
            o, rw, d = b.step(dev(a0[sub]), noise=dev(z0[sub]))
            o, rw, d = o.cpu().numpy(), rw.cpu().numpy(), d.cpu().numpy()
            assert np.array_equal(d.astype(bool), e_d), (kind, cfg)
            if kind == "v2":
                ok = np.isclose(o, e_s, rtol=1e-12, atol=1e-12, equal_nan=True) | (np.isinf(o) & (o == e_s))
                assert ok.all() and np.array_equal(rw, e_r, equal_nan=True)
            else:
                assert np.array_equal(o, e_s, equal_nan=True) and np.array_equal(rw, e_r, equal_nan=True), (kind, cfg)
                assert np.array_equal(np.signbit(o[~np.isnan(o)]), np.signbit(e_s[~np.isnan(e_s)]))
    # fused kernels (generic with trace, fast without) on the same special states, escapement policy, sigma = 0
    for K in (1.0, 2.5):
        outs = []
        for traced in (True, False):
            b = EnvBatch(1, n, r=0.3, K=K, sigma=0.0, Tmax=100, x0=0.75)
            b.reset()
            b.state.copy_(dev(s0))
            out = b.rollout(6, policy="escapement", S=0.4 * K, trace=traced, trace_action=traced)
            outs.append((out["ret"].cpu().numpy(), out["done_step"].cpu().numpy(), b.state.cpu().numpy()))
        with np.errstate(all="ignore"):
            e = fo.rollout_vec("v1", 6, n, r=0.3, K=K, sigma=0.0, x0=0.75, Tmax=100, policy="escapement", S=0.4 * K, s0=s0,
                               record=False)
        for ret, ds, st in outs:
            assert np.array_equal(ret, e["ret"], equal_nan=True) and np.array_equal(ds, e["done_step"])
            assert np.array_equal(st, e["final_state"], equal_nan=True)


@pytest.mark.parametrize("n", [66, 130, 4098])
def test_v4_vector_path_with_partial_last_warp(n):
    """fishing-v4 keeps the idle lanes of the last warp alive for the warp-cooperative reset draw; sizes that are even
    (vector path) but not multiples of 64, with Philox and with injected reset noise."""
    torch = _torch()
    rng = np.random.default_rng(n)
    T = 30
    cfg = dict(r=0.3, K=1.0, sigma=0.05, x0=0.75, Tmax=6, sigma_p=0.2, seed=4)
    acts = rng.uniform(-1.0, -0.3, size=(T, n))
    from gym_fishing_b200 import _native as nat

    for inject in (False, True):
        b = make_batch("v4", n, auto_reset=True, **cfg)
        rnoise = rng.standard_normal((T + 1, 2, n))
        b.reset(reset_noise=dev(rnoise[0]) if inject else None)
        K0, r0 = b.K_env.cpu().numpy().copy(), b.r_env.cpu().numpy().copy()
        z = b.noise(T, step0=0).cpu().numpy()
        if not inject:
            rnoise = np.concatenate([np.zeros((1, 2, n)), b.noise(T, stream=nat.STREAM_AUTORESET, step0=0).cpu().numpy()])
        obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
        rew = torch.empty_like(obs)
        done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
        d_act, d_rn = dev(acts), dev(rnoise)
        for t in range(T):
            b.step(d_act[t], reset_noise=d_rn[t + 1] if inject else None, out_obs=obs[t], out_reward=rew[t], out_done=done[t])
        exp = fo.rollout_vec("v4", T, n, r=r0, K=K0, sigma=0.05, x0=0.75, Tmax=6, policy="actions", actions=acts, noise=z,
                             auto_reset=True, reset_noise=rnoise, K_mean=1.0, r_mean=0.3, sigma_p=0.2)
        assert np.array_equal(done.cpu().numpy(), exp["done"]) and exp["done"].sum() > n
        assert np.array_equal(obs.cpu().numpy(), exp["obs"]) and np.array_equal(rew.cpu().numpy(), exp["reward"])
        assert np.array_equal(b.K_env.cpu().numpy(), exp["K"]) and np.array_equal(b.r_env.cpu().numpy(), exp["r"])


@pytest.mark.parametrize("adt", ["f64", "f32"])
def test_v4_whole_batch_equals_its_parts_with_noise_and_redraws(adt):
    """fishing-v4, 393 282 environments (a partial last warp) over several episodes each, process noise on, ~a sixth of the
    quotas taking everything: the whole batch, stepped per launch and through step_many (programmatic dependent launch
    between the steps, phase-A L2 prefetch on), against the same environment ids stepped as three smaller batches.
    Observations, rewards, dones, years and the redrawn (K, r) agree bit for bit -- including the episode of an
    environment whose redrawn K was clipped to 0 (x / 0: NaN observations until its next reset; the oracle does the same).
    (Written while a persistent variant of the kernel was measured, profiles/ab_r02.md r02o / r02p.)"""
    torch = _torch()
    n = (1 << 18) + (1 << 17) + 66
    T = 14
    cfg = dict(r=0.3, K=1.0, sigma=0.05, x0=0.75, Tmax=5, sigma_p=0.2, seed=11)
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * 1.2 - 1.0     # ~ a sixth of the quotas take everything
    if adt == "f32":
        acts = acts.float()

    def run(n_part, lo, many):
        b = make_batch("v4", n_part, env_id0=lo, auto_reset=True, **cfg)
        b.reset()
        obs = torch.empty(T, n_part, dtype=torch.float64, device="cuda")
        rew = torch.empty_like(obs)
        done = torch.empty(T, n_part, dtype=torch.uint8, device="cuda")
        a = acts[:, lo:lo + n_part].contiguous()
        if many:
            b.step_many(a, obs, rew, done)
        else:
            for t in range(T):
                b.step(a[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t])
        torch.cuda.synchronize()
        return obs, rew, done, b.K_env.clone(), b.r_env.clone(), b.years.clone()

    parts = []
    for lo, n_part in ((0, 1 << 17), (1 << 17, 1 << 17), (1 << 18, n - (1 << 18))):
        assert n_part // 2 < (1 << 17)
        parts.append(run(n_part, lo, many=False))
    ref = [torch.cat([p[k] for p in parts], dim=-1) for k in range(6)]
    assert int(ref[2].sum()) > 2 * n                     # every environment went through several resets
    for many in (False, True):
        got = run(n, 0, many)
        for k, what in enumerate(("obs", "reward", "done", "K", "r", "years")):
            # bit patterns: an environment whose redrawn K was clipped to 0 carries NaN observations for an episode (x / 0)
            a_, b_ = (got[k].view(torch.int64), ref[k].view(torch.int64)) if got[k].dtype == torch.float64 else (got[k], ref[k])
            assert torch.equal(a_, b_), (what, many)
