"""FishingVecEnv (the new batched front end) on the GPU: SB3-VecEnv protocol, both I/O modes, rollout-buffer stepping,
the reference's own VecEnv consumer, and size-independent properties at BASELINE.json's full batch size."""
import numpy as np
import pytest

from oracle import fishing_oracle as fo

pytestmark = pytest.mark.gpu


def test_protocol_shapes_dtypes_autoreset_and_terminal_observation():
    import torch

    import gym_fishing_b200 as gf

    n = 512
    venv = gf.FishingVecEnv("fishing-v1", num_envs=n, sigma=0.0, Tmax=5)
    assert venv.num_envs == n and venv.action_space.shape == (1,) and venv.observation_space.dtype == np.float32
    obs = venv.reset()
    assert obs.shape == (n, 1) and obs.dtype == torch.float64 and obs.is_cuda
    assert torch.all(obs == -0.25)
    act = torch.full((n, 1), -1.0, dtype=torch.float64, device="cuda")
    act[::2] = 1.0  # even envs harvest everything -> extinct on the first step
    obs, rew, done, infos = venv.step(act)
    assert rew.shape == (n,) and done.shape == (n,) and done.dtype == torch.bool and len(infos) == n
    assert bool(done[0]) and not bool(done[1])
    assert float(obs[0, 0]) == -0.25, "auto-reset returns the reset observation"
    assert infos[0]["terminal_observation"].shape == (1,) and float(infos[0]["terminal_observation"][0]) == -1.0
    assert infos[1] == {}
    assert float(rew[0]) == 0.75 and float(rew[1]) == 0.0
    # time limit: done on step Tmax + 1 (strict '>' of base_fishing_env.py:76), counted per env since its own reset
    act[:] = -1.0
    dones = []
    for t in range(6):
        venv.step_async(act)
        obs, rew, done, infos = venv.step_wait()
        dones.append(done.cpu().numpy().copy())
    dones = np.array(dones)
    assert dones[:4].sum() == 0 and dones[4, 1::2].all() and not dones[4, ::2].any() and dones[5, ::2].all()
    assert venv.get_attr("Tmax") == [5] * n and venv.get_attr("Tmax", indices=[0, 3]) == [5, 5]
    assert venv.env_method("get_fish_population", (np.array([-0.25]),), indices=0)[0][0] == 0.75
    assert venv.env_is_wrapped(object) == [False] * n
    venv.close()
    with pytest.raises(RuntimeError):
        venv.step(act)


def test_discrete_env_and_numpy_mode_equal_torch_mode():
    import torch

    import gym_fishing_b200 as gf

    n, T = 3000, 30
    rng = np.random.default_rng(0)
    for env_id, kw, mk in (("fishing-v0", dict(sigma=0.05, Tmax=10, n_actions=50), lambda: rng.integers(0, 12, size=n)),
                           ("fishing-v1", dict(sigma=0.05, Tmax=10, K=2.0, init_state=1.2), lambda: rng.uniform(-1, -0.5, size=(n, 1))),
                           ("fishing-v4", dict(sigma=0.02, Tmax=7), lambda: rng.uniform(-1, -0.5, size=(n, 1)).astype(np.float32))):
        a = gf.FishingVecEnv(env_id, num_envs=n, io="torch", seed=3, **kw)
        b = gf.FishingVecEnv(env_id, num_envs=n, io="numpy", seed=3, **kw)
        oa, ob = a.reset(), b.reset()
        assert isinstance(ob, np.ndarray) and ob.shape == (n, 1)
        assert np.array_equal(oa.cpu().numpy(), ob)
        ndone = 0
        for t in range(T):
            act = mk()
            oa, ra, da, ia = a.step(torch.as_tensor(act, device="cuda"))
            ob, rb, db, ib = b.step(act)
            assert np.array_equal(oa.cpu().numpy(), ob) and np.array_equal(ra.cpu().numpy(), rb)
            assert np.array_equal(da.cpu().numpy(), db) and db.dtype == np.bool_
            idx = np.flatnonzero(db)
            ndone += idx.size
            if idx.size:
                i = int(idx[0])
                assert np.array_equal(ia[i]["terminal_observation"].cpu().numpy(), ib[i]["terminal_observation"])
        assert ndone > n, "several episodes per env ended"
        a.close()
        b.close()


def test_step_into_rollout_buffer_matches_oracle_with_exported_noise():
    import torch

    import gym_fishing_b200 as gf

    n, T = 4096, 64
    venv = gf.FishingVecEnv("fishing-v1", num_envs=n, sigma=0.1, Tmax=20, seed=17)
    venv.reset()
    z = venv.batch.noise(T, step0=0).cpu().numpy()
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, -0.5, size=(T, n))
    d_act = torch.as_tensor(acts, device="cuda")
    obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
    rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
    for t in range(T):
        venv.step_into(d_act[t], obs[t], rew[t], done[t])
    exp = fo.rollout_vec("v1", T, n, r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=20, policy="actions", actions=acts, noise=z,
                         auto_reset=True)
    assert np.array_equal(obs.cpu().numpy(), exp["obs"]) and np.array_equal(rew.cpu().numpy(), exp["reward"])
    assert np.array_equal(done.cpu().numpy(), exp["done"]) and exp["done"].sum() > 0
    # the C-side loop (fishing_step_many) gives the same buffers
    venv2 = gf.FishingVecEnv("fishing-v1", num_envs=n, sigma=0.1, Tmax=20, seed=17)
    venv2.reset()
    obs2, rew2, done2 = torch.empty_like(obs), torch.empty_like(rew), torch.empty_like(done)
    venv2.batch.step_many(d_act, obs2, rew2, done2)
    assert torch.equal(obs, obs2) and torch.equal(rew, rew2) and torch.equal(done, done2)
    assert torch.equal(venv2.batch.state, obs2[-1])


def test_reference_vecenv_consumer_simulate_mdp_vec():
    """shared_env.py:57-79 run against the batched env (numpy mode) with a constant-action model."""
    import gym_fishing_b200 as gf
    from gym_fishing_b200.shared_env import simulate_mdp_vec

    class Const:
        def predict(self, obs, state=None, mask=None):
            return np.full((len(obs), 1), -0.9), state

    venv = gf.FishingVecEnv("fishing-v1", num_envs=8, io="numpy", sigma=0.0, Tmax=12, auto_reset=False)
    df = simulate_mdp_vec(venv, Const(), 16)
    assert list(df.columns) == ["time", "state", "action", "reward", "rep"] and len(df) == 2 * 13 * 8
    one = df[df.rep == 11].sort_values("time")
    exp = fo.rollout_vec("v1", 12, 1, r=0.3, K=1.0, sigma=0.0, x0=0.75, Tmax=12, policy="actions",
                         actions=np.full((12, 1), -0.9))
    assert np.array_equal(one.state.values[1:], (exp["obs"][:, 0] + 1) * 1.0)
    assert np.array_equal(one.reward.values[1:], exp["reward"][:, 0])
    venv.close()


def test_full_size_properties_2pow20():
    """BASELINE.json config 2 at full size: fishing-v1, 2^20 envs, sigma = 0.01, T = 1000 lock-steps,
    U(-1, 1) actions (a third of the envs fish gently so that not everything collapses at once).  The oracle is too slow
    to replay 10^9 env-steps, so: (i) K1 x T == fused K2 bit for bit on every output, (ii) two half shards with global
    ids == the whole batch, (iii) 4096 random env ids x ALL 1000 steps replayed through the oracle with the same actions
    and the engine's own exported noise."""
    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.engine import EnvBatch

    n, T = 1 << 20, 1000
    cfg = dict(r=0.3, K=1.0, sigma=0.01, x0=0.75, Tmax=100, seed=0)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    acts[:, ::3] = acts[:, ::3] * 0.1 - 0.9   # a third of the envs fish gently and survive
    whole = EnvBatch(nat.V1, n, auto_reset=True, **cfg)
    whole.reset()
    obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
    rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
    whole.step_many(acts, obs, rew, done)
    fused = EnvBatch(nat.V1, n, auto_reset=True, **cfg)
    fused.reset()
    out = fused.rollout(T, policy="actions", actions=acts, trace=True)
    assert torch.equal(out["obs"], obs) and torch.equal(out["reward"], rew) and torch.equal(out["done"], done)
    del out
    half = n // 2
    Th = 64                                    # the shard check does not need all 1000 steps
    for lo in (0, half):
        part = EnvBatch(nat.V1, half, env_id0=lo, auto_reset=True, **cfg)
        part.reset()
        o2 = torch.empty(Th, half, dtype=torch.float64, device="cuda")
        r2 = torch.empty_like(o2)
        d2 = torch.empty(Th, half, dtype=torch.uint8, device="cuda")
        part.step_many(acts[:Th, lo:lo + half].contiguous(), o2, r2, d2)
        assert torch.equal(o2, obs[:Th, lo:lo + half]) and torch.equal(d2, done[:Th, lo:lo + half])
    idx = np.sort(np.random.default_rng(0).choice(n, 4096, replace=False))
    tidx = torch.as_tensor(idx, device="cuda")
    noise_src = EnvBatch(nat.V1, n, **cfg)
    z = np.stack([noise_src.noise(50, step0=t0)[:, tidx].cpu().numpy() for t0 in range(0, T, 50)]).reshape(T, idx.size)
    exp = fo.rollout_vec("v1", T, idx.size, r=0.3, K=1.0, sigma=0.01, x0=0.75, Tmax=100, policy="actions",
                         actions=acts[:, tidx].cpu().numpy(), noise=z, auto_reset=True)
    assert np.array_equal(obs[:, tidx].cpu().numpy(), exp["obs"]) and np.array_equal(rew[:, tidx].cpu().numpy(), exp["reward"])
    assert np.array_equal(done[:, tidx].cpu().numpy(), exp["done"])
    assert 0.2 < exp["done"].mean() < 0.8


def test_step_many_host_pipeline_equals_per_step_calls():
    import torch

    import gym_fishing_b200 as gf

    n, T = 5000, 12
    rng = np.random.default_rng(4)
    acts = rng.uniform(-1, -0.4, size=(T, n))
    a = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", seed=9, sigma=0.1, Tmax=5)
    b = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", seed=9, sigma=0.1, Tmax=5)
    c = gf.FishingVecEnv("fishing-v1", num_envs=n, io="torch", seed=9, sigma=0.1, Tmax=5)
    a.reset(), b.reset(), c.reset()
    obs, rew, done = a.step_many(acts)
    obs_t, rew_t, done_t = c.step_many(torch.as_tensor(acts, device="cuda"))
    for t in range(T):
        o, r, d, _ = b.step(acts[t])
        assert np.array_equal(o[:, 0], obs[t]) and np.array_equal(r, rew[t]) and np.array_equal(d, done[t])
    assert np.array_equal(obs_t.cpu().numpy(), obs) and np.array_equal(done_t.cpu().numpy(), done)
    assert done.sum() > n
    for v in (a, b, c):
        v.close()


def test_fused_simulate_matches_the_reference_simulate_tables():
    """simulate_mdp_fused (one K2 launch for all repetitions) against tests/golden/simulate.json (recorded from the
    reference's env.simulate) with the reference's noise injected, and against the host-loop simulate_mdp."""
    import json
    import os

    import gym_fishing_b200 as gf
    from gym_fishing_b200.shared_env import simulate_mdp_fused

    G = os.path.join(os.path.dirname(__file__), "golden")
    Z = np.load(os.path.join(G, "Z2024.npy"))
    tabs = json.load(open(os.path.join(G, "simulate.json")))
    checked = 0
    for tab in tabs:
        if tab["policy"] != "escapement":
            continue     # the msy tables use float32 actions (numpy promotion of the float32 msy constant); host loop covers them
        kind = "v" + tab["env"].split("-v")[1]
        z = Z[tab["noise_offset"] + 1:]            # BMSY consumed one draw
        ref = tab["rows"]
        reps = tab["reps"]
        # the reference consumes one draw per executed step, repetition after repetition
        lens = [sum(1 for r in ref if r[4] == rep) for rep in range(reps)]
        Tmax = 100
        noise = np.zeros((Tmax, reps))
        pos = 0
        for rep, L in enumerate(lens):
            noise[:L, rep] = z[pos:pos + L]
            pos += L
        df = simulate_mdp_fused(tab["env"], "escapement", reps, noise=noise, **tab["kw"])
        assert len(df) == len(ref)
        for mine, r in zip(df.values.tolist(), ref):
            assert int(mine[0]) == r[0] and int(mine[4]) == r[4]
            for j in (1, 2, 3):
                rv = fo.unhex(r[j])
                assert abs(mine[j] - rv) <= (1e-12 if kind == "v2" else 0.0) * max(1.0, abs(rv)), (tab["env"], mine, r)
        checked += 1
    assert checked >= 2
    # Philox noise, many repetitions: shapes and the lag convention
    df = simulate_mdp_fused("fishing-v1", "msy", 1001, sigma=0.05, Tmax=30)
    assert df.rep.max() == 1000 and (df[df.time == 0].action == 0).all() and (df[df.time == 0].reward == 0).all()
    assert df.groupby("rep").time.max().max() == 29


def test_step_many_zero_copy_path_with_pinned_host_arrays(monkeypatch):
    """With FISHING_VEC_ZEROCOPY=1 and page-locked host arrays the kernels read / write host memory directly.  The
    switch is read once per process, so this test runs in a subprocess-free way only when the variable is already set;
    otherwise it checks the default staged path on the same pinned arrays."""
    import gym_fishing_b200 as gf
    from gym_fishing_b200.vec_env import _PinnedArray

    n, T = 6000, 9
    rng = np.random.default_rng(8)
    pa, po, pr_, pd = (_PinnedArray(T * n, np.float64), _PinnedArray(T * n, np.float64), _PinnedArray(T * n, np.float64),
                       _PinnedArray(T * n, np.uint8))
    acts = pa.array.reshape(T, n)
    acts[:] = rng.uniform(-1, -0.4, size=(T, n))
    outs = (po.array.reshape(T, n), pr_.array.reshape(T, n), pd.array.reshape(T, n))
    for env_id, kw in (("fishing-v1", dict(sigma=0.1, Tmax=5)), ("fishing-v4", dict(sigma=0.05, Tmax=4)),
                       ("fishing-v10", dict(sigma=0.05, Tmax=6))):
        a = gf.FishingVecEnv(env_id, num_envs=n, io="numpy", seed=9, **kw)
        b = gf.FishingVecEnv(env_id, num_envs=n, io="numpy", seed=9, **kw)
        a.reset(), b.reset()
        obs, rew, done = a.step_many(acts, out=outs)                       # pinned in and out: zero-copy
        obs2, rew2, done2 = b.step_many(np.array(acts))                    # pageable: staged copies
        assert np.array_equal(obs, obs2) and np.array_equal(rew, rew2) and np.array_equal(done, done2)
        assert done.sum() > n
        o1, r1, d1, _ = a.step(acts[0])                                    # the handle keeps stepping consistently
        o2, r2, d2, _ = b.step(acts[0])
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2)
        a.close(), b.close()
    for p in (pa, po, pr_, pd):
        p.free()


def test_monitor_episode_statistics():
    import torch

    import gym_fishing_b200 as gf

    n, T = 256, 40
    venv = gf.FishingVecEnv("fishing-v1", num_envs=n, sigma=0.0, Tmax=7, monitor=True)
    venv.reset()
    rng = np.random.default_rng(0)
    ret, length = np.zeros(n), np.zeros(n, dtype=int)
    seen = 0
    for t in range(T):
        a = rng.uniform(-1, -0.7, size=(n, 1))
        obs, rew, done, infos = venv.step(torch.as_tensor(a, device="cuda"))
        ret += rew.cpu().numpy()
        length += 1
        for i in np.flatnonzero(done.cpu().numpy())[:5]:
            ep = infos[int(i)]["episode"]
            assert ep["l"] == length[i] and abs(ep["r"] - ret[i]) < 1e-12
            seen += 1
        d = done.cpu().numpy()
        ret[d], length[d] = 0.0, 0
    assert seen > 10


def test_render_file_logging_and_abi_error_codes(tmp_path):
    """f4: render() writes the reference's CSV row [years_passed, state, action, reward] (shared_env.py:8-12; upstream
    raises AttributeError there because `self.action` is never assigned).  Plus the argument errors of the C ABI."""
    import ctypes as ct

    import torch

    import gym_fishing_b200 as gf
    from gym_fishing_b200 import _native as nat

    path = tmp_path / "log.csv"
    env = gf.make("fishing-v1", sigma=0.0, file=str(path))
    env.reset()
    for a in (-0.9, -0.8):
        env.step(np.array([a]))
        row = env.render()
        assert row[0] == env.years_passed and row[1] == env.state[0] and row[3] == env.reward
    env.close()
    lines = path.read_text().strip().splitlines()
    assert len(lines) == 2 and lines[0].startswith("1,")

    lib = nat.lib()
    p = nat.make_params(nat.V1)
    n = 64
    st = torch.zeros(n, dtype=torch.float64, device="cuda")
    yr = torch.zeros(n, dtype=torch.int32, device="cuda")
    act = torch.zeros(n, dtype=torch.float64, device="cuda")
    rew = torch.zeros(n, dtype=torch.float64, device="cuda")
    dn = torch.zeros(n, dtype=torch.uint8, device="cuda")
    soa = nat.FishingSoA(st.data_ptr(), yr.data_ptr(), None, None, None)
    args = lambda **kw: (ct.byref(kw.get("p", p)), kw.get("n", n), 0, 0, kw.get("act", act.data_ptr()), kw.get("dt", nat.ACT_F64),  # noqa: E731
                         None, None, kw.get("soa", soa), st.data_ptr(), rew.data_ptr(), dn.data_ptr(), None, None, None, None)
    assert lib.fishing_step(*args()) == 0
    assert lib.fishing_step(*args(n=-1)) == -1                      # FISHING_E_ARG
    assert lib.fishing_step(*args(act=None)) == -1
    assert lib.fishing_step(*args(dt=nat.ACT_I64)) == -2            # FISHING_E_KIND: integer action on a Box env
    assert lib.fishing_step(*args(dt=7)) == -1
    assert lib.fishing_step(*args(act=act.data_ptr() + 4)) == -3    # FISHING_E_ALIGN
    bad = nat.make_params(3)
    assert lib.fishing_step(*args(p=bad)) == -1                     # there is no fishing-v3
    v4 = nat.make_params(nat.V4)
    assert lib.fishing_step(*args(p=v4)) == -1                      # v4 needs K_env / r_env
    assert lib.fishing_step(*args(n=0)) == 0
    assert b"not valid" in lib.fishing_error_string(-2) and lib.fishing_error_string(700) is not None
    torch.cuda.synchronize()
    h = ct.c_void_p()
    assert lib.fishing_vec_create(ct.byref(p), 16, 0, 99, ct.byref(h)) == -4       # FISHING_E_NODEVICE
    assert lib.fishing_vec_create(ct.byref(p), 16, 0, 0, ct.byref(h)) == 0
    buf = np.zeros(16)
    d8 = np.zeros(16, dtype=np.uint8)
    assert lib.fishing_vec_step(h, buf.ctypes.data, nat.ACT_F64, None, buf.ctypes.data, buf.ctypes.data, d8.ctypes.data, None) == -5
    assert lib.fishing_vec_reset(h, buf.ctypes.data) == 0 and buf[0] == -0.25
    assert lib.fishing_vec_destroy(h) == 0


def test_integration_md_snippets_run(monkeypatch):
    """The ctypes binding shown in INTEGRATION.md (sections 2 and 3) is executable as written."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.chdir(root)
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    ns = {}
    exec(blocks[0], ns)                                   # section 2: fishing_reset + fishing_step on torch memory
    import torch

    torch.cuda.synchronize()
    assert ns["rc"] == 0 and ns["done"].sum().item() > 0 and torch.isfinite(ns["state"]).all()
    exec(blocks[1], ns)                                   # section 3: host-buffer VecEnv
    v = ns["B200VecEnv"](64, sigma=0.0)
    obs = v.reset()
    assert obs.shape == (64, 1) and obs[0, 0] == -0.25
    obs, rew, done, infos = v.step(np.full((64, 1), -0.9))
    assert abs(rew[0] - 0.1) < 1e-12 and not done.any() and len(infos) == 64
    v.close()


def test_examples_run_and_ppo_improves():
    """examples/: the escapement walk-through and a short PPO run on the batched env (a few iterations must already
    beat the untrained policy)."""
    import importlib.util
    import os
    import runpy

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runpy.run_path(os.path.join(root, "examples", "const_escapement.py"), run_name="__main__")
    spec = importlib.util.spec_from_file_location("ppo_torch", os.path.join(root, "examples", "ppo_torch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rets = mod.main(["--envs", "8192", "--iters", "25", "--steps", "64"])
    # constant escapement (the optimum) earns 7.67 per episode on this model, MSY 6.32; PPO passes MSY within seconds
    assert len(rets) >= 20 and max(rets[-3:]) > 6.3


def test_small_and_odd_batches():
    import torch

    import gym_fishing_b200 as gf
    from gym_fishing_b200.engine import EnvBatch

    for n in (1, 2, 3, 7):
        v = gf.FishingVecEnv("fishing-v1", num_envs=n, sigma=0.1, Tmax=4, seed=2)
        big = gf.FishingVecEnv("fishing-v1", num_envs=8, sigma=0.1, Tmax=4, seed=2)
        o, ob = v.reset(), big.reset()
        for t in range(9):
            a = torch.full((8, 1), -0.8 + 0.01 * t, dtype=torch.float64, device="cuda")
            o, r, d, _ = v.step(a[:n])
            ob, rb, db, _ = big.step(a)
            assert torch.equal(o, ob[:n]) and torch.equal(r, rb[:n]) and torch.equal(d, db[:n])   # same global env ids
    b = EnvBatch(1, 3)
    b.reset()
    with pytest.raises(ValueError, match="pairs"):
        b.rollout(4, policy="escapement", S=0.5)
    b = EnvBatch(1, 8)
    b.reset()
    with pytest.raises(ValueError, match="multiples"):
        b.rollout(4, policy="escapement", S=0.5, noise_id_div=3)
    assert b.rollout(0, policy="escapement", S=0.5)["ret"].abs().sum().item() == 0.0     # T = 0 is a no-op


def test_sb3_dtypes_are_the_float32_casts_of_the_float64_path():
    import torch

    import gym_fishing_b200 as gf

    n, T = 5000, 20
    rng = np.random.default_rng(3)
    acts32 = rng.uniform(-1, -0.4, size=(T, n)).astype(np.float32)        # what an SB3 policy emits
    ref = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", seed=5, sigma=0.1, Tmax=5)
    a = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", seed=5, sigma=0.1, Tmax=5, sb3_dtypes=True)
    b = gf.FishingVecEnv("fishing-v1", num_envs=n, io="numpy", seed=5, sigma=0.1, Tmax=5, sb3_dtypes=True)
    c = gf.FishingVecEnv("fishing-v1", num_envs=n, io="torch", seed=5, sigma=0.1, Tmax=5, sb3_dtypes=True)
    o0 = ref.reset()
    assert a.reset().dtype == np.float32 and np.array_equal(a.reset(), o0.astype(np.float32))
    b.reset(), c.reset()
    obs64, rew64, done64 = ref.step_many(acts32)
    obs, rew, done = a.step_many(acts32)
    assert obs.dtype == np.float32 and rew.dtype == np.float32 and done.dtype == np.bool_
    assert np.array_equal(obs, obs64.astype(np.float32)) and np.array_equal(rew, rew64.astype(np.float32))
    assert np.array_equal(done, done64)
    for t in range(T):
        o, r, d, _ = b.step(acts32[t])
        ot, rt, dt, _ = c.step(torch.as_tensor(acts32[t], device="cuda"))
        assert o.dtype == np.float32 and np.array_equal(o[:, 0], obs[t]) and np.array_equal(r, rew[t]) and np.array_equal(d, done[t])
        assert ot.dtype == torch.float32 and np.array_equal(ot.cpu().numpy()[:, 0], obs[t]) and np.array_equal(rt.cpu().numpy(), rew[t])
    for v in (ref, a, b, c):
        v.close()


def test_user_action_policy_reads_a_quota_from_stdin(monkeypatch):
    """models/policies.py:34-48 of the reference: prompt with the stock, parse the quota, map it through get_action."""
    import builtins

    import gym_fishing_b200 as gf
    from gym_fishing_b200.models import user_action

    for env_id, expect in (("fishing-v1", 0.2 / 1.0 - 1.0), ("fishing-v0", 20)):
        env = gf.make(env_id)
        obs = env.reset()
        prompts = []
        monkeypatch.setattr(builtins, "input", lambda prompt: (prompts.append(prompt), "0.2")[1])
        action, state = user_action(env).predict(obs)
        assert "fish population: 0.75" in prompts[0] and state is obs
        assert action == expect
        env.step(action)


def test_dqn_example_learns_on_the_discrete_env():
    """examples/dqn_torch.py, the counterpart of the reference's tests/test-DQN.py: epsilon-greedy DQN on the batched
    fishing-v0 (int64 actions, terminal observations for the bootstrap) must leave the random policy well behind."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dqn_torch", os.path.join(root, "examples", "dqn_torch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rets = mod.main(["--envs", "4096", "--iters", "300"])
    assert len(rets) > 20 and rets[0] < 1.5 and max(rets[-5:]) > 2.5


def test_wire_formats_packed_dones_and_float32_per_array():
    """fishing_vec_step_many_wire (ABI v3): bit-packed dones and float32 observations / rewards are exactly the packed /
    cast forms of the float64 / uint8 arrays, for an environment count that leaves a partial last word in every chunk."""
    import gym_fishing_b200 as gf

    n, T = 10_010, 12                       # even (pairs) but not a multiple of 8, 64 or the 1024-env chunk granularity
    rng = np.random.default_rng(11)
    acts = rng.uniform(-1, 0.2, size=(T, n))
    kw = dict(num_envs=n, io="numpy", seed=9, sigma=0.1, Tmax=4)
    ref = gf.FishingVecEnv("fishing-v1", **kw)
    pk = gf.FishingVecEnv("fishing-v1", packed_done=True, **kw)
    sb = gf.FishingVecEnv("fishing-v1", packed_done=True, sb3_dtypes=True, **kw)
    for v in (ref, pk, sb):
        v.reset()
    obs, rew, done = ref.step_many(acts)
    o1, r1, bits = pk.step_many(acts)
    assert bits.dtype == np.uint8 and bits.shape == (T, (n + 7) // 8)
    assert np.array_equal(o1, obs) and np.array_equal(r1, rew) and np.array_equal(pk.unpack_dones(bits), done)
    o2, r2, bits2 = sb.step_many(acts)
    assert o2.dtype == np.float32 and np.array_equal(o2, obs.astype(np.float32)) and np.array_equal(r2, rew.astype(np.float32))
    assert np.array_equal(bits2, bits) and done.any() and not done.all()
    # a second call continues the same trajectories (state, step counter, staging reuse)
    obs_b, _, done_b = ref.step_many(acts[:5])
    o1b, _, bits_b = pk.step_many(acts[:5])
    assert np.array_equal(o1b, obs_b) and np.array_equal(pk.unpack_dones(bits_b), done_b)
    # a longer call than any before: the [T, n] staging blocks are re-allocated
    more = rng.uniform(-1, 0.2, size=(2 * T + 3, n))
    obs_c, rew_c, done_c = ref.step_many(more)
    o1c, r1c, bits_c = pk.step_many(more)
    assert bits_c.shape == (2 * T + 3, (n + 7) // 8) and np.array_equal(o1c, obs_c) and np.array_equal(r1c, rew_c)
    assert np.array_equal(pk.unpack_dones(bits_c), done_c)
    assert sb.step_many(more)[0].shape == (2 * T + 3, n)
    with pytest.raises(ValueError, match="even"):
        gf.FishingVecEnv("fishing-v1", num_envs=7, io="numpy", packed_done=True)
    for v in (ref, pk, sb):
        v.close()


def test_step_many_host_path_equals_single_host_steps():
    """The staged [T, chunk] pipeline of fishing_vec_step_many against T calls of fishing_vec_step, all kinds of env."""
    import gym_fishing_b200 as gf

    n, T = 6144 + 2, 9
    rng = np.random.default_rng(2)
    for env_id, kw in (("fishing-v1", dict(sigma=0.05, Tmax=3)), ("fishing-v2", dict(sigma=0.05, Tmax=3)),
                       ("fishing-v4", dict(sigma=0.05, Tmax=3)), ("fishing-v0", dict(sigma=0.05, Tmax=3)),
                       ("fishing-v10", dict(sigma=0.05, Tmax=3))):
        acts = rng.integers(0, 60, size=(T, n)) if env_id == "fishing-v0" else rng.uniform(-1, -0.2, size=(T, n))
        a = gf.FishingVecEnv(env_id, num_envs=n, io="numpy", seed=4, **kw)
        b = gf.FishingVecEnv(env_id, num_envs=n, io="numpy", seed=4, **kw)
        assert np.array_equal(a.reset(), b.reset())
        obs, rew, done = a.step_many(acts)
        for t in range(T):
            o, r, d, _ = b.step(acts[t])
            assert np.array_equal(o[:, 0], obs[t]) and np.array_equal(r, rew[t]) and np.array_equal(d, done[t]), (env_id, t)
        a.close(), b.close()


def test_attributes_of_the_numpy_mode_batch_and_render():
    """get_attr / set_attr / render without a NotImplementedError: per-environment arrays travel through
    fishing_vec_fetch_array / fishing_vec_store_array, shared parameters through fishing_vec_set_params."""
    import gym_fishing_b200 as gf

    n = 64
    for io in ("numpy", "torch"):
        v = gf.FishingVecEnv("fishing-v4", num_envs=n, io=io, seed=1, sigma=0.0, Tmax=50)
        v.reset()
        K, r = v.get_attr("K"), v.get_attr("r")
        assert len(K) == n and len(set(K)) > n // 2 and all(0 <= k <= 1e6 for k in K) and len(set(r)) > n // 2
        assert v.get_attr("K", indices=[3]) == [K[3]]
        v.set_attr("K", 2.0, indices=[0, 1])               # per-environment array: any subset
        assert v.get_attr("K", indices=[0, 1, 2]) == [2.0, 2.0, K[2]]
        v.set_attr("state", -0.5, indices=5)
        assert float(v.get_attr("state", indices=5)[0][0]) == -0.5
        assert v.get_attr("years_passed") == [0] * n
        tab = v.render()
        assert tab.shape == (n, 3) and tab[5, 1] == -0.5 and tab[5, 2] == (-0.5 + 1) * v.get_attr("K")[5]
        with pytest.raises(ValueError, match="whole batch"):
            v.set_attr("sigma", 0.1, indices=[0])
        v.set_attr("Tmax", 1)                              # shared parameter: takes effect at the next step
        a = np.full((n, 1), -1.0)
        if io == "torch":
            import torch

            a = torch.as_tensor(a, device="cuda")
        _, _, d1, _ = v.step(a)
        _, _, d2, _ = v.step(a)
        d1 = d1.cpu().numpy() if hasattr(d1, "cpu") else d1
        d2 = d2.cpu().numpy() if hasattr(d2, "cpu") else d2
        assert not d1.any() and d2.all()                   # years_passed = 2 > Tmax = 1
        v.close()
    v1 = gf.FishingVecEnv("fishing-v1", num_envs=8, io="numpy", K=3.0)
    v1.reset()
    assert v1.get_attr("K") == [3.0] * 8 and v1.get_attr("fish_population") == [0.75] * 8
    with pytest.raises(AttributeError):
        v1.get_attr("model")
    v1.close()


def test_lognormal_kinds_are_refused_by_the_logistic_bmsy_and_population_draw():
    """fishing_bmsy / fishing_population_draw only know the logistic and tipping-point models: FISHING_E_KIND for v5..v11
    (they used to fall through to a logistic answer), and simulate_mdp_fused asks for S / msy_q instead of computing them."""
    import ctypes as ct

    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.shared_env import simulate_mdp_fused

    S, q = ct.c_double(), ct.c_double()
    for kind in nat.LOGNORMAL_KINDS:
        assert nat.lib().fishing_bmsy(kind, 0.3, 1.0, 0.5, 0, ct.byref(S), ct.byref(q)) == -2
        p = nat.make_params(kind, models=[("ricker", {"r": 0.3, "K": 1.0})] if kind == nat.V11 else None)
        x = torch.ones(16, dtype=torch.float64, device="cuda")
        assert nat.lib().fishing_population_draw(ct.byref(p), 16, x.data_ptr(), 0, 0.0, x.data_ptr(), None) == -2
    assert nat.lib().fishing_bmsy(nat.V1, 0.3, 1.0, 0.5, 0, ct.byref(S), ct.byref(q)) == 0 and 0.49 < S.value < 0.51
    with pytest.raises(ValueError, match="msy_q"):
        simulate_mdp_fused("fishing-v6", policy="msy", reps=2, S=0.4)
    with pytest.raises(ValueError, match="pass S"):
        simulate_mdp_fused("fishing-v9", policy="escapement", reps=2)


def test_step_ex_done_bits_need_the_vector_path():
    import ctypes as ct

    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.engine import EnvBatch

    b = EnvBatch(nat.V1, 7)                # odd n: scalar kernel
    b.reset()
    a = torch.zeros(7, dtype=torch.float64, device="cuda")
    rew = torch.empty(7, dtype=torch.float64, device="cuda")
    done = torch.empty(7, dtype=torch.uint8, device="cuda")
    bits = torch.zeros(1, dtype=torch.int64, device="cuda")
    o32 = torch.empty(7, dtype=torch.float32, device="cuda")
    x = nat.FishingStepExtra(out_done_bits=bits.data_ptr())
    p = b.params()
    args = (ct.byref(p), 7, 0, 0, a.data_ptr(), nat.ACT_F64, None, None, b.soa(), b.state.data_ptr(), rew.data_ptr(), done.data_ptr(), None)
    assert nat.lib().fishing_step_ex(*args, ct.byref(x), None) == -3              # FISHING_E_ALIGN
    x = nat.FishingStepExtra(out_obs32=o32.data_ptr())
    assert nat.lib().fishing_step_ex(*args, ct.byref(x), None) == 0              # float32 copies work on the scalar path too
    torch.cuda.synchronize()
    assert np.array_equal(o32.cpu().numpy(), b.state.cpu().numpy().astype(np.float32))
    x = nat.FishingStepExtra(ep_return=rew.data_ptr())                            # statistics: all three arrays or none
    assert nat.lib().fishing_step_ex(*args, ct.byref(x), None) == -1


def test_env_simulate_runs_on_the_device_with_the_reference_use_of_numpys_generator():
    """env.simulate(escapement | msy) as ONE fused rollout (shared_env.simulate_on_device): the table and the state of
    numpy's global generator afterwards are those of the host loop over step() -- i.e. of the reference's simulate_mdp
    (shared_env.py:29-54) -- including repetitions that end early and fishing-v4's two draws per reset."""
    import gym_fishing_b200 as gf
    from gym_fishing_b200.shared_env import simulate_mdp, simulate_on_device

    cases = [("fishing-v1", dict(sigma=0.05, Tmax=40), "escapement", 7), ("fishing-v0", dict(sigma=0.1, Tmax=25), "msy", 6),
             ("fishing-v2", dict(sigma=0.08, Tmax=30), "escapement", 5), ("fishing-v4", dict(sigma=0.05, Tmax=20), "escapement", 6),
             # constant quotas under strong noise: some repetitions go extinct and end early (the draws they leave unused
             # move to the next repetition); with many of them the device route gives up and the host loop answers
             ("fishing-v1", dict(sigma=0.25, Tmax=60), "msy", 12), ("fishing-v1", dict(sigma=0.4, Tmax=60), "msy", 12),
             ("fishing-v1", dict(sigma=0.9, Tmax=60), "msy", 12)]
    early_on_device = 0
    for env_id, kw, pol, reps in cases:
        np.random.seed(123)
        env = gf.make(env_id, **kw)
        model = {"escapement": gf.escapement, "msy": gf.msy}[pol](env)
        st = np.random.get_state()
        host = simulate_mdp(env, model, reps)
        after_host = np.random.get_state()[1].copy(), np.random.get_state()[2:]
        n_early = reps - int((host.groupby("rep").size() == kw["Tmax"]).sum())
        np.random.set_state(st)
        dev = simulate_on_device(env, model, reps)
        assert dev is not None or n_early > 2, (env_id, kw, n_early)
        np.random.set_state(st)
        pub = env.simulate(model, reps)                               # the public entry point: device route, else the host loop
        after_pub = np.random.get_state()[1].copy(), np.random.get_state()[2:]
        assert np.array_equal(after_host[0], after_pub[0]) and after_host[1] == after_pub[1], (env_id, kw)   # same draws consumed
        assert list(pub.columns) == list(host.columns) and len(pub) == len(host), (env_id, len(pub), len(host))
        tol = 1e-12 if env_id == "fishing-v2" else 0.0
        a, b = pub.values.astype(np.float64), host.values.astype(np.float64)
        assert np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1.0)), (env_id, kw)
        if dev is not None:
            assert dev.equals(pub)
            early_on_device += int(n_early > 0)
    assert early_on_device >= 1, "no case with an early ending went through the device route"
    # Philox mode: no host generator involved, repetitions are independent environments
    env = gf.make("fishing-v1", sigma=0.1, Tmax=50, noise="philox", seed=3)
    df = env.simulate(gf.escapement(env), reps=200)
    assert len(df) == 200 * 50 and df.groupby("rep").reward.sum().std() > 0
    # a foreign model goes through the host loop
    class Const:
        def predict(self, obs, **kw):
            return np.array([-0.9]), None
    host_df = env.simulate(Const(), reps=2)
    assert 2 <= len(host_df) <= 100 and set(host_df.rep) == {0, 1} and (host_df.action.iloc[1:5] > 0).all()


@pytest.mark.parametrize("env_id,kw", [("fishing-v0", dict(sigma=0.05, Tmax=6)), ("fishing-v2", dict(sigma=0.05, Tmax=6)),
                                       ("fishing-v4", dict(sigma=0.05, Tmax=6)), ("fishing-v11", dict(Tmax=6)),
                                       ("fishing-v1", dict(sigma=0.05, Tmax=6, K=2.0))])
def test_extra_outputs_of_the_step_kernel_for_every_family(env_id, kw):
    """monitor= / sb3_dtypes= (the EXTRA instantiation of K1, fishing_step_ex) against the plain environment: same trajectories,
    float32 values are the casts, episode statistics equal the sums / counts taken from the plain run -- for the Discrete, the
    tipping-point, the per-environment-parameter, the model-uncertainty and the general-K kernels, with an environment count that
    leaves a partial last warp."""
    import torch

    import gym_fishing_b200 as gf

    n, T = 1000 + 38, 30
    g = torch.Generator(device="cuda").manual_seed(5)
    plain = gf.FishingVecEnv(env_id, num_envs=n, seed=21, **kw)
    extra = gf.FishingVecEnv(env_id, num_envs=n, seed=21, monitor=True, sb3_dtypes=True, **kw)
    o0, o1 = plain.reset(), extra.reset()
    assert o1.dtype == torch.float32 and torch.equal(o1, o0.float())
    ret = torch.zeros(n, dtype=torch.float64, device="cuda")
    length = torch.zeros(n, dtype=torch.int64, device="cuda")
    episodes = 0
    for t in range(T):
        if env_id == "fishing-v0":
            a = torch.randint(0, 40, (n, 1), device="cuda", generator=g)
        else:
            a = torch.rand(n, 1, dtype=torch.float64, device="cuda", generator=g) * 0.7 - 1.0
        o, r, d, _ = plain.step(a)
        o32, r32, d2, infos = extra.step(a)
        assert torch.equal(d, d2) and torch.equal(o32, o.float()) and torch.equal(r32, r.float())
        assert torch.equal(extra.batch.state, plain.batch.state)              # the float64 state itself is untouched
        ret += r
        length += 1
        idx, er, el = infos.episode_arrays()
        assert np.array_equal(idx, d.nonzero().reshape(-1).cpu().numpy())
        assert np.array_equal(er, ret[d].cpu().numpy()) and np.array_equal(el, length[d].cpu().numpy())
        episodes += len(idx)
        ret[d] = 0
        length[d] = 0
    assert episodes > n
    plain.close(), extra.close()
