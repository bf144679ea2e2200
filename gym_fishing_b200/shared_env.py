"""Rollout recording and reporting helpers with the reference's outputs (gym_fishing/envs/shared_env.py:8-127):
simulate_mdp / simulate_mdp_vec tables with columns [time, state, action, reward, rep], estimate_policyfn, and the
matplotlib plots (only if matplotlib is installed)."""
from __future__ import annotations

import numpy as np
from pandas import DataFrame

COLUMNS = ["time", "state", "action", "reward", "rep"]
_NUMPY_NORMAL = np.random.normal      # to notice a test harness that has replaced the global generator's draw function


def simulate_mdp(env, model, reps=1):
    """Policy/env loop of at most env.Tmax steps per repetition.  Row t holds the stock BEFORE step t together with
    the quota and reward of step t-1 (one-step lag, shared_env.py:37-38,49), and the loop stops after a done step."""
    rows = []
    for rep in range(int(reps)):
        obs = env.reset()
        last_quota, last_reward = 0.0, 0.0
        t = 0
        while t < env.Tmax:
            rows.append([t, env.get_fish_population(obs), last_quota, last_reward, rep])
            action, _ = model.predict(obs)
            obs, last_reward, done, _ = env.step(action)
            if isinstance(action, np.ndarray):
                action = action[0]
            if isinstance(last_reward, np.ndarray):
                last_reward = last_reward[0]
            last_quota = env.get_quota(action)
            t += 1
            if done:
                break
    return DataFrame(rows, columns=COLUMNS)


def simulate_mdp_vec(env, model, n_eval_episodes):
    """The reference's VecEnv consumer (shared_env.py:57-79) rebuilt on whole-batch arrays: one row per (t, env) for
    t = 0..Tmax, `rep` numbering episodes across repetitions.  Works with FishingVecEnv in either io mode and with
    any object following the SB3 VecEnv protocol."""
    if n_eval_episodes % env.num_envs != 0:
        raise AssertionError("number of evaluations needs to be divisible by the number of parallel environments")
    n = env.num_envs
    reps = n_eval_episodes // n
    Tmax = env.get_attr("Tmax")[0]
    K = np.asarray(env.get_attr("K"), dtype=np.float64)
    chunks = []

    def to_np(a):
        return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)

    for rep in range(reps):
        obs = env.reset()
        state = None
        done = np.zeros(n, dtype=bool)
        action = np.full((n, 1), env.action_space.low[0] if hasattr(env.action_space, "low") else 0)
        reward = np.zeros(n)
        for t in range(Tmax + 1):
            x = (to_np(obs).reshape(n) + 1) * K
            chunks.append(np.column_stack([np.full(n, t), x, to_np(action).reshape(n), to_np(reward).reshape(n),
                                           rep * n + np.arange(n)]))
            if t == Tmax:
                break
            action, state = model.predict(obs, state=state, mask=done)
            obs, reward, done, _ = env.step(action)
    df = DataFrame(np.concatenate(chunks), columns=COLUMNS)
    return df.astype({"time": int, "rep": int})


def estimate_policyfn(env, model, reps=1, n=50):
    """Policy as a function of the stock over n observation points (shared_env.py:82-102)."""
    grid = np.linspace(env.observation_space.low, env.observation_space.high, num=n, dtype=env.observation_space.dtype)
    rows = []
    for rep in range(int(reps)):
        for obs in grid:
            action, _ = model.predict(obs)
            if isinstance(action, np.ndarray):
                action = action[0]
            rows.append([env.get_fish_population(obs), env.get_quota(action), rep])
    return DataFrame(rows, columns=["state", "action", "rep"])


def _plt():
    try:
        import matplotlib.pyplot as plt  # noqa: WPS433
    except Exception as e:  # noqa: BLE001
        raise RuntimeError("plotting needs matplotlib, which is not installed") from e
    return plt


def plot_mdp(self, df, output="results.png"):
    plt = _plt()
    fig, axs = plt.subplots(3, 1)
    for rep, g in df.groupby("rep"):
        axs[0].plot(g.time, g.state, color="blue", alpha=0.3)
        axs[1].plot(g.time, g.action, color="blue", alpha=0.3)
        axs[2].plot(g.time, np.cumsum(g.reward), color="blue", alpha=0.3)
    for ax, lab in zip(axs, ("state", "action", "reward")):
        ax.set_ylabel(lab)
    fig.tight_layout()
    plt.savefig(output)
    plt.close("all")


def plot_policyfn(self, df, output="policy.png"):
    plt = _plt()
    for rep, g in df.groupby("rep"):
        plt.plot(g.state, g.state - g.action, color="blue")
    plt.savefig(output)
    plt.close("all")


def simulate_mdp_fused(env_id, policy="escapement", reps=1, *, S=None, msy_q=None, noise=None, reset_noise=None, seed=0,
                       device="cuda", return_steps=False, **env_kwargs):
    """simulate_mdp (shared_env.py:29-54) for one of the built-in policies with ALL repetitions run as one fused rollout on
    the device: `reps` environments, at most Tmax steps, an environment stops at its first done (the `break` at :51).
    Returns the same table [time, state, action, reward, rep] (same one-step lag of the quota / reward columns).

    policy: "escapement" (S = BMSY by default) or "msy" (msy_q = surplus production at BMSY by default);
    noise: optional [Tmax, reps] standard normals to inject (else the engine's Philox stream keyed by `seed`);
    reset_noise: optional [2, reps] (zK, zr) for the parameter draws of fishing-v4's reset;
    return_steps: also return the number of steps every repetition took."""
    import torch

    from . import _native as nat
    from .engine import EnvBatch, bmsy
    from .vec_env import growth_kwargs, normalise_kwargs

    kind = nat.KIND_OF_ID[env_id]
    cfg = normalise_kwargs(kind, env_kwargs)
    if kind in nat.LOGNORMAL_KINDS and (S is None or (policy == "msy" and msy_q is None)):
        # the device BMSY kernel only knows the logistic / tipping-point models: never let it stand in for this family
        raise ValueError("fishing-v5..v11: pass S (and msy_q for policy='msy'); models.policies.BMSY / msy on a scalar env "
                         "compute them for the log-normal growth family")
    if S is None or (policy == "msy" and msy_q is None):
        S0, q0 = bmsy(kind, r=cfg["r"], K=cfg["K"], C=cfg["C"], K_is_f64=(kind == nat.V4))
        S = S0 if S is None else S
        msy_q = q0 if msy_q is None else msy_q
    reps = int(reps)
    n = reps + (reps & 1)          # the fused kernel owns environments in pairs
    T = int(cfg["Tmax"])
    b = EnvBatch(kind, n, device=device, seed=seed, r=cfg["r"], K=cfg["K"], sigma=cfg["sigma"], C=cfg["C"], x0=cfg["x0"],
                 Tmax=T, n_actions=cfg["n_actions"], sigma_p=cfg["sigma_p"], **growth_kwargs(kind, env_kwargs))
    if reset_noise is not None:
        zr = torch.zeros((2, n), dtype=torch.float64, device=b.device)
        zr[:, :reps] = torch.as_tensor(np.asarray(reset_noise, dtype=np.float64), device=b.device).reshape(2, reps)
        b.reset(reset_noise=zr.contiguous())
    else:
        b.reset()
    obs0 = b.state.clone()
    K = b.K_env.clone() if b.K_env is not None else torch.full((n,), float(cfg["K"]), dtype=torch.float64, device=b.device)
    if noise is not None:
        z = torch.zeros((T, n), dtype=torch.float64, device=b.device)
        z[:, :reps] = torch.as_tensor(np.asarray(noise, dtype=np.float64), device=b.device).reshape(T, reps)
        noise = z
    out = b.rollout(T, policy=policy, S=float(S), msy_q=None if msy_q is None else float(msy_q), noise=noise,
                    freeze_done=True, trace=True, trace_action=True)
    obs, rew, act = out["obs"].cpu().numpy(), out["reward"].cpu().numpy(), out["action"].cpu().numpy()
    dstep = out["done_step"].cpu().numpy()[:reps]
    obs0, K = obs0.cpu().numpy(), K.cpu().numpy()
    steps = np.where(dstep < 0, T, dstep + 1).astype(np.int64)          # steps actually taken by every repetition
    # the table, built for all repetitions at once: row (t, rep) exists for t < steps[rep]
    tt, rr = np.meshgrid(np.arange(T), np.arange(reps), indexing="ij")
    keep = tt < steps[None, :]
    s_before = np.concatenate([obs0[None, :reps], obs[:T - 1, :reps]], axis=0)
    a_prev = np.concatenate([np.zeros((1, reps)), act[:T - 1, :reps]], axis=0)
    Kr = K[None, :reps]
    quota = (a_prev / cfg["n_actions"]) * Kr if kind == nat.V0 else (np.clip(a_prev, -1.0, 1.0) + 1) * Kr
    quota[0, :] = 0.0
    r_prev = np.concatenate([np.zeros((1, reps)), rew[:T - 1, :reps]], axis=0)
    order = np.argsort(rr[keep], kind="stable")                           # rows grouped by repetition, time ascending
    table = np.column_stack([tt[keep], ((s_before + 1) * Kr)[keep], quota[keep], r_prev[keep], rr[keep]])[order]
    df = DataFrame(table, columns=COLUMNS).astype({"time": int, "rep": int})
    return (df, steps) if return_steps else df


_ENV_ID_OF_KIND = {0: "fishing-v0", 1: "fishing-v1", 2: "fishing-v2", 4: "fishing-v4"}


def simulate_on_device(env, model, reps=1):
    """env.simulate(model, reps) without the host loop, when `model` is this package's escapement / msy policy of this very
    environment (fishing-v0/v1/v2/v4): all repetitions run as ONE fused rollout (simulate_mdp_fused) and the table comes
    out identical to simulate_mdp's.  Returns None when the combination is not covered -- the caller runs the host loop.

    noise="philox": the engine's counter-based stream, repetition k = environment k of a fresh key.
    noise="numpy" (the default, reference-compatible mode): the normals are drawn from numpy's GLOBAL generator in exactly
    the order and number the reference's loop would consume them -- per repetition (2 for fishing-v4's reset,
    fishing_model_error.py:42-43, then) one per step, a repetition that ends early leaves its unused draws to the next one
    (shared_env.py:51) -- and injected into the kernel, so np.random.seed(s) gives the reference's table bit for bit and
    leaves the generator where the reference leaves it.  Because the number of steps is only known afterwards, the run is
    repeated with corrected offsets when a repetition ended early: every pass fixes at least the first repetition whose
    offset was wrong, so a handful of early endings converge in a few passes; after six the host loop takes over (a
    policy that drives most repetitions extinct is inherently sequential in the reference's draw order)."""
    from . import _native as nat
    from .models.policies import escapement, msy

    kind = env._kind
    if kind not in _ENV_ID_OF_KIND or type(model) not in (escapement, msy) or getattr(model, "env", None) is not env:
        return None
    if int(reps) < 1 or int(env.Tmax) < 1:
        return None
    policy = "msy" if type(model) is msy else "escapement"
    kw = dict(sigma=env.sigma, init_state=env.init_state, Tmax=env.Tmax)
    if kind == nat.V4:
        kw.update(K_mean=env.K_mean, r_mean=env.r_mean, sigma_p=env.sigma_p)
    else:
        kw.update(r=env.r, K=env.K)
    if kind == nat.V0:
        kw["n_actions"] = env.n_actions
    if kind == nat.V2:
        kw["C"] = env.C
    common = dict(policy=policy, reps=reps, S=float(model.S), msy_q=float(model.msy) if policy == "msy" else None,
                  device=env._slot.device, **kw)
    reps, T = int(reps), int(env.Tmax)
    if env._noise_mode == "philox":
        env._epoch += 1
        df = simulate_mdp_fused(_ENV_ID_OF_KIND[kind], seed=env._seed + (env._epoch << 20), **common)
        env.reset()
        return df
    if np.random.normal is not _NUMPY_NORMAL:
        return None        # someone injects the draws call by call (the parity tests do): keep making scalar calls in order
    per_reset = 2 if kind == nat.V4 else 0
    state0 = np.random.get_state()
    steps = np.full(reps, T, dtype=np.int64)
    for _ in range(6):
        np.random.set_state(state0)
        per_rep = steps + per_reset
        z = np.random.normal(0, 1, size=int(per_rep.sum()))            # the legacy generator: same stream as scalar calls
        offs = np.concatenate([[0], np.cumsum(per_rep)[:-1]])
        noise = np.zeros((T, reps))
        tt = np.arange(T)[:, None]
        valid = tt < steps[None, :]
        noise[valid] = z[(offs[None, :] + per_reset + tt)[valid]]
        reset_noise = np.stack([z[offs], z[offs + 1]]) if per_reset else None
        df, took = simulate_mdp_fused(_ENV_ID_OF_KIND[kind], noise=noise, reset_noise=reset_noise, return_steps=True, **common)
        if np.array_equal(took, steps):
            _leave_like_the_host_loop(env)
            return df
        steps = took
    np.random.set_state(state0)
    return None


def _leave_like_the_host_loop(env):
    """After the device run: counters as after a reset (the generator has already been advanced by the draws that were used)."""
    env.years_passed = 0
    env.reward = 0
    env.harvest = 0
