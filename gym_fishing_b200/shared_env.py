"""Rollout recording and reporting helpers with the reference's outputs (gym_fishing/envs/shared_env.py:8-127):
simulate_mdp / simulate_mdp_vec tables with columns [time, state, action, reward, rep], estimate_policyfn, and the
matplotlib plots (only if matplotlib is installed)."""
from __future__ import annotations

import numpy as np
from pandas import DataFrame

COLUMNS = ["time", "state", "action", "reward", "rep"]


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


def simulate_mdp_fused(env_id, policy="escapement", reps=1, *, S=None, msy_q=None, noise=None, seed=0, device="cuda",
                       **env_kwargs):
    """simulate_mdp (shared_env.py:29-54) for one of the built-in policies with ALL repetitions run as one fused rollout on
    the device: `reps` environments, at most Tmax steps, an environment stops at its first done (the `break` at :51).
    Returns the same table [time, state, action, reward, rep] (same one-step lag of the quota / reward columns).

    policy: "escapement" (S = BMSY by default) or "msy" (msy_q = surplus production at BMSY by default);
    noise: optional [Tmax, reps] standard normals to inject (else the engine's Philox stream keyed by `seed`)."""
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
    n = int(reps) + (int(reps) & 1)          # the fused kernel owns environments in pairs
    T = int(cfg["Tmax"])
    b = EnvBatch(kind, n, device=device, seed=seed, r=cfg["r"], K=cfg["K"], sigma=cfg["sigma"], C=cfg["C"], x0=cfg["x0"],
                 Tmax=T, n_actions=cfg["n_actions"], sigma_p=cfg["sigma_p"], **growth_kwargs(kind, env_kwargs))
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
    dstep = out["done_step"].cpu().numpy()
    obs0, K = obs0.cpu().numpy(), K.cpu().numpy()
    chunks = []
    for rep in range(int(reps)):
        L = T if dstep[rep] < 0 else int(dstep[rep]) + 1           # steps actually taken
        s_before = np.concatenate([[obs0[rep]], obs[:L - 1, rep]])
        a_prev = act[:L - 1, rep]
        if kind == nat.V0:
            quota = (a_prev / cfg["n_actions"]) * K[rep]
        else:
            quota = (np.clip(a_prev, -1.0, 1.0) + 1) * K[rep]
        chunks.append(np.column_stack([np.arange(L), (s_before + 1) * K[rep], np.concatenate([[0.0], quota]),
                                       np.concatenate([[0.0], rew[:L - 1, rep]]), np.full(L, rep)]))
    df = DataFrame(np.concatenate(chunks), columns=COLUMNS)
    return df.astype({"time": int, "rep": int})
