"""High-level fused rollouts on top of EnvBatch: the loop of simulate_mdp (gym_fishing/envs/shared_env.py:29-54) with
the classical policies (gym_fishing/models/policies.py:16-31) applied in-kernel, for millions of replicate
environments, plus the statistics a study needs (mean / std of the episode return, extinction counts) combined across
GPUs with ONE small collective.  Environments are sharded by contiguous global-id ranges; Philox noise is keyed by
global id, so the numbers do not depend on the number of GPUs."""
from __future__ import annotations

from . import _native as nat
from .distributed import allreduce_stats, shard_range
from .engine import EnvBatch
from .vec_env import growth_kwargs, normalise_kwargs


def _rank_world():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # noqa: BLE001
        pass
    return 0, 1


def _finish(stats, n_total):
    """[groups, 3] sums -> dict of mean / std (population) / count tensors."""
    s, s2, nd = stats[:, 0], stats[:, 1], stats[:, 2]
    mean = s / n_total
    var = (s2 / n_total - mean * mean).clamp_min(0.0)
    return dict(sum=s, sumsq=s2, mean=mean, std=var.sqrt(), n_done=nd, n=n_total)


def rollout_stats(env_id, num_envs, T, *, policy="escapement", S=None, msy_q=None, device="cuda", seed=0,
                  freeze_done=True, deterministic=True, keep_per_env=False, **env_kwargs):
    """`num_envs` (global) replicate episodes of `T` fused steps under a built-in policy; returns return statistics
    over ALL ranks.  BASELINE.json config 3: fishing-v2, 2^24 envs over 8 GPUs, constant escapement, T = Tmax = 1000."""
    kind = nat.KIND_OF_ID[env_id]
    cfg = normalise_kwargs(kind, env_kwargs)
    rank, world = _rank_world()
    lo, hi = shard_range(num_envs, rank, world, align=2)
    b = EnvBatch(kind, hi - lo, device=device, env_id0=lo, seed=seed, r=cfg["r"], K=cfg["K"], sigma=cfg["sigma"], C=cfg["C"],
                 x0=cfg["x0"], Tmax=cfg["Tmax"], n_actions=cfg["n_actions"], sigma_p=cfg["sigma_p"],
                 **growth_kwargs(kind, env_kwargs))
    b.reset()
    out = b.rollout(T, policy=policy, S=S, msy_q=msy_q, freeze_done=freeze_done)
    # fixed 256-wide tree (partition-invariant) when EVERY rank's shard is a multiple of 256 --
    # decided from the global size, so that all ranks agree on the shape of the block they exchange
    groups = 256 if num_envs % (256 * world) == 0 else 1
    local = b.reduce_stats(out["ret"], out["done_step"], groups=groups)
    total = allreduce_stats(local, deterministic=deterministic)
    from .distributed import tree_fold_rows

    folded = tree_fold_rows(total).unsqueeze(0)      # the 256 group columns, same adjacent-pair tree as on the device
    res = _finish(folded, num_envs)
    res = {k: (v[0] if hasattr(v, "shape") and v.ndim else v) for k, v in res.items()}
    if keep_per_env:
        res.update(ret=out["ret"], done_step=out["done_step"], final_state=b.state, env_range=(lo, hi))
    return res


def escapement_sweep(env_id="fishing-v1", *, levels=1024, replicates=16384, T=1000, S_levels=None, device="cuda", seed=0,
                     deterministic=True, **env_kwargs):
    """Grid sweep of the constant-escapement level (BASELINE.json config 5): `levels` escapement targets x `replicates`
    noise replicates.  Environment id = replicate * levels + level; the Philox stream is keyed by the REPLICATE
    (noise_id_div = levels), i.e. common random numbers across levels.  Replicates are sharded over ranks; each GPU
    reduces its replicates per level with a fixed pairwise tree, then one collective combines the [levels, 3] blocks.
    Returns dict(S, mean, std, n_done, ...) of length-`levels` tensors (identical on every rank)."""
    import torch

    kind = nat.KIND_OF_ID[env_id]
    if kind == nat.V0:
        raise ValueError("the sweep uses the continuous get_action path; use fishing-v1/v2/v4")
    cfg = normalise_kwargs(kind, env_kwargs)
    rank, world = _rank_world()
    r_lo, r_hi = shard_range(replicates, rank, world, align=2)
    n_local = (r_hi - r_lo) * levels
    dev = torch.device(device)
    if S_levels is None:
        K = float(cfg["K"])
        S_levels = torch.arange(levels, dtype=torch.float64, device=dev) * K / (levels - 1)   # S_j = K * j / (levels-1)
    else:
        S_levels = torch.as_tensor(S_levels, dtype=torch.float64, device=dev)
    S_env = S_levels.repeat(r_hi - r_lo).contiguous()
    b = EnvBatch(kind, n_local, device=dev, env_id0=r_lo * levels, seed=seed, r=cfg["r"], K=cfg["K"], sigma=cfg["sigma"],
                 C=cfg["C"], x0=cfg["x0"], Tmax=cfg["Tmax"], sigma_p=cfg["sigma_p"], **growth_kwargs(kind, env_kwargs))
    b.reset()
    out = b.rollout(T, policy="escapement", S_env=S_env, noise_id_div=levels, freeze_done=True)
    local = b.reduce_stats(out["ret"], out["done_step"], groups=levels)
    total = allreduce_stats(local, deterministic=deterministic)
    res = _finish(total, replicates)
    res["S"] = S_levels
    return res
