"""Multi-GPU plumbing: one process per GPU (torchrun), environments sharded by contiguous global-id ranges, and ONE
small collective at the end of a rollout for the return statistics.  There is no exchange step in the data path:
environments are independent and the Philox noise is keyed by GLOBAL environment id, so every trajectory is
identical whatever the number of GPUs."""
from __future__ import annotations

import os


def shard_range(n_total, rank, world, align=2):
    """Contiguous [start, stop) of global env ids owned by `rank`; boundaries are multiples of `align` (the kernels own
    environments in pairs, sweeps in 2*levels blocks) and the shards differ in size by at most one `align` block."""
    if n_total % align != 0:
        raise ValueError(f"n_total={n_total} must be a multiple of {align}")
    blocks = n_total // align
    base, extra = divmod(blocks, world)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start * align, stop * align


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_*).
    Returns (rank, world, local_rank).  Single-process runs return (0, 1, 0) without initialising anything."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def tree_sum(parts):
    """Pairwise (adjacent, level by level) sum of a list of tensors -- the same shape of tree the on-device
    reduction uses, so shard results combine to exactly the single-GPU value when the shards are aligned
    power-of-two blocks."""
    parts = list(parts)
    while len(parts) > 1:
        nxt = [parts[i] + parts[i + 1] for i in range(0, len(parts) - 1, 2)]
        if len(parts) % 2:
            nxt.append(parts[-1])
        parts = nxt
    return parts[0]


def allreduce_stats(stats, deterministic=True):
    """Combine per-rank statistics tensors (any shape, float64).  deterministic=True: all_gather + fixed pairwise tree
    in rank order (bit-identical for 1/2/4/8 GPUs); False: a plain NCCL/gloo all_reduce(SUM)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats
    if not deterministic:
        out = stats.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out
    world = dist.get_world_size()
    gathered = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats.contiguous())
    return tree_sum(gathered)


def barrier_max_ms(ms, device=None):
    """Max over ranks of a per-rank duration (device-timed milliseconds)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
