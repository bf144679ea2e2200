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


def bind_cpu_to_gpu(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (same NUMA node / PCIe
    root).  Pinned host buffers allocated afterwards land on that node (Linux allocates locally), which is what keeps
    the host<->device copies of the numpy front end off the inter-socket link when 8 ranks stream at once.  Returns
    the number of cores bound to, or 0 if NVML / affinity information is not available (nothing is changed then)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device_index
            if visible:
                entry = visible.split(",")[device_index].strip()
                idx = int(entry) if entry.isdigit() else None
            h = pynvml.nvmlDeviceGetHandleByIndex(idx) if idx is not None else pynvml.nvmlDeviceGetHandleByUUID(entry)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        allowed = os.sched_getaffinity(0)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 -- best effort: no NVML, no permission, exotic topology
        return 0


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


def tree_fold_rows(t):
    """tree_sum over the rows of a 2-D tensor, level by level on whole slices (log2(rows) kernel launches instead of one per
    row): t[0::2] + t[1::2] pairs adjacent rows exactly like tree_sum does, an odd last row is carried up unchanged."""
    while t.shape[0] > 1:
        even = t.shape[0] & ~1
        nxt = t[0:even:2] + t[1:even:2]
        if t.shape[0] & 1:
            import torch

            nxt = torch.cat([nxt, t[-1:]], dim=0)
        t = nxt
    return t[0]


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
