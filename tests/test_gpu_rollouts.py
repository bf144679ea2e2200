"""The two replicate-rollout workloads of BASELINE.json (configs 3 and 5) at test size, against known answers from the
reference and the oracle."""
import os

import numpy as np
import pytest

from oracle import fishing_oracle as fo
from oracle import philox_ref as pr

pytestmark = pytest.mark.gpu


def test_sweep_sigma0_known_answers_from_the_reference():
    """sigma = 0, T = 1000, fishing-v1 defaults, continuous get_action path: the reference gives
    S = 0.25 / 0.50 / 0.75 -> sum of rewards 0x1.c58cccccccc46p+5 / 0x1.2cb3333333397p+6 / 0x1.c18cccccccc48p+5."""
    import gym_fishing_b200 as gf

    res = gf.escapement_sweep("fishing-v1", levels=5, replicates=4, T=1000, S_levels=[0.0, 0.25, 0.5, 0.75, 1.0],
                              sigma=0.0, Tmax=1000)
    mean = res["mean"].cpu().numpy()
    assert [float(m).hex() for m in mean[1:4]] == ["0x1.c58cccccccc46p+5", "0x1.2cb3333333397p+6", "0x1.c18cccccccc48p+5"]
    assert np.all(res["std"].cpu().numpy() < 1e-6)          # replicates identical without noise
    assert float(res["n_done"][0]) == 4.0                    # S = 0: everything is harvested in the first step


def test_sweep_with_noise_matches_oracle_and_uses_common_random_numbers():
    import gym_fishing_b200 as gf

    levels, reps, T = 16, 64, 100
    res = gf.escapement_sweep("fishing-v1", levels=levels, replicates=reps, T=T, sigma=0.1, Tmax=1000, seed=5)
    z = np.stack([pr.step_noise(5, np.arange(reps, dtype=np.uint64), np.uint64(t)) for t in range(T)])
    S = np.arange(levels) / (levels - 1.0)
    for j in (0, 3, 8, 15):
        e = fo.rollout_vec("v1", T, reps, r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=1000, policy="escapement", S=S[j], noise=z,
                           record=False)
        assert abs(float(res["mean"][j]) - e["ret"].mean()) <= 1e-9 * max(1.0, abs(e["ret"].mean()))
        assert float(res["n_done"][j]) == float((e["done_step"] >= 0).sum())
    assert float(res["std"][8]) > 0


def test_rollout_stats_cfg3_shape_at_test_size():
    """fishing-v2, constant escapement at the reference's S, sigma = 0: every replicate equals KAT-3b."""
    import gym_fishing_b200 as gf

    S = float.fromhex("0x1.9374bcp-1")
    res = gf.rollout_stats("fishing-v2", 4096, 1000, policy="escapement", S=S, sigma=0.0, Tmax=1000, keep_per_env=True)
    kat = float.fromhex("0x1.d06ecd09cebeep+3")
    assert abs(float(res["mean"]) - kat) <= 1e-12 * kat and float(res["n_done"]) == 0.0 and float(res["std"]) < 1e-6
    assert abs(float(res["final_state"][0]) - float.fromhex("-0x1.9457f2e95d710p-3")) < 1e-12
    # with noise: statistics equal a numpy reduction of the per-env returns
    res = gf.rollout_stats("fishing-v2", 1 << 16, 200, policy="escapement", S=S, sigma=0.05, Tmax=1000, keep_per_env=True)
    ret = res["ret"].cpu().numpy()
    assert abs(float(res["mean"]) - ret.mean()) <= 1e-12 * ret.mean()
    assert abs(float(res["std"]) - ret.std()) <= 1e-9
    assert float(res["n_done"]) == float((res["done_step"].cpu().numpy() >= 0).sum())


def test_max_size_2pow24_shards_equal_whole_batch():
    """BASELINE.json config 3 at its full per-box size (2^24 fishing-v2 envs) on one GPU for a few steps: eight shards with
    global ids (what the 8 ranks run) reproduce the whole batch bit for bit, statistics included."""
    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.distributed import tree_sum
    from gym_fishing_b200.engine import EnvBatch

    n, T, S = 1 << 24, 12, float.fromhex("0x1.9374bcp-1")
    cfg = dict(r=0.3, K=1.0, sigma=0.05, C=0.5, x0=0.75, Tmax=1000, seed=0)
    whole = EnvBatch(nat.V2, n, **cfg)
    whole.reset()
    out = whole.rollout(T, policy="escapement", S=S, freeze_done=True)
    stats = whole.reduce_stats(out["ret"], out["done_step"], groups=256)
    parts = []
    shard = n // 8
    for g in range(8):
        b = EnvBatch(nat.V2, shard, env_id0=g * shard, **cfg)
        b.reset()
        o = b.rollout(T, policy="escapement", S=S, freeze_done=True)
        assert torch.equal(o["ret"], out["ret"][g * shard:(g + 1) * shard])
        assert torch.equal(b.state, whole.state[g * shard:(g + 1) * shard])
        parts.append(b.reduce_stats(o["ret"], o["done_step"], groups=256))
    assert torch.equal(tree_sum(parts), stats)
    assert float(stats[:, 0].sum()) > 0 and torch.isfinite(whole.state).all()


def test_full_size_cfg4_v4_rollout_buffer_2pow22():
    """BASELINE.json config 4 at its full size: fishing-v4, 2^22 environments, a 128-step rollout buffer filled by K1
    (per-env K, r redrawn from the Philox stream at every auto-reset).  (i) the fused kernel with per-step rows == K1 x 128
    bit for bit; (ii) two half shards with global ids == the whole batch; (iii) four blocks of 512 environments replayed
    through the oracle over all 128 steps with the same actions and the engine's own exported reset draws."""
    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.engine import EnvBatch

    n, T = 1 << 22, 128
    cfg = dict(r=0.3, K=1.0, sigma=0.0, sigma_p=0.1, x0=0.75, Tmax=100, seed=3)
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.rand((T, n), dtype=torch.float64, device="cuda", generator=g) * 0.5 - 1.0
    whole = EnvBatch(nat.V4, n, auto_reset=True, **cfg)
    whole.reset()
    K0, r0 = whole.K_env.clone(), whole.r_env.clone()
    obs = torch.empty(T, n, dtype=torch.float64, device="cuda")
    rew = torch.empty_like(obs)
    done = torch.empty(T, n, dtype=torch.uint8, device="cuda")
    whole.step_many(acts, obs, rew, done)
    fused = EnvBatch(nat.V4, n, auto_reset=True, **cfg)
    fused.reset()
    out = fused.rollout(T, policy="actions", actions=acts, trace=True)
    assert torch.equal(out["obs"], obs) and torch.equal(out["reward"], rew) and torch.equal(out["done"], done)
    assert torch.equal(fused.K_env, whole.K_env) and torch.equal(fused.r_env, whole.r_env)
    del out, fused
    half, Th = n // 2, 32
    for lo in (0, half):
        part = EnvBatch(nat.V4, half, env_id0=lo, auto_reset=True, **cfg)
        part.reset()
        o2 = torch.empty(Th, half, dtype=torch.float64, device="cuda")
        r2 = torch.empty_like(o2)
        d2 = torch.empty(Th, half, dtype=torch.uint8, device="cuda")
        part.step_many(acts[:Th, lo:lo + half].contiguous(), o2, r2, d2)
        assert torch.equal(o2, obs[:Th, lo:lo + half]) and torch.equal(d2, done[:Th, lo:lo + half])
        del part, o2, r2, d2
    m = 512
    for lo in (0, 1_000_000, 2_718_282 & ~1, n - m):
        src = EnvBatch(nat.V4, m, env_id0=lo, **cfg)
        rn = np.empty((T + 1, 2, m))
        rn[0] = 0.0                                     # the explicit reset's draws are taken from the engine's K0, r0 below
        rn[1:] = src.noise(T, stream=nat.STREAM_AUTORESET, step0=0).cpu().numpy()
        exp = fo.rollout_vec("v4", T, m, r=r0[lo:lo + m].cpu().numpy(), K=K0[lo:lo + m].cpu().numpy(), sigma=0.0, x0=0.75,
                             Tmax=100, policy="actions", actions=acts[:, lo:lo + m].cpu().numpy(), auto_reset=True,
                             reset_noise=rn, K_mean=1.0, r_mean=0.3, sigma_p=0.1)
        assert np.array_equal(obs[:, lo:lo + m].cpu().numpy(), exp["obs"]), lo
        assert np.array_equal(rew[:, lo:lo + m].cpu().numpy(), exp["reward"]) and np.array_equal(done[:, lo:lo + m].cpu().numpy(), exp["done"])
        assert np.array_equal(whole.K_env[lo:lo + m].cpu().numpy(), exp["K"]) and exp["done"].sum() > m


def test_full_size_cfg5_sweep_1024_levels():
    """BASELINE.json config 5 at one GPU's share of its full size: 1024 escapement levels x 2048 noise replicates x 1000
    steps, common random numbers (warp-cooperative sweep kernel).  (i) two half shards of the replicate axis with global
    ids == the whole batch, per environment and in the reduced statistics; (ii) two replicates x all 1024 levels replayed
    through the oracle with the engine's exported noise; (iii) the sigma = 0 known answers of the reference at the three
    levels that hit S = 0.25 / 0.5 / 0.75 exactly are reproduced by the levels of a 5-point grid (test above) -- here the
    1024-level grid must give the same mean return for every replicate when sigma = 0."""
    import torch

    from gym_fishing_b200 import _native as nat
    from gym_fishing_b200.engine import EnvBatch

    levels, reps, T = 1024, 2048, 1000
    n = levels * reps
    cfg = dict(r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=1000, seed=0)
    S_levels = torch.arange(levels, dtype=torch.float64, device="cuda") / (levels - 1)

    def run(rep_lo, rep_n, **over):
        c = dict(cfg, **over)
        b = EnvBatch(nat.V1, rep_n * levels, env_id0=rep_lo * levels, **c)
        b.reset()
        out = b.rollout(T, policy="escapement", S_env=S_levels.repeat(rep_n).contiguous(), noise_id_div=levels, freeze_done=True)
        return b, out

    b, whole = run(0, reps)
    stats = b.reduce_stats(whole["ret"], whole["done_step"], groups=levels)
    parts = []
    for lo in (0, reps // 2):
        bp, o = run(lo, reps // 2)
        sl = slice(lo * levels, (lo + reps // 2) * levels)
        assert torch.equal(o["ret"], whole["ret"][sl]) and torch.equal(o["done_step"], whole["done_step"][sl])
        parts.append(bp.reduce_stats(o["ret"], o["done_step"], groups=levels))
    assert torch.equal(parts[0] + parts[1], stats)          # aligned power-of-two shards combine exactly
    # oracle replay of replicates 6 and 7 (one Philox pair) at every level
    z = EnvBatch(nat.V1, 2, env_id0=6, **cfg).noise(T, step0=0).cpu().numpy()          # [T, 2]: replicate 6 -> z0, 7 -> z1
    for k in (6, 7):
        e = fo.rollout_vec("v1", T, levels, r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=1000, policy="escapement",
                           S=S_levels.cpu().numpy(), noise=np.repeat(z[:, k - 6:k - 5], levels, axis=1), record=False)
        got = whole["ret"][k * levels:(k + 1) * levels].cpu().numpy()
        assert np.array_equal(got, e["ret"]), k
        assert np.array_equal(whole["done_step"][k * levels:(k + 1) * levels].cpu().numpy(), e["done_step"])
    # sigma = 0: every replicate of a level is the same deterministic trajectory
    _, det = run(0, 4, sigma=0.0)
    r4 = det["ret"].view(4, levels)
    assert torch.equal(r4[0], r4[1]) and torch.equal(r4[0], r4[3])
    assert float(r4[0].max()) > 75.0


NCCL_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch
import gym_fishing_b200 as gf
from gym_fishing_b200.distributed import init_from_env
rank, world, local = init_from_env()          # backend nccl: one process per GPU
assert world == 2 and torch.distributed.get_backend() == "nccl"
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
S = float.fromhex("0x1.9374bcp-1")
a = gf.rollout_stats("fishing-v2", 1 << 18, 200, policy="escapement", S=S, device=dev, sigma=0.05, Tmax=1000)
b = gf.escapement_sweep("fishing-v1", levels=256, replicates=512, T=150, device=dev, sigma=0.1, Tmax=1000)
plain = gf.rollout_stats("fishing-v2", 1 << 18, 200, policy="escapement", S=S, device=dev, sigma=0.05, Tmax=1000, deterministic=False)
if rank == 0:
    print("RESULT " + json.dumps({"sum": float(a["sum"]).hex(), "sumsq": float(a["sumsq"]).hex(), "n_done": float(a["n_done"]),
                                  "sweep_sum": float(b["sum"].sum()).hex(), "sweep_sumsq_17": float(b["sumsq"][17]).hex(),
                                  "sweep_mean_200": float(b["mean"][200]).hex(), "plain_sum": float(plain["sum"])}))
torch.distributed.barrier()
torch.distributed.destroy_process_group()
'''


def test_two_ranks_over_nccl_reproduce_the_single_gpu_statistics(tmp_path):
    """cfg 3 / cfg 5 in small: two processes, one GPU each, NCCL all-gather of the statistics (the only collective of the
    path) -- the sums must be bit-identical to the one-GPU run, because the Philox streams are keyed by GLOBAL environment id and
    the reduction tree is fixed.  Skipped on a one-GPU box; the 8-GPU records are profiles/configs_8gpu_*.jsonl."""
    import json
    import socket
    import subprocess
    import sys

    import torch

    import gym_fishing_b200 as gf

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(NCCL_WORKER)
    for attempt in range(3):
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            port = sock.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", str(port), str(script), root]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        if res.returncode == 0 or "AssertionError" in res.stderr:
            break
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    two = json.loads([l for l in res.stdout.splitlines() if l.startswith("RESULT ")][0][7:])
    S = float.fromhex("0x1.9374bcp-1")
    a = gf.rollout_stats("fishing-v2", 1 << 18, 200, policy="escapement", S=S, sigma=0.05, Tmax=1000)
    b = gf.escapement_sweep("fishing-v1", levels=256, replicates=512, T=150, sigma=0.1, Tmax=1000)
    assert two["sum"] == float(a["sum"]).hex() and two["sumsq"] == float(a["sumsq"]).hex() and two["n_done"] == float(a["n_done"])
    assert two["sweep_sum"] == float(b["sum"].sum()).hex() and two["sweep_sumsq_17"] == float(b["sumsq"][17]).hex()
    assert two["sweep_mean_200"] == float(b["mean"][200]).hex()
    assert abs(two["plain_sum"] - float(a["sum"])) <= 1e-9 * abs(float(a["sum"]))      # plain NCCL all-reduce: same value, free order
