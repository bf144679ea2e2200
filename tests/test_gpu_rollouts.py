"""The two replicate-rollout workloads of BASELINE.json (configs 3 and 5) at test size, against known answers from the
reference and the oracle."""
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
