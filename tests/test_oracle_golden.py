"""Pins oracle/fishing_oracle.py to the golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only.  Bit-exact everywhere except fishing-v2, whose np.exp may differ by an
ulp between numpy builds (tolerance 1e-12 relative, stated by BASELINE.json's north_star)."""
import json
import os

import numpy as np
import pytest

from oracle import fishing_oracle as fo

G = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(G, "kat.json")))
Z = np.load(os.path.join(G, "Z2024.npy"))
uh = fo.unhex


def close(a, b, kind):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if kind == "v2":
        return np.allclose(a, b, rtol=1e-12, atol=1e-300, equal_nan=True)
    return np.array_equal(a, b, equal_nan=True)


def test_generator_noise_is_the_documented_one():
    assert [float(z).hex() for z in Z[:3]] == KAT["Z_head"]
    assert KAT["Z_head"][0] == "0x1.07632a01e361cp+0"


@pytest.mark.parametrize("kind", ["v0", "v1", "v2", "v4"])
def test_step_trajectories_match_reference(kind):
    d = np.load(os.path.join(G, f"traj_{kind}.npz"))
    T, M = d["actions"].shape
    s = d["obs0"].copy()
    years = np.zeros(M, dtype=np.int64)
    K = d["K_env"][0] if kind == "v4" else d["p_K"]
    r = d["r_env"][0] if kind == "v4" else d["p_r"]
    for t in range(T):
        s, rew, done, years = fo.step_vec(kind, s, d["actions"][t], d["noise"][t], years, r=r, K=K,
                                          sigma=d["p_sigma"], C=d["p_C"], n_actions=d["p_n_actions"],
                                          Tmax=d["p_Tmax"])
        assert close(s, d["obs"][t], kind), (kind, t)
        assert close(rew, d["reward"][t], kind), (kind, t)
        assert np.array_equal(done.astype(np.uint8), d["done"][t]), (kind, t)
    assert d["done"].any() and not d["done"].all()


@pytest.mark.parametrize("kind", ["v0", "v1", "v2", "v4"])
def test_v4_and_base_reset(kind):
    d = np.load(os.path.join(G, f"traj_{kind}.npz"))
    if kind == "v4":
        K, r = fo.v4_draw(d["p_K"], d["p_r"], d["p_sigma_p"], d["reset_noise"][0, 0], d["reset_noise"][0, 1])
        assert np.array_equal(K, d["K_env"][0]) and np.array_equal(r, d["r_env"][0])
        assert np.array_equal(fo.reset_state("v4", d["p_x0"], K), d["obs0"])
    else:
        assert np.array_equal(fo.reset_state(kind, d["p_x0"], d["p_K"]), d["obs0"])


@pytest.mark.parametrize("kind", ["v0", "v1", "v2", "v4"])
def test_autoreset_rollout_matches_reference(kind):
    d = np.load(os.path.join(G, f"autoreset_{kind}.npz"))
    T, M = d["actions"].shape
    K0 = d["K_env"][0] if kind == "v4" else d["p_K"]
    r0 = d["r_env"][0] if kind == "v4" else d["p_r"]
    out = fo.rollout_vec(kind, T, M, r=r0, K=K0, sigma=d["p_sigma"], C=d["p_C"], x0=d["p_x0"],
                         n_actions=d["p_n_actions"], Tmax=d["p_Tmax"], policy="actions", actions=d["actions"],
                         noise=d["noise"], auto_reset=True, reset_noise=d["reset_noise"], K_mean=d["p_K"],
                         r_mean=d["p_r"], sigma_p=d["p_sigma_p"], s0=d["obs0"])
    assert close(out["obs"], d["obs"], kind)
    assert close(out["reward"], d["reward"], kind)
    assert np.array_equal(out["done"], d["done"])
    assert close(out["term"], d["term"], kind)
    assert d["done"].sum() > M  # resets really happened, several per env
    if kind == "v4":
        assert np.array_equal(out["K"], d["K_env"][T]) and np.array_equal(out["r"], d["r_env"][T])


def test_kat1_v0_escapement_1000_steps():
    k = KAT["kat1"]
    S = fo.bmsy("v0", r=0.3, K=1)
    assert S.dtype == np.float32 and float(S).hex() == k["S"] == "0x1.ff97240000000p-2"
    out = fo.rollout_vec("v0", 1000, 1, r=0.3, K=1.0, sigma=0.0, x0=0.75, n_actions=100, Tmax=1000,
                         policy="escapement", S=S)
    acts = out["action"][:, 0].astype(int).tolist()
    assert acts == k["actions"]
    assert sum(a * (t + 1) for t, a in enumerate(acts)) == k["wsum"] == 3753467
    assert float(out["final_state"][0]).hex() == k["final_s"]
    tot = 0.0
    for v in out["reward"][:, 0]:
        tot += float(v)
    assert tot.hex() == k["sumR"]
    assert not out["done"].any() and int(out["years"][0]) == 1000
    # step 1001 trips the strict '>' time limit (base_fishing_env.py:76)
    a = fo.escapement_action("v0", out["final_state"], S, 1.0, 100)
    _, _, done, _ = fo.step_vec("v0", out["final_state"], a, 0.0, out["years"], r=0.3, K=1.0, sigma=0.0, Tmax=1000)
    assert bool(done[0]) is True and k["done_1001"] is True


def test_kat2_v1_constant_action():
    k = KAT["kat2"]
    out = fo.rollout_vec("v1", 101, 1, r=0.3, K=1.0, sigma=0.0, Tmax=100, policy="actions",
                         actions=np.full((101, 1), -0.85))
    assert [float(v).hex() for v in out["obs"][:, 0]] == k["s"]
    assert [float(v).hex() for v in out["reward"][:, 0]] == k["r"]
    assert [bool(v) for v in out["done"][:, 0]] == k["done"]
    assert k["s"][0] == "-0x1.4fdf3b645a1ccp-2" and k["done"][99] is True


def test_kat2b_overharvest_then_nothing():
    (s1, r1, d1), (s2, r2, d2) = KAT["kat2b"]
    out = fo.rollout_vec("v1", 2, 1, r=0.3, K=1.0, sigma=0.0, policy="actions", actions=np.array([[0.1], [-1.0]]))
    assert [float(v).hex() for v in out["obs"][:, 0]] == [s1, s2]
    assert [float(v).hex() for v in out["reward"][:, 0]] == [r1, r2]
    assert [bool(v) for v in out["done"][:, 0]] == [d1, d2] == [True, True]


def test_kat3_tipping_point_upstream_assertions():
    """The only numeric assertions upstream: tests/test-envs.py:99-100 and :105-106."""
    k = KAT["kat3"]
    for x0, key, cmp in ((0.75, "x_075", lambda x: x >= 0.75), (0.3, "x_030", lambda x: x <= 0.3)):
        a = fo.action_from_quota("v2", 0, 1.0)
        s, *_ = fo.step_vec("v2", fo.reset_state("v2", x0, 1.0), a, 0.0, 0, r=0.3, K=1.0, sigma=0.0, C=0.5)
        x = float(fo.population_from_state(s, 1.0))
        assert cmp(x)
        assert abs(x - uh(k[key])) <= 1e-12 * abs(x)


def test_kat3b_v2_escapement():
    k = KAT["kat3b"]
    S = fo.bmsy("v2", r=0.3, K=1, C=0.5)
    assert float(S).hex() == k["S"] == "0x1.9374bc0000000p-1"
    out = fo.rollout_vec("v2", 1000, 1, r=0.3, K=1.0, sigma=0.0, C=0.5, Tmax=1000, policy="escapement", S=S)
    assert abs(out["final_state"][0] - uh(k["final_s"])) <= 1e-12 * abs(uh(k["final_s"]))
    assert abs(out["total"][0] - uh(k["sumR"])) <= 1e-12 * uh(k["sumR"])
    assert not out["done"].any()


def test_kat5_v1_injected_noise():
    k = KAT["kat5"]
    out = fo.rollout_vec("v1", 1000, 1, r=0.3, K=1.0, sigma=0.05, Tmax=1000, policy="actions",
                         actions=np.full((1000, 1), -0.95), noise=Z[:1000, None])
    assert float(out["final_state"][0]).hex() == k["final_s"]
    tot = 0.0
    for v in out["reward"][:, 0]:
        tot += float(v)
    assert tot.hex() == k["sumR"]


@pytest.mark.parametrize("name,kind", [("kat6", "v0"), ("kat7", "v2")])
def test_kat6_kat7_escapement_with_noise(name, kind):
    k = KAT[name]
    S = fo.bmsy(kind, r=0.3, K=1, C=0.5, z=Z[0])          # BMSY eats Z[0]
    assert float(S).hex() == k["S"]
    out = fo.rollout_vec(kind, 1000, 1, r=0.3, K=1.0, sigma=0.05, C=0.5, Tmax=1000, policy="escapement", S=S,
                         noise=Z[1:1001, None])
    tot = 0.0
    for v in out["reward"][:, 0]:
        tot += float(v)
    if kind == "v0":
        assert float(out["final_state"][0]).hex() == k["final_s"] and tot.hex() == k["sumR"]
        assert sum(int(a) * (t + 1) for t, a in enumerate(out["action"][:, 0])) == k["wsum"]
    else:
        assert abs(out["final_state"][0] - uh(k["final_s"])) <= 1e-12 * abs(uh(k["final_s"]))
        assert abs(tot - uh(k["sumR"])) <= 1e-12 * uh(k["sumR"])
    assert not out["done"].any()


def test_kat8_v4_parameter_draws():
    k = KAT["kat8"]
    K, r = fo.v4_draw(1.0, 0.3, 0.1, Z[0], Z[1])
    assert (float(K).hex(), float(r).hex()) == (k["ctor_K"], k["ctor_r"])
    K, r = fo.v4_draw(1.0, 0.3, 0.1, Z[2], Z[3])
    assert (float(K).hex(), float(r).hex()) == (k["reset_K"], k["reset_r"])
    assert float(fo.reset_state("v4", 0.75, K)).hex() == k["reset_obs"]
    out = fo.rollout_vec("v4", 147, 1, r=r, K=K, sigma=0.05, x0=0.75, Tmax=1000, policy="actions",
                         actions=np.full((147, 1), -0.95), noise=Z[4:151, None])
    assert float(out["obs"][0, 0]).hex() == k["s0"] and float(out["reward"][0, 0]).hex() == k["r0"]
    assert int(out["done_step"][0]) == k["done_t"] == 146
    tot = 0.0
    for v in out["reward"][:, 0]:
        tot += float(v)
    assert tot.hex() == k["sumR"]


def test_edge_cases_nan_out_of_range_post_done():
    for case in KAT["edge"]:
        kind = "v" + case["env"].split("-v")[1]
        kw = dict(r=0.3, K=1, sigma=0.0, C=0.5, n_actions=100, Tmax=100, init_state=0.75)
        kw.update(case["kw"])
        z = Z[case["noise_offset"]:]
        s = fo.reset_state(kind, kw["init_state"], kw["K"])
        assert float(s).hex() == case["obs0"]
        years = 0
        for i, (a, (es, er, ed)) in enumerate(zip(case["actions"], case["steps"])):
            a = a if kind == "v0" else uh(a)
            s, rew, done, years = fo.step_vec(kind, s, a, z[i], years, r=kw["r"], K=kw["K"], sigma=kw["sigma"],
                                              C=kw["C"], n_actions=kw["n_actions"], Tmax=kw["Tmax"])
            assert close(s, uh(es), kind) and close(rew, uh(er), kind) and bool(done) == ed, (case, i)


def test_policy_constants():
    rows = json.load(open(os.path.join(G, "policy.json")))
    for row in rows:
        kind = "v" + row["env"].split("-v")[1]
        kw = row["kw"]
        z = Z[row["noise_offset"]:]
        if kind == "v4":
            # ctor draws z[0:2]; BMSY's population_draw eats z[2]; its env.reset() redraws from z[3:5]
            K, r = fo.v4_draw(kw.get("K_mean", 1.0), kw.get("r_mean", 0.3), kw.get("sigma_p", 0.1), z[0], z[1])
            assert (float(K).hex(), float(r).hex()) == (row["K_ctor"], row["r_ctor"])
            S = fo.bmsy("v4", r=r, K=K, z=z[2])
            assert S.dtype == np.float64
            K2, r2 = fo.v4_draw(kw.get("K_mean", 1.0), kw.get("r_mean", 0.3), kw.get("sigma_p", 0.1), z[3], z[4])
            assert (float(K2).hex(), float(r2).hex()) == (row["K_bmsy"], row["r_bmsy"])
            q = fo.msy_quota("v4", S, r=r2, K=K2)     # msy() draws at the K, r left by BMSY's reset
        else:
            r, K, C = kw.get("r", 0.3), kw.get("K", 1), kw.get("C", 0.5)
            S = fo.bmsy(kind, r=r, K=K, C=C, z=z[0])
            assert S.dtype == np.float32
            q = fo.msy_quota(kind, S, r=r, K=K, C=C)
        assert float(S).hex() == row["S"], row
        assert str(np.asarray(q).dtype) == row["msy_dtype"]
        if kind == "v2":
            assert abs(float(q) - uh(row["msy"])) <= 2e-7 * abs(uh(row["msy"]))   # float32 exp, 1 ulp of f32
        else:
            assert float(q).hex() == row["msy"], row


def test_simulate_tables():
    tabs = json.load(open(os.path.join(G, "simulate.json")))
    for tab in tabs:
        kind = "v" + tab["env"].split("-v")[1]
        kw = dict(r=0.3, K=1, sigma=0.0, C=0.5)
        kw.update(tab["kw"])
        z = Z[tab["noise_offset"]:]
        S = fo.bmsy(kind, r=kw["r"], K=kw["K"], C=kw["C"], z=z[0])
        if tab["policy"] == "msy":
            q = fo.msy_quota(kind, S, r=kw["r"], K=kw["K"], C=kw["C"], z=z[1])
            rows = fo.simulate_table(kind, "msy", reps=tab["reps"], r=kw["r"], K=kw["K"], sigma=kw["sigma"],
                                     C=kw["C"], msy_q=q, noise=z[2:], f32_actions=True)
        else:
            rows = fo.simulate_table(kind, "escapement", reps=tab["reps"], r=kw["r"], K=kw["K"], sigma=kw["sigma"],
                                     C=kw["C"], S=S, noise=z[1:])
        assert len(rows) == len(tab["rows"])
        for mine, ref in zip(rows, tab["rows"]):
            assert mine[0] == ref[0] and mine[4] == ref[4]
            for j in (1, 2, 3):
                rv = uh(ref[j])
                assert abs(mine[j] - rv) <= (1e-12 if kind == "v2" else 0.0) * abs(rv), (tab["env"], mine, ref)


def test_scalar_port_equals_vectorised_oracle():
    rng = np.random.default_rng(7)
    for kind in fo.KINDS:
        z = rng.standard_normal(300)
        it = iter(z)
        env = fo.ScalarPortEnv(kind, r=0.4, K=1.3, sigma=0.05, C=0.3, n_actions=50, init_state=0.9, Tmax=40,
                               normal=lambda m=0.0, s=1.0: float(m + s * next(it)))
        s = env.reset()[0]
        years, K, r = 0, env.K, env.r
        for t in range(120):
            a = int(rng.integers(0, 12)) if kind == "v0" else np.array([rng.uniform(-1.1, -0.7)])
            obs, rew, done, _ = env.step(a)
            zz = z[(4 if kind == "v4" else 0) + t]
            s, r2, d2, years = fo.step_vec(kind, s, a if kind == "v0" else a[0], zz, years, r=r, K=K, sigma=0.05,
                                           C=0.3, n_actions=50, Tmax=40)
            assert close(obs[0], s, kind) and close(rew, r2, kind) and bool(d2) == done


# ------------------------------------------------------------------------------------- next row f1: fishing-v5..v11
def _ln_params(d, m, kind):
    p = {k: float(d["p_" + k][m]) for k in ("r", "K", "C", "M", "theta", "q", "b", "a", "alpha", "sigma")}
    return p


@pytest.mark.parametrize("kind", ["v5", "v6", "v7", "v8", "v9", "v10", "v11"])
def test_lognormal_family_trajectories_match_reference(kind):
    d = np.load(os.path.join(G, f"traj_{kind}.npz"))
    T, M = d["actions"].shape
    for m in range(M):
        p = _ln_params(d, m, kind)
        model = fo.MODEL_NAMES[int(d["p_model"][m])] if kind == "v11" else fo.MODEL_OF_KIND[kind]
        K_state = 1.0 if kind == "v11" else p["K"]
        s, years = d["obs0"][m], 0
        assert s == d["p_x0"][m] / K_state - 1
        for t in range(T):
            if kind == "v10":
                p["r"] = p["r"] + p["alpha"]          # growth_models.py:148, before the draw, never reset
            s, rew, done, years = fo.step_ln(model, s, d["actions"][t, m], d["noise"][t, m], years, p, K_state=K_state,
                                             Tmax=int(d["p_Tmax"][m]))
            assert np.isclose(s, d["obs"][t, m], rtol=1e-13, atol=1e-15, equal_nan=True), (kind, m, t, float(s), d["obs"][t, m])
            assert np.isclose(rew, d["reward"][t, m], rtol=1e-13, atol=1e-15, equal_nan=True)
            assert bool(done) == bool(d["done"][t, m]), (kind, m, t)
    assert d["done"].any() and not d["done"].all()
