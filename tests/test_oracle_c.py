"""Pins the C restatement (oracle/fishing_oracle.c, used only as bench.py's fast CPU baseline) to the numpy oracle.  CPU only."""
import numpy as np
import pytest

from oracle import fishing_oracle as fo
from oracle.build_c import load


@pytest.mark.parametrize("kind,code", [("v0", 0), ("v1", 1), ("v2", 2), ("v4", 4)])
def test_c_oracle_equals_numpy_oracle(kind, code):
    lib = load()
    rng = np.random.default_rng(code)
    n, T = 5000, 30
    K = rng.uniform(0.8, 1.2, n) if kind == "v4" else np.full(n, 1.7)
    r = rng.uniform(0.2, 0.4, n) if kind == "v4" else np.full(n, 0.4)
    s = np.full(n, 0.9 / 1.7 - 1) if kind != "v4" else np.full(n, 0.75)
    sc = s.copy()
    years = np.zeros(n, dtype=np.int64)
    yc = np.zeros(n, dtype=np.int32)
    rew, done = np.zeros(n), np.zeros(n, dtype=np.uint8)
    for t in range(T):
        a = rng.integers(0, 20, n).astype(np.float64) if kind == "v0" else rng.uniform(-1.2, -0.5, n)
        z = rng.standard_normal(n)
        s, r_np, d_np, years = fo.step_vec(kind, s, a, z, years, r=r, K=K, sigma=0.05, C=0.4, n_actions=50, Tmax=12)
        lib.oracle_step(code, n, r.ctypes.data, K.ctypes.data, 0.4, 1.7, 0.05, 0.4, 50.0, 12, 0, 0.0, a.ctypes.data, z.ctypes.data,
                        sc.ctypes.data, yc.ctypes.data, rew.ctypes.data, done.ctypes.data)
        if kind == "v2":
            assert np.allclose(sc, s, rtol=1e-12, atol=1e-12) and np.allclose(rew, r_np, rtol=1e-12, atol=1e-12)
        else:
            assert np.array_equal(sc, s) and np.array_equal(rew, r_np)
        assert np.array_equal(done.astype(bool), d_np)
