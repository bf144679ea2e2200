"""Guards the assumption behind the CUDA BMSY kernel's float32 exp: numpy's float32 np.exp on this host is the
simd_exp_FLOAT algorithm restated in oracle/np_expf_ref.py (and in csrc/step_kernels.cu: np_expf).  CPU only."""
import numpy as np

from oracle.np_expf_ref import np_expf


def test_restatement_equals_numpy_float32_exp():
    rng = np.random.default_rng(0)
    for lo, hi in [(-0.2, 0.2), (-5, 5), (-87, 88)]:
        x = rng.uniform(lo, hi, 500_000).astype(np.float32)
        assert np.array_equal(np.exp(x), np_expf(x))
    edge = np.array([0.0, -0.0, 1.0, -1.0, 88.7, 88.8, -103.9, -104.5, np.inf, -np.inf, np.nan], dtype=np.float32)
    with np.errstate(all="ignore"):
        assert np.array_equal(np.exp(edge), np_expf(edge), equal_nan=True)


def test_float32_exp_is_not_correctly_rounded_which_is_why_bmsy_needs_it():
    x = np.linspace(-0.1, 0.1, 10001, dtype=np.float32)
    cr = np.exp(x.astype(np.float64)).astype(np.float32)
    assert (np.exp(x) != cr).mean() > 0.05
