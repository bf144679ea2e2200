"""numpy restatement of numpy's own float32 `exp` kernel (simd_exp_FLOAT, shared by the AVX2+FMA3 and AVX512F builds of
numpy >= 1.20; numpy is an un-vendored dependency of the reference, `examples/requirements.txt:14` pins 1.19.2, this
container has 2.3.5).  TEST INFRASTRUCTURE ONLY.

Why it exists: BMSY for fishing-v2 (gym_fishing/models/policies.py:51-67) is an argmax over float32 values that tie to
the last bit, so its result depends on the exact float32 `exp` used.  The CUDA BMSY kernel
(gym_fishing_b200/csrc/step_kernels.cu: np_expf) reproduces this algorithm operation for operation;
tests/test_np_expf.py checks THIS restatement against np.exp on the running host, which guards the assumption."""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    # float32 FMA through float64: the product of two float32 is exact in float64
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def np_expf(x):
    x = np.asarray(x, dtype=f32)
    magic = f32(12582912.0)
    q = (x * f32(1.442695040888963407359924681001892137)).astype(f32)
    q = ((q + magic).astype(f32) - magic).astype(f32)
    r = _fma(q, f32(-6.93145752e-1), x)
    r = _fma(q, f32(-1.42860677e-6), r)
    num = _fma(f32(5.082762527590693718096e-04), r, f32(6.757896990527504603057e-03))
    for c in (5.114512081637298353406e-02, 2.473615434895520810817e-01, 7.257664613233124478488e-01,
              9.999999999980870924916e-01):
        num = _fma(num, r, f32(c))
    den = _fma(f32(2.159509375685829852307e-02), r, f32(-2.742335390411667452936e-01))
    den = _fma(den, r, f32(1.0))
    with np.errstate(all="ignore"):
        out = np.ldexp((num / den).astype(f32), q.astype(np.int32)).astype(f32)
    out = np.where(x > f32(88.72283935546875), f32(np.inf), out)
    out = np.where(x < f32(-103.97208404541015625), f32(0), out)
    return np.where(np.isnan(x), x, out).astype(f32)
