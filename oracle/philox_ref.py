"""numpy restatement of the engine's noise spec (gym_fishing_b200/csrc/philox.cuh).  TEST INFRASTRUCTURE ONLY.

This is NOT part of the reference (which draws from numpy's global MT19937); it restates the *new* counter-based
generator so that tests can check the CUDA implementation word for word: Philox4x32-10 (Salmon, Moraes, Dror, Shaw,
"Parallel random numbers: as easy as 1, 2, 3", SC'11; known-answer vectors from the Random123 distribution's
kat_vectors file are checked in tests/test_philox_ref.py), 52-bit open-interval uniforms and Box-Muller.

This is the MATHEMATICAL definition of the draws (numpy's log / sqrt / cos / sin).  The device evaluates the same formulas
with its own table-assisted routines (gym_fishing_b200/csrc/normal.cuh, round 2; round 1: fdlibm-style polynomials), so
device draws agree with this file to a few ulps, not bit for bit (tests/test_gpu_parity.py:
test_philox_normals_against_long_double_evaluation bounds the difference by 3e-15); parity tests of the dynamics therefore
export the engine's own draws (fishing_noise_fill) and inject them here, never the other way round."""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint32).copy()
    k1 = np.asarray(k1, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def _open_unit(lo, hi):
    k = ((hi.astype(np.uint64) & np.uint64(0xFFFFF)) << np.uint64(32)) | lo.astype(np.uint64)
    return (k.astype(np.float64) + 0.5) * 2.0 ** -52


def uniforms(seed, ids, t, stream):
    ids = np.asarray(ids, dtype=np.uint64)
    t = np.asarray(t, dtype=np.uint64)
    seed = np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)
    c3 = (np.uint64(stream) << np.uint64(24)) | ((t >> np.uint64(32)) & np.uint64(0xFFFFFF))
    w = philox4x32_10((ids & MASK32).astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32),
                      (t & MASK32).astype(np.uint32), c3.astype(np.uint32),
                      np.uint32(seed & MASK32), np.uint32(seed >> np.uint64(32)))
    return _open_unit(w[0], w[1]), _open_unit(w[2], w[3])


def normal_pair(seed, ids, t, stream):
    u1, u2 = uniforms(seed, ids, t, stream)
    R = np.sqrt(-2.0 * np.log(u1))
    return R * np.cos(2.0 * np.pi * u2), R * np.sin(2.0 * np.pi * u2)


def step_noise(seed, env_ids, t):
    """z(env, t) of stream STEP: pair id = env >> 1, even env takes z0, odd env z1."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    z0, z1 = normal_pair(seed, env_ids >> np.uint64(1), t, 0)
    return np.where((env_ids & np.uint64(1)) == 1, z1, z0)


def reset_noise(seed, env_ids, t, stream):
    """(zK, zr) of the v4 streams (1 = auto-reset at step t, 2 = explicit reset number t)."""
    return normal_pair(seed, np.asarray(env_ids, dtype=np.uint64), t, stream)
