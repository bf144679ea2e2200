"""The numpy restatement of the noise generator against Random123's published known-answer vectors, and basic
distribution sanity.  CPU only."""
import numpy as np

from oracle import philox_ref as pr


def test_philox4x32_10_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, exp in kat:
        got = pr.philox4x32_10(*[np.uint32(c) for c in ctr], np.uint32(key[0]), np.uint32(key[1]))
        assert tuple(int(g) for g in got) == exp


def test_uniforms_open_interval_and_normals_moments():
    ids = np.arange(200000, dtype=np.uint64)
    u1, u2 = pr.uniforms(1234, ids, 7, 0)
    assert u1.min() > 0 and u1.max() < 1 and u2.min() > 0 and u2.max() < 1
    z0, z1 = pr.normal_pair(1234, ids, 7, 0)
    z = np.concatenate([z0, z1])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3) < 0.06
    assert abs(np.corrcoef(z0, z1)[0, 1]) < 0.01


def test_streams_and_steps_are_distinct():
    ids = np.arange(64, dtype=np.uint64)
    a = pr.step_noise(5, ids, 0)
    assert not np.array_equal(a, pr.step_noise(5, ids, 1))
    assert not np.array_equal(a, pr.step_noise(6, ids, 0))
    assert np.array_equal(a, pr.step_noise(5, ids, 0))
    # partition invariance: any sub-range gives the same values
    assert np.array_equal(a[10:20], pr.step_noise(5, ids[10:20], 0))
