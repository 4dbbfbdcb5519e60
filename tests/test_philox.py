"""Philox4x32-10 restatement pinned to the Random123 known-answer vectors (kat_vectors)."""
import numpy as np

from oracle import philox

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_uniform_is_open_interval_and_exact():
    r = np.array([0, 1, 511, 512, 0xFFFFFFFF], dtype=np.uint32)
    u = philox.u01(r)
    assert u.dtype == np.float32
    assert (u > 0).all() and (u < 1).all()
    assert u[0] == np.float32(0.5 * 2.0 ** -23) and u[-1] == np.float32(1.0 - 2.0 ** -24)
    assert u[2] == u[0] and u[3] == np.float32(1.5 * 2.0 ** -23)


def test_noise_statistics_and_keying():
    z = philox.clip_noise(seed=7, clip_index=3, length=200001)
    assert z.shape == (200001,)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert not np.array_equal(z[:100], philox.clip_noise(7, 4, 100))
    assert np.array_equal(z[:100], philox.clip_noise(7, 3, 100))  # counter-based: prefix-stable


def test_augment_shift_and_gain():
    x = np.arange(1, 21, dtype=np.float32)[None, :]
    y = philox.augment_waveform(x, seed=1, first_clip_index=0, gain_db=(6.0, 6.0), max_shift=0)
    np.testing.assert_allclose(y, x * 10 ** (6.0 / 20), rtol=1e-6)
    ks = set()
    for c in range(40):
        _, k, _ = philox.clip_params(1, c, 0.0, 0.0, 3, 0.0, 0.0)
        ks.add(k)
        y = philox.augment_waveform(x, 1, c, max_shift=3)[0]
        want = np.zeros(20, np.float32)
        if k >= 0:
            want[k:] = x[0, : 20 - k]
        else:
            want[: 20 + k] = x[0, -k:]
        np.testing.assert_array_equal(y, want)
    assert ks == set(range(-3, 4))
