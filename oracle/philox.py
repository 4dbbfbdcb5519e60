"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatement of Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random
numbers: as easy as 1, 2, 3", SC'11; Random123 `philox.h`) and of the uniform /
Box-Muller draws used by the waveform augmentation spec (the
reference has no waveform augmentation, so this file IS the spec).  The integer
stream must be bit-exact with audio_fewshot_b200/csrc/philox.cuh; it is pinned
to the Random123 known-answer vectors in tests/test_philox.py.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

STREAM_PARAMS = 0
STREAM_NOISE = 1


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01(r):
    """uint32 -> float32 in (0, 1): ((r >> 9) + 0.5) * 2^-23 (every step exact in fp32)."""
    r = np.asarray(r, dtype=np.uint32)
    return (((r >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)).astype(np.float32)


def clip_params(seed, clip_index, gain_db_lo, gain_db_hi, max_shift, noise_std_lo, noise_std_hi):
    """Per-clip draws: (gain g, shift k, noise sigma).  Counter = (0, STREAM_PARAMS, clip_lo, clip_hi)."""
    seed = int(seed)
    clip_index = int(clip_index)
    r0, r1, r2, _ = philox4x32_10(0, STREAM_PARAMS, clip_index & 0xFFFFFFFF, (clip_index >> 32) & 0xFFFFFFFF,
                                  seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    f = np.float32
    gain_db = f(gain_db_lo) + (f(gain_db_hi) - f(gain_db_lo)) * u01(r0)
    g = f(10.0) ** (f(gain_db) * f(0.05))
    span = 2 * int(max_shift) + 1
    draw = int(np.floor(u01(r1) * f(span)))
    draw = min(draw, span - 1)
    k = draw - int(max_shift)
    sigma = f(noise_std_lo) + (f(noise_std_hi) - f(noise_std_lo)) * u01(r2)
    return f(g), k, f(sigma)


def clip_noise(seed, clip_index, length):
    """Standard-normal noise n[0:length] of a clip: n[2i], n[2i+1] = Box-Muller (cos, sin) of the
    first two words of Philox(counter = (i, STREAM_NOISE, clip_lo, clip_hi), key = seed)."""
    seed = int(seed)
    clip_index = int(clip_index)
    nblk = (int(length) + 1) // 2
    i = np.arange(nblk, dtype=np.uint32)
    r0, r1, _, _ = philox4x32_10(i, STREAM_NOISE, clip_index & 0xFFFFFFFF, (clip_index >> 32) & 0xFFFFFFFF,
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = u01(r0).astype(np.float64)
    u2 = u01(r1).astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * nblk, dtype=np.float64)
    z[0::2] = rad * np.cos(2.0 * np.pi * u2)
    z[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return z[:length].astype(np.float32)


def augment_waveform(x, seed, first_clip_index, gain_db=(0.0, 0.0), max_shift=0, noise_std=(0.0, 0.0)):
    """y[b, n] = g_b * x[b, n - k_b] (zero fill) + sigma_b * noise_b[n].   x: [B, L] float32."""
    x = np.asarray(x, dtype=np.float32)
    B, L = x.shape
    y = np.zeros_like(x)
    for b in range(B):
        g, k, sigma = clip_params(seed, first_clip_index + b, gain_db[0], gain_db[1], max_shift,
                                  noise_std[0], noise_std[1])
        shifted = np.zeros(L, dtype=np.float32)
        if k >= 0:
            shifted[k:] = x[b, : L - k] if k > 0 else x[b]
        else:
            shifted[: L + k] = x[b, -k:]
        v = (g * shifted).astype(np.float32)
        if sigma > 0:
            v = (v + (sigma * clip_noise(seed, first_clip_index + b, L)).astype(np.float32)).astype(np.float32)
        y[b] = v
    return y
