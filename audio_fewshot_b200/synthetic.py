"""Synthetic inputs and checkpoint-free weights for benchmarks, smoke tests and examples (no dataset
and no checkpoint can be downloaded; throughput is measured on synthetic episodes of the repo's audio shape).

Both generators are keyed so that every rank of every world size produces the same content for the same global
episode index / parameter name (numpy PCG64 streams, identical on every machine)."""
import zlib

import numpy as np


def synthetic_clip_batch(seed, first_episode, n_episodes, W, S, Q, L, sample_rate=16000):
    """[E*W*(S+Q), L] fp32, class-major rows: N(0,1)*0.1 noise + a class-dependent tone
    0.05*sin(2 pi f_c t), f_c = 200*(c+1) Hz.  Content depends only on the global episode index."""
    t = np.arange(L, dtype=np.float64) / sample_rate
    out = np.empty((n_episodes, W, S + Q, L), dtype=np.float32)
    for e in range(n_episodes):
        r = np.random.default_rng((seed, first_episode + e))
        noise = r.standard_normal((W, S + Q, L)).astype(np.float32) * np.float32(0.1)
        for c in range(W):
            tone = (0.05 * np.sin(2.0 * np.pi * 200.0 * (c + 1) * t)).astype(np.float32)
            out[e, c] = noise[c] + tone[None, :]
    return out.reshape(n_episodes * W * (S + Q), L)


def synthetic_clip_batch_device(seed, first_episode, n_episodes, W, S, Q, L, device, sample_rate=16000):
    """The same kind of batch generated ON the device (torch Philox streams keyed by the global episode index), for
    runs whose episode count makes host generation the bottleneck (10 000-episode evaluation).  Not bit-identical to
    synthetic_clip_batch; identical across world sizes for the same global episode index."""
    import torch

    t = torch.arange(L, dtype=torch.float64, device=device) / sample_rate
    tones = torch.stack([(0.05 * torch.sin(2.0 * np.pi * 200.0 * (c + 1) * t)).float() for c in range(W)])  # [W, L]
    out = torch.empty((n_episodes, W, S + Q, L), dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    for e in range(n_episodes):
        gen.manual_seed((int(seed) << 32) + int(first_episode) + e)
        out[e].normal_(0.0, 0.1, generator=gen)
    out += tones[None, :, None, :]
    return out.reshape(n_episodes * W * (S + Q), L)


def name_seeded_weights_(net):
    """Overwrite EVERY parameter and buffer with values derived from its name and shape, so that two modules with
    the same state_dict layout (this package's and the reference's) get identical weights and non-trivial BatchNorm
    statistics without shipping a checkpoint."""
    import torch

    sd = net.state_dict()
    for key in sorted(sd.keys()):
        t = sd[key]
        if key.endswith("num_batches_tracked"):
            continue
        r = np.random.default_rng(zlib.crc32(key.encode()))
        shape = tuple(t.shape)
        if key.endswith("running_var") or (key.endswith("weight") and t.dim() == 1):
            v = r.uniform(0.5, 1.5, size=shape)
        elif key.endswith("running_mean") or key.endswith("bias"):
            v = r.standard_normal(shape) * 0.1
        elif key.endswith("temperature"):
            v = np.full(shape, np.log(1.0 / 200.0))
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            v = r.standard_normal(shape) * np.sqrt(2.0 / max(fan_in, 1))
        with torch.no_grad():
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
    return net
