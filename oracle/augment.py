"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (torch, one plane at a time) of the reference's spectrogram-domain augmentations,
libfewshot_core/audio_augmentations.py.  Random parameters are drawn with Python's `random` in the
reference's order, so `random.seed(k)` before a call reproduces the reference's choice.  Pinned by
tests/golden/augment.npz (outputs of the real reference, oracle/make_golden.py --augment)."""
import random

import numpy as np
import torch

CHOICES = ["cutout", "linear_filter", "noise_suppression", "noise_matching", "background_subtraction",
           "contrast_enhancement", "foreground_norm", "wiener_filter"]  # audio_augmentations.py:552-555


def _per_plane(spec, fn):
    """Apply fn to every [H, W] plane of a 2/3/4-D tensor (the reference's mode == '2D'/'3D'/'4D' loops)."""
    h, w = spec.shape[-2:]
    flat = spec.reshape(-1, h, w).clone()
    for i in range(flat.shape[0]):
        flat[i] = fn(flat[i])
    return flat.reshape(spec.shape)


def cutout(spec, num_cutouts, ratio, fill):  # :56-103
    out = spec.clone()
    h, w = spec.shape[-2:]
    for _ in range(num_cutouts):
        ch = int(h * random.uniform(*ratio))
        cw = int(w * random.uniform(*ratio))
        top = random.randint(0, max(0, h - ch))
        left = random.randint(0, max(0, w - cw))
        out[..., top:top + ch, left:left + cw] = fill
    return out


def linear_filter(spec, num_points, strength):  # :467-528
    h = spec.shape[-2]
    pts = sorted(random.sample(range(h), min(num_points, h)))
    vals = [1.0 + random.uniform(-strength, strength) for _ in pts]
    curve = torch.tensor(np.interp(np.arange(h), pts, vals), dtype=spec.dtype)
    return spec * curve.view(h, 1)


def _soft_mask(p, thr):
    return torch.sigmoid((p.abs() - thr) / (thr * 0.1 + 1e-8))


def noise_suppression(spec, percentile, strength):  # :106-158
    def f(p):
        thr = torch.quantile(p.abs(), percentile / 100.0)
        return p * (1 - strength * (1 - _soft_mask(p, thr)))
    return _per_plane(spec, f)


def noise_matching(spec, target, window):  # :388-464
    target = 0.1 if target is None else target
    w = spec.shape[-1]

    def f(p):
        est = p.abs().min(dim=0, keepdim=True)[0]
        if window > 1 and w > window:
            k = torch.ones(1, 1, window) / window
            est = torch.nn.functional.pad(est.unsqueeze(0), (window // 2, window // 2), mode="reflect")
            est = torch.nn.functional.conv1d(est, k).squeeze(0)
        cur = est.mean()
        scale = torch.clamp(target / (cur + 1e-8), 0.5, 2.0) if cur > 1e-8 else 1.0
        m = _soft_mask(p, torch.quantile(p.abs(), 0.3))
        return p * (m + (1 - m) * scale)
    return _per_plane(spec, f)


def background_subtraction(spec, percentile):  # :161-209
    return _per_plane(spec, lambda p: torch.clamp(p - torch.quantile(p, percentile / 100.0, dim=1, keepdim=True),
                                                  min=0.0))


def contrast(spec, factor, clip_percentile):  # :212-266
    def f(p):
        m = p.mean()
        p = m + (p - m) * factor
        if clip_percentile < 100:
            mx = torch.quantile(p.abs(), clip_percentile / 100.0)
            p = torch.clamp(p, -mx, mx)
        return p
    return _per_plane(spec, f)


def foreground_norm(spec, top_k_percent):  # :269-325
    def f(p):
        e = p.abs()
        fg = e >= torch.quantile(e, 1.0 - top_k_percent / 100.0)
        if fg.sum() > 0:
            v = p[fg]
            p = (p - v.mean()) / (v.std() + 1e-8)
        return p
    return _per_plane(spec, f)


def wiener(spec, percentile, gain):  # :328-385
    def f(p):
        snr = p.abs() / (torch.quantile(p.abs(), percentile / 100.0) + 1e-8)
        return p * (snr / (snr + 1.0) * gain)
    return _per_plane(spec, f)


def augment_spectrogram(spec, mean, std, augmentation_type="random", **kw):  # :531-604
    x = spec * torch.tensor(std, dtype=spec.dtype) + torch.tensor(mean, dtype=spec.dtype)
    t = random.choice(CHOICES) if augmentation_type == "random" else augmentation_type
    if t == "cutout":
        x = cutout(x, kw.get("num_cutouts", random.randint(1, 3)), kw.get("cutout_size_ratio", (0.1, 0.3)),
                   kw.get("fill_value", 0.0))
    elif t == "linear_filter":
        n = kw.get("num_points", random.randint(3, 6))
        x = linear_filter(x, n, kw.get("filter_strength", random.uniform(0.3, 0.7)))
    elif t == "noise_suppression":
        pct = kw.get("noise_percentile", random.uniform(15, 25))
        x = noise_suppression(x, pct, kw.get("suppression_strength", random.uniform(0.4, 0.7)))
    elif t == "noise_matching":
        x = noise_matching(x, kw.get("target_noise_level", None), kw.get("smoothing_window", random.choice([3, 5, 7])))
    elif t == "background_subtraction":
        x = background_subtraction(x, kw.get("percentile", random.uniform(5, 15)))
    elif t == "contrast_enhancement":
        f = kw.get("contrast_factor", random.uniform(1.3, 2.0))
        x = contrast(x, f, kw.get("clip_percentile", random.uniform(90, 98)))
    elif t == "foreground_norm":
        x = foreground_norm(x, kw.get("top_k_percent", random.uniform(15, 25)))
    elif t == "wiener_filter":
        pct = kw.get("noise_floor_percentile", random.uniform(10, 20))
        x = wiener(x, pct, kw.get("gain_factor", random.uniform(1.5, 2.5)))
    else:
        raise ValueError(t)
    return (x - torch.tensor(mean, dtype=spec.dtype)) / torch.tensor(std, dtype=spec.dtype)
