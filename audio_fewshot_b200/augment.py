"""Spectrogram-domain augmentation behind the reference's API
(libfewshot_core/audio_augmentations.py: augment_spectrogram :531-604, batch_augment_spectrogram
:607-648, the eight augmentations :56-528, denormalize/normalize :16-53).

Same function names, arguments, defaults and -- because the random parameters are drawn here with
Python's `random` in exactly the reference's order -- the same results for the same `random.seed`.
The arithmetic is one sm_100a kernel per call (csrc/specaug.cu): every [H, W] plane is de-normalised,
augmented and re-normalised in shared memory, quantiles by exact radix select instead of a sort.
CUDA tensors only; a CPU tensor raises (no fallback).
"""
import ctypes as C
import random

import numpy as np
import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream

TYPES = {"cutout": 0, "linear_filter": 1, "noise_suppression": 2, "noise_matching": 3,
         "background_subtraction": 4, "contrast_enhancement": 5, "foreground_norm": 6, "wiener_filter": 7}
RANDOM_CHOICES = ["cutout", "linear_filter", "noise_suppression", "noise_matching", "background_subtraction",
                  "contrast_enhancement", "foreground_norm", "wiener_filter"]  # :552-555, order matters


def _scalar(v):
    return float(v.item()) if isinstance(v, torch.Tensor) else float(v)


def _planes(spec):
    if spec.dim() not in (2, 3, 4):
        raise ValueError("Expected 2D, 3D or 4D tensor, got shape %s" % (tuple(spec.shape),))
    h, w = spec.shape[-2], spec.shape[-1]
    return int(spec.numel() // (h * w)), int(h), int(w)


def _launch(spec, mean, std, type_name, rects=(), fill=0.0, p0=0.0, p1=0.0, i0=0, curve=None):
    _need_cuda(spec, "spectrogram")
    spec = spec.contiguous()
    planes, h, w = _planes(spec)
    cfg = _lib.SpecAugCfg()
    cfg.type = TYPES[type_name]
    cfg.n_rect = len(rects)
    for k, r in enumerate(rects):
        for j in range(4):
            cfg.rect[k][j] = int(r[j])
    cfg.fill, cfg.p0, cfg.p1, cfg.i0 = float(fill), float(p0), float(p1), int(i0)
    out = torch.empty_like(spec)
    _lib.check(_lib.lib().afs_spec_augment(_ptr(spec), planes, h, w, _scalar(mean), _scalar(std), C.byref(cfg),
                                           _ptr(curve), _ptr(out), _stream()), "afs_spec_augment")
    return out


# ----------------------------------------------------------------- parameter draws (reference order)
def _draw_cutouts(h, w, num_cutouts, cutout_size_ratio):
    rects = []
    for _ in range(num_cutouts):  # :86-95
        ch = int(h * random.uniform(*cutout_size_ratio))
        cw = int(w * random.uniform(*cutout_size_ratio))
        top = random.randint(0, max(0, h - ch))
        left = random.randint(0, max(0, w - cw))
        rects.append((top, left, ch, cw))
    return rects


def _draw_filter_curve(h, num_points, filter_strength):
    freq_points = sorted(random.sample(range(h), min(num_points, h)))  # :498
    values = [1.0 + random.uniform(-filter_strength, filter_strength) for _ in freq_points]  # :502
    return np.interp(np.arange(h), freq_points, values)  # :505-509 (float64, cast to the tensor dtype :512)


# ----------------------------------------------------------------- the reference's public functions
def denormalize_spectrogram(spectrogram, mean, std):
    return spectrogram * std + mean


def normalize_spectrogram(spectrogram, mean, std):
    return (spectrogram - mean) / std


def random_cutout(spectrogram, num_cutouts=1, cutout_size_ratio=(0.1, 0.3), fill_value=0.0, mean=0.0, std=1.0):
    _, h, w = _planes(spectrogram)
    rects = _draw_cutouts(h, w, num_cutouts, cutout_size_ratio)
    if len(rects) > 8:
        raise ValueError("at most 8 cutouts per call")
    return _launch(spectrogram, mean, std, "cutout", rects=rects, fill=fill_value)


def apply_linear_filteraugment(spectrogram, num_points=4, filter_strength=0.5, mean=0.0, std=1.0):
    _, h, _ = _planes(spectrogram)
    curve = torch.tensor(_draw_filter_curve(h, num_points, filter_strength), device=spectrogram.device,
                         dtype=torch.float32)
    return _launch(spectrogram, mean, std, "linear_filter", curve=curve)


def background_noise_suppression(spectrogram, noise_percentile=20, suppression_strength=0.5, mean=0.0, std=1.0):
    return _launch(spectrogram, mean, std, "noise_suppression", p0=noise_percentile / 100.0, p1=suppression_strength)


def adaptive_noise_profile_matching(spectrogram, target_noise_level=None, smoothing_window=5, mean=0.0, std=1.0):
    if target_noise_level is None:
        target_noise_level = 0.1  # :416-417
    return _launch(spectrogram, mean, std, "noise_matching", p0=target_noise_level, i0=smoothing_window)


def temporal_median_background_subtraction(spectrogram, percentile=10, mean=0.0, std=1.0):
    return _launch(spectrogram, mean, std, "background_subtraction", p0=percentile / 100.0)


def spectral_contrast_enhancement(spectrogram, contrast_factor=1.5, clip_percentile=95, mean=0.0, std=1.0):
    p1 = clip_percentile / 100.0 if clip_percentile < 100 else 2.0
    return _launch(spectrogram, mean, std, "contrast_enhancement", p0=contrast_factor, p1=p1)


def foreground_energy_normalization(spectrogram, top_k_percent=20, mean=0.0, std=1.0):
    return _launch(spectrogram, mean, std, "foreground_norm", p0=1.0 - top_k_percent / 100.0)


def wiener_like_filtering(spectrogram, noise_floor_percentile=15, gain_factor=2.0, mean=0.0, std=1.0):
    return _launch(spectrogram, mean, std, "wiener_filter", p0=noise_floor_percentile / 100.0, p1=gain_factor)


def augment_spectrogram(spectrogram, mean, std, augmentation_type="random", **kwargs):
    """De-normalise, apply one augmentation, re-normalise (:531-604) -- fused into one kernel launch."""
    if augmentation_type == "random":
        augmentation_type = random.choice(RANDOM_CHOICES)
    if augmentation_type == "cutout":
        n = kwargs.get("num_cutouts", random.randint(1, 3))
        ratio = kwargs.get("cutout_size_ratio", (0.1, 0.3))
        return random_cutout(spectrogram, n, ratio, kwargs.get("fill_value", 0.0), mean, std)
    if augmentation_type == "linear_filter":
        n = kwargs.get("num_points", random.randint(3, 6))
        strength = kwargs.get("filter_strength", random.uniform(0.3, 0.7))
        return apply_linear_filteraugment(spectrogram, n, strength, mean, std)
    if augmentation_type == "noise_suppression":
        pct = kwargs.get("noise_percentile", random.uniform(15, 25))
        strength = kwargs.get("suppression_strength", random.uniform(0.4, 0.7))
        return background_noise_suppression(spectrogram, pct, strength, mean, std)
    if augmentation_type == "noise_matching":
        target = kwargs.get("target_noise_level", None)
        win = kwargs.get("smoothing_window", random.choice([3, 5, 7]))
        return adaptive_noise_profile_matching(spectrogram, target, win, mean, std)
    if augmentation_type == "background_subtraction":
        return temporal_median_background_subtraction(spectrogram, kwargs.get("percentile", random.uniform(5, 15)),
                                                      mean, std)
    if augmentation_type == "contrast_enhancement":
        factor = kwargs.get("contrast_factor", random.uniform(1.3, 2.0))
        clip = kwargs.get("clip_percentile", random.uniform(90, 98))
        return spectral_contrast_enhancement(spectrogram, factor, clip, mean, std)
    if augmentation_type == "foreground_norm":
        return foreground_energy_normalization(spectrogram, kwargs.get("top_k_percent", random.uniform(15, 25)),
                                               mean, std)
    if augmentation_type == "wiener_filter":
        pct = kwargs.get("noise_floor_percentile", random.uniform(10, 20))
        gain = kwargs.get("gain_factor", random.uniform(1.5, 2.5))
        return wiener_like_filtering(spectrogram, pct, gain, mean, std)
    raise ValueError("Unknown augmentation_type: %s" % augmentation_type)


def batch_augment_spectrogram(spectrograms, mean, std, num_augmentations=10, **kwargs):
    """[B, C, H, W] (or [C, H, W]) -> [B * num_augmentations, C, H, W] (:607-648)."""
    if spectrograms.dim() == 3:
        spectrograms = spectrograms.unsqueeze(0)
    elif spectrograms.dim() != 4:
        raise ValueError("Expected 3D or 4D tensor, got shape %s" % (tuple(spectrograms.shape),))
    out = []
    for i in range(spectrograms.shape[0]):
        spec = spectrograms[i:i + 1]
        for _ in range(num_augmentations):
            out.append(augment_spectrogram(spec, mean, std, **kwargs).squeeze(0))
    return torch.stack(out, dim=0)
