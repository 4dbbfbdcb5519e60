"""TEST INFRASTRUCTURE ONLY.  Seeded parity cases shared by oracle/make_golden.py (which runs the
real reference on them) and tests/ (which run the oracle and the CUDA kernels on them).
Inputs come from numpy PCG64 streams, so they are identical on every machine."""
import zlib

import numpy as np


def _rng(seed):
    return np.random.default_rng(seed)


# ------------------------------------------------------------------ prototype heads
PROTO_CASES = {
    # C1 ProtoNet/Conv64F, C2 ProtoNet/ResNet-12 1-shot, C4 DeepBDC vectors
    "c1_euclid": dict(seed=11, E=1, W=5, S=5, Q=15, D=1600, head="proto", mode="euclidean"),
    "c2_euclid_d12800": dict(seed=12, E=2, W=5, S=1, Q=15, D=12800, head="proto", mode="euclidean"),
    "c4_euclid_d2080": dict(seed=13, E=4, W=5, S=5, Q=10, D=2080, head="proto", mode="euclidean"),
    "cos_sim": dict(seed=14, E=2, W=5, S=5, Q=15, D=1600, head="proto", mode="cos_sim"),
    "odd_way7": dict(seed=15, E=3, W=7, S=3, Q=4, D=64, head="proto", mode="euclidean"),
    "way20": dict(seed=16, E=2, W=20, S=2, Q=3, D=128, head="proto", mode="euclidean"),
    "bdc_1shot_dot": dict(seed=17, E=2, W=5, S=1, Q=10, D=2080, head="deepbdc", mode="dot"),
    "bdc_5shot": dict(seed=18, E=2, W=5, S=5, Q=10, D=2080, head="deepbdc", mode="euclidean"),
}


def proto_features(c):
    n = c["E"] * c["W"] * (c["S"] + c["Q"])
    return _rng(c["seed"]).standard_normal((n, c["D"])).astype(np.float32)


# ------------------------------------------------------------------ DN4
DN4_CASES = {
    "c3_5shot": dict(seed=21, E=1, W=5, S=5, Q=15, C=64, H=4, Wd=5, n_k=3),
    "c3_1shot_q10": dict(seed=22, E=1, W=5, S=1, Q=10, C=64, H=4, Wd=5, n_k=3),
    "nk1_two_episodes": dict(seed=23, E=2, W=5, S=5, Q=10, C=64, H=4, Wd=5, n_k=1),
    "nk5_odd": dict(seed=24, E=2, W=3, S=2, Q=3, C=40, H=3, Wd=7, n_k=5),
    "resnet12_map": dict(seed=25, E=1, W=5, S=5, Q=2, C=640, H=8, Wd=9, n_k=3),
}


def dn4_features(c):
    n = c["E"] * c["W"] * (c["S"] + c["Q"])
    # non-negative like post-ReLU maps, with a few exact zeros
    x = _rng(c["seed"]).standard_normal((n, c["C"], c["H"], c["Wd"])).astype(np.float32)
    return np.maximum(x, 0.0) + 0.01 * np.abs(x)


# ------------------------------------------------------------------ BDC
BDC_CASES = {
    "c4_map": dict(seed=31, B=4, C=64, H=16, Wd=19, log_temp=float(np.log(1.0 / 200.0))),
    "small_map": dict(seed=32, B=3, C=64, H=5, Wd=5, log_temp=-3.0),
    "dim48": dict(seed=33, B=2, C=48, H=7, Wd=9, log_temp=-4.0),
}


def bdc_features(c):
    x = _rng(c["seed"]).standard_normal((c["B"], c["C"], c["H"], c["Wd"])).astype(np.float32)
    return np.maximum(x, 0.0)


# ------------------------------------------------------------------ ragged split / vote
SPLIT_CASES = {
    "ragged_small": dict(seed=41, E=1, W=2, S=1, Q=2, W_logits=2, fixed=[2, 1, 1, 3]),
    "ragged_e3": dict(seed=42, E=3, W=5, S=5, Q=4, W_logits=5, fixed=None),
    "ones": dict(seed=43, E=2, W=5, S=1, Q=15, W_logits=5, fixed="ones"),
}


def split_repeats(c):
    n = c["E"] * c["W"] * c["Q"]
    if c["fixed"] == "ones":
        return np.ones(n, dtype=np.int64)
    if c["fixed"] is not None:
        return np.asarray(c["fixed"], dtype=np.int64)
    return _rng(c["seed"]).integers(1, 4, size=n).astype(np.int64)


def split_logits(c, n_rows):
    # coarse values so that window argmaxes tie often enough to exercise torch.mode's tie rule
    r = _rng(c["seed"] + 1000)
    return np.round(r.standard_normal((n_rows, c["W_logits"])) * 2.0).astype(np.float32) / 2.0


CI_DATA = tuple(float(v) for v in _rng(51).uniform(40.0, 90.0, size=37))


# ------------------------------------------------------------------ backbones
BACKBONE_CASES = {
    "conv64f_flat": ("Conv64F", dict(is_flatten=True, is_feature=False, leaky_relu=False, negative_slope=0.2,
                                     last_pool=True, maxpool_last2=True, num_channels=1)),
    "conv64f_dn4": ("Conv64F", dict(is_flatten=False, is_feature=False, leaky_relu=False, negative_slope=0.2,
                                    last_pool=False, maxpool_last2=True, num_channels=1)),
    "resnet12": ("resnet12", dict(keep_prob=0.0, avg_pool=True, is_flatten=True, maxpool_last2=True,
                                  num_channels=1)),
    "resnet12bdc": ("resnet12Bdc", dict(reduce_dim=64, num_channels=1)),
}


def backbone_input():
    return (_rng(61).standard_normal((2, 1, 128, 157)) * 0.5).astype(np.float32)


# the weight recipe and the waveform generator live in the product package (bench.py and smoke() must not depend on
# test infrastructure); the parity cases use the same functions under their historical names
from audio_fewshot_b200.synthetic import name_seeded_weights_ as perturb_bn_  # noqa: E402,F401
from audio_fewshot_b200.synthetic import synthetic_clip_batch  # noqa: E402,F401


# ------------------------------------------------------------------ spectrogram-domain augmentations
AUG_MEAN, AUG_STD = -15.114207, 26.22313  # Auxiliary/Clean_Mean_Std.npy
AUG_FIXED = {  # explicit kwargs: no random draw except the ones inside cutout / linear_filter
    "cutout": dict(num_cutouts=2, cutout_size_ratio=(0.1, 0.3), fill_value=0.0),
    "linear_filter": dict(num_points=4, filter_strength=0.5),
    "noise_suppression": dict(noise_percentile=20, suppression_strength=0.5),
    # smoothing_window=1: with any window > 1 the reference itself raises (F.pad(reflect) on a 4-D tensor with a
    # 2-tuple, audio_augmentations.py:432) -- recorded in the golden as "raises", see DESIGN.md
    "noise_matching": dict(target_noise_level=None, smoothing_window=1),
    "background_subtraction": dict(percentile=10),
    "contrast_enhancement": dict(contrast_factor=1.5, clip_percentile=95),
    "foreground_norm": dict(top_k_percent=20),
    "wiener_filter": dict(noise_floor_percentile=15, gain_factor=2.0),
}
AUG_RANDOM_SEEDS = tuple(range(100, 114))  # augmentation_type='random': the type and its parameters are drawn


def aug_input(shape=(2, 1, 32, 41), seed=81):
    """A NORMALISED spectrogram batch (what augment_spectrogram receives)."""
    return _rng(seed).standard_normal(shape).astype(np.float32)


# ------------------------------------------------------------------ MAML (config C5, shrunk)
MAML_CASE = dict(seed=71, torch_seed=5, E=2, W=5, S=2, Q=3, lr=0.01, train_iter=2, test_iter=3, feat_dim=1600,
                 backbone=dict(is_flatten=True, is_feature=False, leaky_relu=False, negative_slope=0.2,
                               last_pool=True, num_channels=1))
MAML_GRAD_KEYS = ("emb_func.layer1.0.weight", "emb_func.layer3.1.bias", "emb_func.logits.1.weight",
                  "emb_func.logits.2.bias", "classifier.layers.0.weight")


def maml_images(c=MAML_CASE):
    n = c["E"] * c["W"] * (c["S"] + c["Q"])
    return (_rng(c["seed"]).standard_normal((n, 1, 128, 157)) * 0.5).astype(np.float32)
