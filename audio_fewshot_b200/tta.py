"""Energy-gated test-time augmentation: the caller-side step right after DeepBDC.set_forward.

Mirrors the reference's evaluation loop body (libfewshot_core/test.py:380-414) and its two helpers
(`map_q_to_s_runs` test.py:33-72, `augment_images_with_mask` test.py:74-152):

    1. model(batch, enhance_classification_via_energy=True) -> (output, acc, uncertains, ood_query_mask, query_mask);
    2. every window of a query flagged out-of-distribution (top 20 % energy, deepbdc.py:326-351) is replaced by
       `num_augmentations` augmented copies (augment_spectrogram(..., augmentation_type='noise_suppression') with the
       Clean_Mean_Std statistics), the other rows are kept, order preserved;
    3. the flagged queries' `repeats` grow accordingly and the model votes again over the enlarged window sets.

The augmentations run on the GPU (one fused kernel per augmented window, csrc/specaug.cu) with parameters drawn from
Python's `random` in the reference's call order, so a seeded run reproduces the reference's images.

Reference quirk (kept out, documented): test.py:405 does `repeats[idxs] += num_augmentations - 1`, which matches the
number of rows produced only for queries with ONE window (r windows become r * num_augmentations rows, not
r + num_augmentations - 1); `updated_repeats` below returns r * num_augmentations, which coincides for r == 1.
"""
import numpy as np
import torch

from . import augment as _augment


def map_q_to_s_runs(s, r, q):
    """Per-row flags from per-query flags: row i (a query row where s[i]) gets q[j] of the run j it belongs to
    (runs of lengths r over the True entries of s, in order).  test.py:33-72, vectorised."""
    s = np.asarray(s, dtype=bool)
    r = np.asarray(r, dtype=int)
    q = np.asarray(q, dtype=bool)
    if r.sum() != s.sum():
        raise ValueError("Sum of r (%d) must equal number of True in s (%d)" % (r.sum(), s.sum()))
    if len(r) != len(q):
        raise ValueError("Length of r (%d) must equal length of q (%d)" % (len(r), len(q)))
    mapped = np.zeros_like(s, dtype=bool)
    mapped[np.flatnonzero(s)] = np.repeat(q, r)
    return mapped


def augment_images_with_mask(images, repeats, is_query_mask, mask, augmentation_fn, num_augmentations=10):
    """Rows of `images` that belong to a flagged query are replaced by `num_augmentations` outputs of
    `augmentation_fn(row[None])` (called once per copy, in row order -- the reference's call order, so random
    parameters are drawn identically); all other rows are kept.  test.py:74-152."""
    repeats_np = repeats.cpu().numpy() if isinstance(repeats, torch.Tensor) else np.asarray(repeats)
    mask_np = mask.cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
    flagged = map_q_to_s_runs(is_query_mask, repeats_np, mask_np)
    counts = np.where(flagged, num_augmentations, 1)
    starts = np.concatenate([[0], np.cumsum(counts)])
    out = torch.empty((int(starts[-1]),) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    keep = np.flatnonzero(~flagged)
    if keep.size:
        out[torch.as_tensor(starts[keep], device=images.device)] = images[torch.as_tensor(keep, device=images.device)]
    for i in np.flatnonzero(flagged):
        for a in range(num_augmentations):
            res = augmentation_fn(images[i].clone().unsqueeze(0))
            res = res if isinstance(res, torch.Tensor) else torch.stack(list(res), dim=0)
            out[starts[i] + a] = res.reshape(images.shape[1:])
    return out


def updated_repeats(repeats, ood_query_mask, num_augmentations):
    """Windows per query after augment_images_with_mask: flagged queries have num_augmentations x as many."""
    rep = repeats.clone() if isinstance(repeats, torch.Tensor) else torch.as_tensor(np.asarray(repeats)).clone()
    idx = torch.as_tensor(np.flatnonzero(np.asarray(ood_query_mask)), dtype=torch.long)
    rep[idx] = rep[idx] * num_augmentations
    return rep


def energy_tta_step(model, batch, num_augmentations=10, mean=0.0, std=1.0, augmentation_type="noise_suppression",
                    **aug_kwargs):
    """One evaluation batch with energy-gated test-time augmentation (test.py:380-414).  `batch` is the reference's
    flat list [image, global_target, repeats, support_size]; the model must implement the 5-tuple return of
    DeepBDC.set_forward(enhance_classification_via_energy=True).  Returns (acc, info) where acc is the accuracy after
    re-voting (the first-pass accuracy when nothing was flagged) and info holds the first-pass results."""
    image, global_target, repeats, support_size = batch
    _, acc0, uncertains, ood_mask, query_mask = model.set_forward(
        [image, global_target, repeats, support_size], update_threshold=False, enhance_classification_via_energy=True)
    info = {"acc_before": acc0, "uncertains": uncertains, "ood_query_mask": ood_mask, "query_mask": query_mask,
            "n_flagged": int(np.asarray(ood_mask).sum())}
    if info["n_flagged"] == 0:
        return acc0, info
    image = image.to(model.device, non_blocking=True)
    fn = lambda x: _augment.augment_spectrogram(x, mean=mean, std=std, augmentation_type=augmentation_type, **aug_kwargs)
    augmented = augment_images_with_mask(image, repeats, query_mask, ood_mask, fn, num_augmentations)
    rep2 = updated_repeats(repeats, ood_mask, num_augmentations)
    _, acc1 = model.set_forward([augmented, global_target, rep2, support_size])[:2]
    info["repeats_after"] = rep2
    return acc1, info
