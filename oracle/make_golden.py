"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz from the REAL reference classes.

Run in the authoring container (needs /root/reference):

    python -m oracle.make_golden

Inputs are NOT stored: every case regenerates them with
numpy.random.default_rng(seed) (PCG64, stable across numpy versions) through
`oracle.cases`, so the fixtures stay small (outputs only).  The GPU box has no
/root/reference; tests there compare the CUDA kernels with these files.
"""
import os
import sys

import numpy as np
import torch

from . import cases
from .ref_import import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def make_maml():
    """MAML.set_forward_loss (maml.py:91-123) of the real reference on the CPU.  The reference's
    BatchNorm2d_fw calls .cuda() on its scratch buffers (maml_module.py:85-86); on this GPU-less
    machine Tensor.cuda is patched to the identity for the duration of the run -- nothing else changes."""
    arch, utils, _ = import_reference()
    c = cases.MAML_CASE
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        torch.manual_seed(0)
        emb = arch.Conv64F(**c["backbone"])
        model = arch.MAML(inner_param=dict(lr=c["lr"], train_iter=c["train_iter"], test_iter=c["test_iter"]),
                          feat_dim=c["feat_dim"], way_num=c["W"], shot_num=c["S"], query_num=c["Q"],
                          test_way=c["W"], test_shot=c["S"], test_query=c["Q"], emb_func=emb, device="cpu")
        cases.perturb_bn_(model)
        model.train()
        x = torch.from_numpy(cases.maml_images(c))
        torch.manual_seed(c["torch_seed"])  # Dropout(0.3) of Conv64F.logits is live during adaptation
        output, acc, loss = model([x, torch.zeros(x.shape[0])])
        loss.backward()
        named = dict(model.named_parameters())
        out = {"output": output.detach().numpy(), "acc": np.asarray(acc), "loss": np.asarray(loss.item()),
               "keys": np.asarray(sorted(model.state_dict().keys()))}
        for k in cases.MAML_GRAD_KEYS:
            out["grad/" + k] = named[k].grad.numpy()
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez_compressed(os.path.join(OUT, "maml.npz"), **out)
    print("maml.npz", os.path.getsize(os.path.join(OUT, "maml.npz")), "loss", out["loss"], "acc", out["acc"])


def make_augment():
    """augment_spectrogram of the real reference (audio_augmentations.py:531-604) on seeded planes."""
    import random
    _, _, aug = import_reference()
    out = {}
    x = torch.from_numpy(cases.aug_input())
    for t, kw in cases.AUG_FIXED.items():
        random.seed(7)
        out["fixed/" + t] = aug.augment_spectrogram(x, cases.AUG_MEAN, cases.AUG_STD, augmentation_type=t, **kw).numpy()
    raised = []
    for sd in cases.AUG_RANDOM_SEEDS:
        random.seed(sd)
        try:
            out["random/%d" % sd] = aug.augment_spectrogram(x, cases.AUG_MEAN, cases.AUG_STD).numpy()
        except NotImplementedError:  # 'noise_matching' drawn: the reference's reflect pad fails (see cases.py)
            raised.append(sd)
    out["random/raised"] = np.asarray(raised, dtype=np.int64)
    random.seed(7)
    try:
        aug.augment_spectrogram(x, cases.AUG_MEAN, cases.AUG_STD, augmentation_type="noise_matching", smoothing_window=5)
        out["noise_matching_window5_raises"] = np.asarray(0)
    except NotImplementedError:
        out["noise_matching_window5_raises"] = np.asarray(1)
    big = torch.from_numpy(cases.aug_input((1, 1, 128, 157), seed=82))
    for t in ("noise_suppression", "background_subtraction"):
        random.seed(9)
        out["full/" + t] = aug.augment_spectrogram(big, cases.AUG_MEAN, cases.AUG_STD, augmentation_type=t,
                                                   **cases.AUG_FIXED[t]).numpy()
    np.savez_compressed(os.path.join(OUT, "augment.npz"), **out)
    print("augment.npz", os.path.getsize(os.path.join(OUT, "augment.npz")))


def tta_cases():
    """Seeded (query-row mask, windows per query, flagged queries, images) cases for the TTA helpers."""
    rng = np.random.default_rng(77)
    out = []
    for _ in range(6):
        nq = int(rng.integers(1, 8))
        r = rng.integers(1, 4, size=nq)
        s = []
        for j in range(nq):
            s += [False] * int(rng.integers(0, 3)) + [True] * int(r[j])
        s = np.array(s + [False] * int(rng.integers(0, 4)), dtype=bool)
        q = rng.random(nq) < 0.4
        imgs = rng.standard_normal((len(s), 1, 4, 5)).astype(np.float32)
        out.append((s, r.astype(np.int64), q, imgs))
    return out


def make_tta():
    """map_q_to_s_runs / augment_images_with_mask of the real reference (test.py:33-152) on seeded cases, with a
    call-counting augmentation so the call ORDER is pinned too."""
    import_reference()
    import libfewshot_core.test as rt
    out = {}
    for i, (s, r, q, imgs) in enumerate(tta_cases()):
        out["%d/s" % i], out["%d/r" % i], out["%d/q" % i], out["%d/img" % i] = s, r, q, imgs
        out["%d/mapped" % i] = rt.map_q_to_s_runs(s, r, q)
        calls = [0]

        def fn(x):
            calls[0] += 1
            return x * 2 + calls[0]

        out["%d/aug" % i] = rt.augment_images_with_mask(torch.from_numpy(imgs), torch.from_numpy(r), s, q, fn, 3).numpy()
    np.savez_compressed(os.path.join(OUT, "tta.npz"), **out)
    print("tta.npz", os.path.getsize(os.path.join(OUT, "tta.npz")))


def main():
    if "--tta" in sys.argv:
        return make_tta()
    if "--maml" in sys.argv:
        return make_maml()
    if "--augment" in sys.argv:
        return make_augment()
    arch, utils, _ = import_reference()
    from libfewshot_core.model.metric.proto_net import ProtoLayer
    from libfewshot_core.model.metric.dn4 import DN4Layer
    from libfewshot_core.model.metric import deepbdc as ref_deepbdc
    from libfewshot_core.model.backbone.utils.bdc_pool import BDCovpool, Triuvec

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    # ---- ProtoLayer (proto_net.py:34-64), fixed layout
    out = {}
    for name, c in cases.PROTO_CASES.items():
        feat = torch.from_numpy(cases.proto_features(c))
        E, W, S, Q, D = c["E"], c["W"], c["S"], c["Q"], c["D"]
        f = feat.view(E, W, S + Q, D)
        sup = f[:, :, :S].contiguous().view(E, W * S, D)
        qry = f[:, :, S:].contiguous().view(E, W * Q, D)
        if c["head"] == "proto":
            logits = ProtoLayer()(qry, sup, W, S, Q, mode=c["mode"])
        else:
            logits = ref_deepbdc.ProtoLayer()(qry, sup, W, S, Q)
        out[name] = logits.reshape(-1, W).numpy()
    np.savez_compressed(os.path.join(OUT, "proto_layer.npz"), **out)

    # ---- DN4Layer (dn4.py:39-75)
    out = {}
    for name, c in cases.DN4_CASES.items():
        feat = torch.from_numpy(cases.dn4_features(c))
        E, W, S, Q, C, H, Wd = c["E"], c["W"], c["S"], c["Q"], c["C"], c["H"], c["Wd"]
        f = feat.view(E, W, S + Q, C, H, Wd)
        sup = f[:, :, :S].contiguous().view(E, W * S, C, H, Wd)
        qry = f[:, :, S:].contiguous().view(E, W * Q, C, H, Wd)
        score = DN4Layer(c["n_k"])(qry, sup, W, S, Q)
        out[name] = score.reshape(-1, W).numpy()
    np.savez_compressed(os.path.join(OUT, "dn4_layer.npz"), **out)

    # ---- BDCovpool + Triuvec (bdc_pool.py:69-93)
    out = {}
    for name, c in cases.BDC_CASES.items():
        x = torch.from_numpy(cases.bdc_features(c))
        t = torch.full((1, 1), float(c["log_temp"]))
        full = BDCovpool(x, t)
        out[name + "/full"] = full.numpy()
        out[name + "/triu"] = Triuvec(full).reshape(x.shape[0], -1).numpy()
    np.savez_compressed(os.path.join(OUT, "bdc_pool.npz"), **out)

    # ---- split_by_episode (abstract_model.py:176-332), majority_vote, vote acc (utils.py:432-446)
    out = {}
    for name, c in cases.SPLIT_CASES.items():
        E, W, S, Q = c["E"], c["W"], c["S"], c["Q"]
        model = arch.ProtoNet(way_num=W, shot_num=S, query_num=Q, test_way=W, test_shot=S, test_query=Q,
                              emb_func=None, device="cpu")
        repeats = torch.from_numpy(cases.split_repeats(c))
        n_rows = E * W * S + int(repeats.sum())
        feats = torch.arange(n_rows, dtype=torch.float32).view(-1, 1).repeat(1, 2)  # row id as the feature
        sup, qry, st, qt, mask = model.split_by_episode(feats, mode=1, repeats=repeats, support_size=E * W * S)
        out[name + "/support_rows"] = sup[..., 0].numpy().astype(np.int32)
        out[name + "/query_rows"] = np.concatenate([q[:, 0].numpy() for q in qry]).astype(np.int32)
        out[name + "/query_len"] = np.asarray([q.shape[0] for q in qry], dtype=np.int32)
        out[name + "/query_target"] = qt.numpy().astype(np.int32)
        out[name + "/query_mask"] = mask
        logits = torch.from_numpy(cases.split_logits(c, int(repeats.sum())))
        pred = utils.majority_vote(logits.softmax(dim=1), repeats)
        out[name + "/vote_pred"] = pred.numpy().astype(np.int32)
        out[name + "/vote_acc"] = np.asarray(utils.vote_catagorical_acc(qt.reshape(-1), pred.to(torch.long)).item())
        avg = utils.average_logits(logits, repeats)
        out[name + "/energy"] = (-torch.logsumexp(avg, dim=1)).numpy()
    m, h = utils.mean_confidence_interval(list(cases.CI_DATA))
    out["ci/mean_h"] = np.asarray([m, h])
    np.savez_compressed(os.path.join(OUT, "episode_vote.npz"), **out)

    # ---- backbones (conv_four.py, resnet_12.py, resnet_bdc.py): eval features of seeded weights
    out = {}
    x = torch.from_numpy(cases.backbone_input())
    for name, (ctor, kwargs) in cases.BACKBONE_CASES.items():
        torch.manual_seed(0)
        net = getattr(arch, ctor)(**kwargs).eval()
        cases.perturb_bn_(net)
        with torch.no_grad():
            y = net(x)
        out[name + "/out"] = y.numpy()
        out[name + "/keys"] = np.asarray(sorted(net.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, "backbones.npz"), **out)
    make_maml()
    make_augment()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
