"""Pins the CPU oracle: against the committed goldens (generated from the REAL reference by
oracle/make_golden.py) and, in the authoring container, against the reference classes directly."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import backbones as obb
from oracle import cases, frontend as fe, heads


def _split_fixed(feat, c):
    E, W, S, Q = c["E"], c["W"], c["S"], c["Q"]
    f = feat.view(E, W, S + Q, *feat.shape[1:])
    sup = f[:, :, :S].contiguous().view(E, W * S, *feat.shape[1:])
    qry = f[:, :, S:].contiguous().view(E, W * Q, *feat.shape[1:])
    return sup, qry


@pytest.mark.parametrize("name", sorted(cases.PROTO_CASES))
def test_proto_layer_matches_reference_golden(golden, name):
    c = cases.PROTO_CASES[name]
    sup, qry = _split_fixed(torch.from_numpy(cases.proto_features(c)), c)
    if c["head"] == "proto":
        got = heads.proto_layer(qry, sup, c["W"], c["S"], c["mode"])
    else:
        got = heads.deepbdc_proto_layer(qry, sup, c["W"], c["S"])
    want = golden("proto_layer.npz")[name]
    np.testing.assert_allclose(got.reshape(-1, c["W"]).numpy(), want, rtol=1e-5, atol=1e-5)
    assert np.array_equal(got.reshape(-1, c["W"]).numpy().argmax(1), want.argmax(1))


@pytest.mark.parametrize("name", sorted(cases.DN4_CASES))
def test_dn4_layer_matches_reference_golden(golden, name):
    c = cases.DN4_CASES[name]
    sup, qry = _split_fixed(torch.from_numpy(cases.dn4_features(c)), c)
    got = heads.dn4_layer(qry, sup, c["W"], c["S"], c["n_k"]).reshape(-1, c["W"]).numpy()
    want = golden("dn4_layer.npz")[name]
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", sorted(cases.BDC_CASES))
def test_bdc_matches_reference_golden(golden, name):
    c = cases.BDC_CASES[name]
    x = torch.from_numpy(cases.bdc_features(c))
    full = heads.bdcovpool(x, torch.full((1, 1), c["log_temp"]))
    g = golden("bdc_pool.npz")
    np.testing.assert_allclose(full.numpy(), g[name + "/full"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(heads.triuvec(full).numpy(), g[name + "/triu"], rtol=1e-5, atol=1e-6)
    # BDC matrix is symmetric with (numerically) zero row/column means
    assert np.abs(full.numpy() - full.numpy().transpose(0, 2, 1)).max() < 1e-4
    assert np.abs(full.numpy().mean(axis=2)).max() < 1e-4


@pytest.mark.parametrize("name", sorted(cases.SPLIT_CASES))
def test_split_vote_energy_match_reference_golden(golden, name):
    c = cases.SPLIT_CASES[name]
    g = golden("episode_vote.npz")
    E, W, S, Q = c["E"], c["W"], c["S"], c["Q"]
    rep = cases.split_repeats(c)
    n_rows = E * W * S + int(rep.sum())
    feats = torch.arange(n_rows, dtype=torch.float32).view(-1, 1).repeat(1, 2)
    sup, qry, st, qt, mask = heads.split_by_episode(feats, W, S, Q, torch.from_numpy(rep), E * W * S)
    assert np.array_equal(sup[..., 0].numpy().astype(np.int32), g[name + "/support_rows"])
    assert np.array_equal(np.concatenate([q[:, 0].numpy() for q in qry]).astype(np.int32), g[name + "/query_rows"])
    assert np.array_equal(np.asarray([q.shape[0] for q in qry]), g[name + "/query_len"])
    assert np.array_equal(qt.numpy(), g[name + "/query_target"])
    assert np.array_equal(mask, g[name + "/query_mask"])
    logits = torch.from_numpy(cases.split_logits(c, int(rep.sum())))
    pred = heads.majority_vote(logits, rep)
    assert np.array_equal(pred.numpy().astype(np.int32), g[name + "/vote_pred"])
    acc = heads.vote_categorical_acc(qt.reshape(-1), pred.to(torch.long)).item()
    assert acc == pytest.approx(float(g[name + "/vote_acc"]))
    np.testing.assert_allclose(heads.energy_score(logits, rep).numpy(), g[name + "/energy"], rtol=1e-6, atol=1e-6)
    # the device table the kernels use describes exactly the same rows
    cls_row = heads.cls_row_table(rep, E, W, S)
    sup_rows = np.concatenate([np.arange(cls_row[k], cls_row[k] + S) for k in range(E * W)])
    assert np.array_equal(sup_rows, g[name + "/support_rows"].reshape(-1))
    qry_rows = np.concatenate([np.arange(cls_row[k] + S, cls_row[k + 1]) for k in range(E * W)])
    assert np.array_equal(qry_rows, g[name + "/query_rows"])


def test_confidence_interval_matches_reference_golden(golden):
    m, h = heads.mean_confidence_interval(list(cases.CI_DATA))
    np.testing.assert_allclose([m, h], golden("episode_vote.npz")["ci/mean_h"], rtol=1e-12)


def test_repeats_ones_equals_fixed_layout():
    c = dict(E=2, W=5, S=5, Q=3)
    feat = torch.randn(2 * 5 * 8, 6)
    a = heads.split_by_episode(feat, 5, 5, 3)
    b = heads.split_by_episode(feat, 5, 5, 3, torch.ones(30, dtype=torch.long), 50)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], torch.stack(b[1]))


@pytest.mark.parametrize("name", sorted(cases.BACKBONE_CASES))
def test_backbone_restatement_matches_reference_golden(golden, name):
    from audio_fewshot_b200 import model as arch
    ctor, kwargs = cases.BACKBONE_CASES[name]
    g = golden("backbones.npz")
    net = getattr(arch, ctor)(**kwargs).eval()
    assert sorted(net.state_dict().keys()) == list(g[name + "/keys"])  # same checkpoint layout
    cases.perturb_bn_(net)
    sd = net.state_dict()
    x = torch.from_numpy(cases.backbone_input())
    with torch.no_grad():
        if ctor == "Conv64F":
            y = obb.conv64f_forward(sd, x, kwargs["is_flatten"], kwargs["last_pool"], kwargs["maxpool_last2"])
        elif ctor == "resnet12":
            y = obb.resnet12_forward(sd, x)
        else:
            y = obb.resnet12bdc_forward(sd, x)
    want = g[name + "/out"]
    scale = np.abs(want).max()
    assert np.abs(y.numpy() - want).max() <= 2e-4 * scale
    if ctor != "resnet12Bdc":  # the product module itself (BdcPool needs CUDA: covered by the gpu tests)
        with torch.no_grad():
            assert np.abs(net(x).numpy() - want).max() <= 2e-4 * scale


def test_frontend_spec_matches_torchaudio():
    torchaudio = pytest.importorskip("torchaudio")
    x = (np.random.default_rng(3).standard_normal((2, 16000)) * 0.1).astype(np.float32)
    ms = torchaudio.transforms.MelSpectrogram(16000, n_fft=1024, hop_length=512, n_mels=128, f_min=0, f_max=8000,
                                              power=2.0, norm="slaney", mel_scale="slaney")
    want = 10 * torch.log10(ms(torch.from_numpy(x)) + fe.LOG_EPS).unsqueeze(1).numpy()
    got = fe.logmel_torch(x).numpy()
    assert got.shape == (2, 1, 128, 32)
    assert np.abs(got - want).max() < 5e-4  # torchaudio builds its filterbank in fp32
    assert np.abs(fe.logmel_f64(x) - got).max() < 1e-4
    assert fe.logmel_torch(np.zeros((1, 80000), np.float32)).shape == (1, 1, 128, 157)


def test_mean_std_files_are_scalar_pairs():
    import os
    from conftest import ROOT
    p = os.path.join(ROOT, "tests", "golden", "Clean_Mean_Std.npy")
    mean, std = fe.load_mean_std(p)
    assert mean == pytest.approx(-15.114207, abs=1e-4) and std == pytest.approx(26.22313, abs=1e-4)


# ---------------------------------------------------------------- direct checks against the reference
def test_oracle_heads_equal_reference_classes(reference):
    arch, utils, _ = reference
    from libfewshot_core.model.metric.proto_net import ProtoLayer
    from libfewshot_core.model.metric.dn4 import DN4Layer
    torch.manual_seed(5)
    q, s = torch.randn(2, 30, 96), torch.randn(2, 10, 96)
    for mode in ("euclidean", "cos_sim"):
        assert torch.equal(ProtoLayer()(q, s, 5, 2, 6, mode=mode), heads.proto_layer(q, s, 5, 2, mode))
    q5, s5 = torch.rand(1, 12, 16, 3, 4), torch.rand(1, 6, 16, 3, 4)
    assert torch.equal(DN4Layer(2)(q5, s5, 3, 2, 4), heads.dn4_layer(q5, s5, 3, 2, 2))
    logits = torch.randn(9, 4)
    rep = torch.tensor([2, 1, 3, 3])
    assert torch.equal(utils.majority_vote(logits.softmax(1), rep), heads.majority_vote(logits, rep))
    assert torch.equal(utils.average_logits(logits, rep), heads.average_logits(logits, rep))


def test_oracle_split_equals_reference_on_random_ragged(reference):
    arch, _, _ = reference
    rng = np.random.default_rng(9)
    for trial in range(5):
        E, W, S, Q = int(rng.integers(1, 4)), int(rng.integers(2, 6)), int(rng.integers(1, 4)), int(rng.integers(1, 5))
        rep = torch.from_numpy(rng.integers(1, 4, size=E * W * Q))
        n = E * W * S + int(rep.sum())
        feat = torch.randn(n, 3, 2, 2)
        model = arch.DN4(way_num=W, shot_num=S, query_num=Q, test_way=W, test_shot=S, test_query=Q, emb_func=None,
                         device="cpu")
        a = model.split_by_episode(feat, mode=2, repeats=rep, support_size=E * W * S)
        b = heads.split_by_episode(feat, W, S, Q, rep, E * W * S)
        assert torch.equal(a[0], b[0])
        assert all(torch.equal(x, y) for x, y in zip(a[1], b[1]))
        assert torch.equal(a[3], b[3]) and np.array_equal(a[4], b[4])


def test_torch_mode_cuda_rule_matches_measured_golden(golden):
    """torch.mode on CUDA does not return the smallest tied label; the oracle restates its rule and
    is pinned by outputs recorded from torch 2.11 on a B200 (tools/probe_torch_mode.py)."""
    g = golden("torch_mode_cuda.npz")
    differs_from_cpu_rule = 0
    for lab, n, m in zip(g["labels"], g["n"], g["mode"]):
        y = lab[:n].astype(np.int64)
        assert heads.torch_mode_cuda(y) == int(m)
        differs_from_cpu_rule += int(torch.mode(torch.from_numpy(y))[0].item() != int(m))
    assert differs_from_cpu_rule > 100  # the two rules really are different


@pytest.mark.parametrize("case", ["s5", "s1", "odd"])
def test_frontend_spec_matches_the_torchaudio_golden(case):
    """Pins oracle/frontend.py to the committed torchaudio outputs (tests/golden/logmel_torchaudio.npz, written by
    oracle/make_frontend_golden.py): with torchaudio's own fp32 filterbank and window the float64 spec reproduces
    torchaudio's dB values to 1e-5 -- the only external anchor the waveform stage can have."""
    from oracle.make_frontend_golden import CASES, waveform
    g = np.load(os.path.join(GOLDEN, "logmel_torchaudio.npz"), allow_pickle=False)
    c = CASES[case]
    x = waveform(c)
    got = fe.logmel_f64(x, hop=c["hop"], n_mels=c["n_mels"], fb=g[case + "_fb"], window=g[case + "_window"])[:, 0]
    assert got.shape == g[case + "_db"].shape
    assert np.abs(got - g[case + "_db"]).max() < 1e-5
    # and the repo's own tables (float64 design, rounded once) are torchaudio's up to its fp32 construction
    assert np.abs(fe.mel_filterbank(n_mels=c["n_mels"]) - g[case + "_fb"]).max() < 1e-6
    assert np.abs(fe.hann_periodic() - g[case + "_window"]).max() < 3e-7  # torch.hann_window evaluates the cosine in fp32
