"""Parity of the fused log-mel kernel with the float64 spec (oracle/frontend.py).

Tolerance (north star: 1e-4 relative, fp32; norm fixed in SURVEY.md 7.3): on the de-normalised
dB value, |a - b| <= 1e-4 * max(|b|, 1)."""
import os

import numpy as np
import pytest
import torch

from oracle import frontend as fe
from oracle import philox

pytestmark = pytest.mark.gpu

MEAN, STD = -15.114207, 26.22313  # Auxiliary/Clean_Mean_Std.npy


def make_frontend(cuda, hop=512, n_mels=128, mean=MEAN, std=STD, **kw):
    from audio_fewshot_b200.frontend import LogMelFrontEnd
    return LogMelFrontEnd(hop_length=hop, n_mels=n_mels, mean=mean, std=std, **kw).to(cuda)


def assert_db_close(got, want_norm, std=STD, mean=MEAN, tol=1e-4):
    got_db = got.astype(np.float64) * std + mean
    want_db = want_norm * std + mean
    bound = tol * np.maximum(np.abs(want_db), 1.0)
    err = np.abs(got_db - want_db)
    assert (err <= bound).all(), "max dB err %.3e (bound %.3e)" % (err.max(), bound[err.argmax()])
    return err.max()


@pytest.mark.parametrize("B,L,hop,n_mels", [
    (3, 80000, 512, 128),   # S5: the repo's shape -> [B,1,128,157]
    (2, 16000, 102, 128),   # S1: 1 s clips -> [B,1,128,157]
    (2, 16000, 512, 128),   # 1 s at hop 512 -> 32 frames (one CTA chunk exactly)
    (1, 4099, 511, 128),    # odd length, odd hop: unaligned scalar-load path
    (2, 12345, 160, 80),    # fewer mel bins
    (1, 513, 512, 64),      # minimum legal length (reflect pad needs L > 512)
    (5, 33 * 512, 512, 128),  # 34 frames: second chunk holds 2 frames
])
def test_logmel_matches_float64_spec(cuda, B, L, hop, n_mels):
    rng = np.random.default_rng(L * 7 + hop)
    x = (rng.standard_normal((B, L)) * 0.1).astype(np.float32)
    x[0, : L // 3] += (0.5 * np.sin(np.arange(L // 3) * 0.3)).astype(np.float32)
    fr = make_frontend(cuda, hop, n_mels)
    got = fr(torch.from_numpy(x).to(cuda)).cpu().numpy()
    want = fe.logmel_f64(x, hop=hop, n_mels=n_mels, mean=MEAN, std=STD)
    assert got.shape == want.shape == (B, 1, n_mels, 1 + L // hop)
    assert_db_close(got, want)


def test_logmel_per_bin_statistics_and_silence(cuda):
    n_mels = 128
    mean = np.linspace(-30, -5, n_mels).astype(np.float32)
    std = np.linspace(10, 30, n_mels).astype(np.float32)
    x = np.zeros((2, 4096), np.float32)
    x[1] = (np.random.default_rng(1).standard_normal(4096) * 1e-3).astype(np.float32)
    fr = make_frontend(cuda, 256, n_mels, mean=mean, std=std)
    got = fr(torch.from_numpy(x).to(cuda)).cpu().numpy()
    want = fe.logmel_f64(x, hop=256, n_mels=n_mels, mean=mean, std=std)
    db_got = got[:, 0] * std[None, :, None] + mean[None, :, None]
    db_want = want[:, 0] * std[None, :, None] + mean[None, :, None]
    assert np.abs(db_got[0] - 10 * np.log10(fe.LOG_EPS)).max() < 1e-3  # silence -> 10 log10(eps)
    assert (np.abs(db_got - db_want) <= 1e-4 * np.maximum(np.abs(db_want), 1.0)).all()


def test_logmel_gain_linearity_and_batch_independence(cuda):
    """Size-independent properties at full bench shape (S5, 200 clips): scaling the waveform by g
    adds 20 log10 g dB; a clip's features do not depend on its batch neighbours."""
    rng = np.random.default_rng(11)
    B, L = 200, 80000
    x = torch.from_numpy((rng.standard_normal((B, L)) * 0.1).astype(np.float32)).to(cuda)
    fr = make_frontend(cuda, 512, 128, mean=0.0, std=1.0)
    a = fr(x)
    b = fr(x * 4.0)
    assert torch.allclose(b - a, torch.full_like(a, 20 * np.log10(4.0)), atol=2e-4)
    for i in (0, 57, 199):
        assert torch.equal(fr(x[i:i + 1]), a[i:i + 1])
    assert torch.equal(fr(x), a)  # deterministic


def test_logmel_matches_torch_fp32_path(cuda):
    """Cross-check with the PyTorch op sequence (torch.stft -> matmul -> log10) on the same device."""
    rng = np.random.default_rng(2)
    x = torch.from_numpy((rng.standard_normal((4, 80000)) * 0.1).astype(np.float32)).to(cuda)
    fr = make_frontend(cuda)
    got = fr(x)
    win = torch.from_numpy(fe.hann_periodic()).to(cuda)
    fb = torch.from_numpy(fe.mel_filterbank()).to(cuda)
    spec = torch.stft(x, 1024, 512, 1024, win, center=True, pad_mode="reflect", return_complex=True)
    mel = torch.matmul((spec.real ** 2 + spec.imag ** 2).transpose(1, 2), fb).transpose(1, 2)
    want = ((10 * torch.log10(mel + fe.LOG_EPS) - MEAN) / STD).unsqueeze(1)
    assert (got - want).abs().max().item() * STD < 2e-3  # the torch fp32/TF32 path is the looser one


@pytest.mark.parametrize("gain,shift,noise", [((-6.0, 6.0), 0, (0.0, 0.0)), ((0.0, 0.0), 300, (0.0, 0.0)),
                                              ((0.0, 0.0), 0, (0.01, 0.05)), ((-3.0, 3.0), 1000, (0.005, 0.02))])
def test_logmel_augmentation_matches_oracle(cuda, gain, shift, noise):
    rng = np.random.default_rng(21)
    B, L, seed, first = 6, 16000, 1234567890123, 40
    x = (rng.standard_normal((B, L)) * 0.1).astype(np.float32)
    aug = dict(gain_db=gain, max_shift=shift, noise_std=noise)
    fr = make_frontend(cuda, 512, 128, aug=aug, seed=seed)
    fr.train()
    got = fr(torch.from_numpy(x).to(cuda), first_clip_index=first).cpu().numpy()
    y = philox.augment_waveform(x, seed, first, gain, shift, noise)
    want = fe.logmel_f64(y, hop=512, mean=MEAN, std=STD)
    assert_db_close(got, want, tol=3e-4)  # Box-Muller in fp32 on the device vs fp64 in the oracle
    fr.eval()  # augmentation is a training-time transform
    plain = fr(torch.from_numpy(x).to(cuda)).cpu().numpy()
    assert_db_close(plain, fe.logmel_f64(x, hop=512, mean=MEAN, std=STD))
    ks = {philox.clip_params(seed, first + b, gain[0], gain[1], shift, noise[0], noise[1])[1] for b in range(B)}
    if shift:
        assert len(ks) > 1  # different clips draw different shifts


@pytest.mark.parametrize("B,L,hop,aug", [
    (3, 80000, 512, None),                # aligned short2 loads
    (2, 4099, 511, None),                 # odd length and hop: scalar int16 loads, reflected edges
    (4, 16000, 512, dict(gain_db=(-3.0, 3.0), max_shift=300, noise_std=(0.005, 0.02))),
])
def test_logmel_pcm16_is_bit_identical_to_converted_fp32(cuda, B, L, hop, aug):
    """16-bit PCM input (afs_logmel_fwd_pcm16): sample = pcm / 32768 is exact in fp32, so the features must equal
    bit for bit those of the fp32 entry point on the converted waveform -- and therefore meet the float64 spec."""
    rng = np.random.default_rng(L + hop)
    pcm = np.clip(np.round(rng.standard_normal((B, L)) * 0.1 * 32768.0), -32768, 32767).astype(np.int16)
    pcm[0, :7] = [-32768, 32767, 0, 1, -1, 12345, -12345]
    x = pcm.astype(np.float32) / np.float32(32768.0)
    fr = make_frontend(cuda, hop, 128, aug=aug, seed=77)
    if aug is not None:
        fr.train()
    a = fr(torch.from_numpy(pcm).to(cuda), first_clip_index=5)
    b = fr(torch.from_numpy(x).to(cuda), first_clip_index=5)
    assert a.dtype == torch.float32 and torch.equal(a, b)
    if aug is None:
        assert_db_close(a.cpu().numpy(), fe.logmel_f64(x, hop=hop, mean=MEAN, std=STD))
    half = make_frontend(cuda, hop, 128, pcm_scale=1.0 / 65536.0)  # another power-of-two scale
    assert torch.equal(half(torch.from_numpy(pcm).to(cuda)), half(torch.from_numpy(x * np.float32(0.5)).to(cuda)))


def test_logmel_rejects_bad_input(cuda):
    from audio_fewshot_b200._lib import AfsError
    fr = make_frontend(cuda)
    with pytest.raises(AfsError):
        fr(torch.zeros(1, 400, device=cuda))  # reflect padding needs L > n_fft/2
    with pytest.raises(AfsError):
        fr(torch.zeros(1, 4000))  # CPU tensor: no fallback
    with pytest.raises(TypeError):
        fr(torch.zeros(1, 4000, device=cuda, dtype=torch.float16))  # fp32 or int16 PCM only
