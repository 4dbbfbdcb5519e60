"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the waveform -> normalised log-mel front-end.

PARITY UNPINNED BY THE REFERENCE: the reference ships no waveform code at all
(SURVEY.md F2 -- datasets are pre-computed `*_spec` folders, reference
config/headers/data.yaml:1), so no golden vector of the reference can pin this
stage.  The spec below is the canonical definition for this repo;
it is cross-checked against torchaudio.transforms.MelSpectrogram (MetaAudio /
librosa convention inferred from the [1,128,157] input the backbones require,
reference libfewshot_core/model/backbone/conv_four.py:87) in tests/test_oracle.py.
The only reference-pinned part is the final normalisation
(x - mean) / std  -- reference libfewshot_core/audio_augmentations.py:36-53 with
the (2,1,1) [mean, std] arrays of Auxiliary/*_Mean_Std.npy (test.py:398-399).

    x[B, L] fp32 -> reflect-pad n_fft/2 -> frames (hop) * hann_periodic(n_fft)
      -> |rFFT|^2 [513] -> mel (slaney scale, slaney norm, f_min 0, f_max sr/2) [n_mels]
      -> log_mult * log10(mel + log_eps) -> (. - mean[m]) / std[m] -> [B, 1, n_mels, T]
"""
import math

import numpy as np

N_FFT = 1024
LOG_EPS = 2.220446049250313e-16  # float64 machine epsilon, as SURVEY.md 8c


def hann_periodic(n=N_FFT, dtype=np.float32):
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(dtype)


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_freqs=513, f_min=0.0, f_max=8000.0, n_mels=128, sample_rate=16000):
    """[n_freqs, n_mels] float32; slaney mel scale + slaney area normalisation
    (the arithmetic of torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney'))."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_slaney(f_min)
    m_max = _hz_to_mel_slaney(f_max)
    m_pts = np.linspace(m_min, m_max, n_mels + 2)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])
    fb = fb * enorm[None, :]
    return fb.astype(np.float32)


def num_frames(L, hop, center=True, n_fft=N_FFT):
    return 1 + L // hop if center else 1 + (L - n_fft) // hop


def frames(x, hop, center=True, n_fft=N_FFT):
    """[B, L] -> [B, T, n_fft] (reflect padding as torch.stft(center=True, pad_mode='reflect'))."""
    x = np.asarray(x)
    if center:
        x = np.pad(x, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
    T = 1 + (x.shape[1] - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    return x[:, idx]


def logmel_f64(x, hop=512, n_mels=128, sample_rate=16000, mean=None, std=None, center=True,
               log_mult=10.0, log_eps=LOG_EPS, fb=None, window=None):
    """float64 evaluation of the spec (used to measure how far fp32 implementations sit from
    the exact value).  fb/window default to the fp32-rounded tables the kernel receives."""
    x = np.asarray(x, dtype=np.float32).astype(np.float64)
    if fb is None:
        fb = mel_filterbank(N_FFT // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate)
    if window is None:
        window = hann_periodic()
    fr = frames(x, hop, center) * window.astype(np.float64)[None, None, :]
    spec = np.fft.rfft(fr, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2  # [B, T, 513]
    mel = power @ fb.astype(np.float64)  # [B, T, n_mels]
    db = log_mult * np.log10(mel + log_eps)
    db = np.transpose(db, (0, 2, 1))  # [B, n_mels, T]
    if mean is not None:
        mean = np.broadcast_to(np.asarray(mean, dtype=np.float64).reshape(-1), (n_mels,))
        std = np.broadcast_to(np.asarray(std, dtype=np.float64).reshape(-1), (n_mels,))
        db = (db - mean[None, :, None]) / std[None, :, None]
    return db[:, None, :, :]


def logmel_torch(x, hop=512, n_mels=128, sample_rate=16000, mean=None, std=None, center=True,
                 log_mult=10.0, log_eps=LOG_EPS, fb=None, window=None):
    """fp32 evaluation with the PyTorch op sequence a user of the reference would write
    (torch.stft -> abs^2 -> matmul(fb) -> log10 -> normalise); this is the CPU baseline arm."""
    import torch

    x = torch.as_tensor(np.asarray(x), dtype=torch.float32)
    if fb is None:
        fb = mel_filterbank(N_FFT // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate)
    if window is None:
        window = hann_periodic()
    fb = torch.as_tensor(np.asarray(fb), dtype=torch.float32)
    win = torch.as_tensor(np.asarray(window), dtype=torch.float32)
    spec = torch.stft(x, n_fft=N_FFT, hop_length=hop, win_length=N_FFT, window=win, center=center,
                      pad_mode="reflect", return_complex=True)  # [B, 513, T]
    power = spec.real ** 2 + spec.imag ** 2
    mel = torch.matmul(power.transpose(1, 2), fb).transpose(1, 2)  # [B, n_mels, T]
    db = log_mult * torch.log10(mel + log_eps)
    if mean is not None:
        mean_t = torch.as_tensor(np.asarray(mean), dtype=torch.float32).reshape(-1)
        std_t = torch.as_tensor(np.asarray(std), dtype=torch.float32).reshape(-1)
        db = (db - mean_t.reshape(1, -1, 1)) / std_t.reshape(1, -1, 1)
    return db.unsqueeze(1)


def load_mean_std(path):
    """Auxiliary/*_Mean_Std.npy: float32 (2,1,1) = [mean, std] (reference test.py:398-399)."""
    arr = np.load(path)
    mean, std = arr.flatten().tolist()
    return float(mean), float(std)
