"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/logmel_torchaudio.npz: the waveform -> log-mel stage as
torchaudio.transforms.MelSpectrogram computes it (the spec SURVEY.md 8c names for the stage the reference itself does
not ship, F2), on seeded inputs at the two BASELINE clip shapes and one odd shape.

    python oracle/make_frontend_golden.py        (needs torchaudio; run in the build container, not on the GPU box)

Stored per case: the input seed/shape, torchaudio's own fp32 filterbank and window (so that the kernel under test
receives exactly the tables torchaudio used -- its filterbank is built in fp32 and differs from this repo's
float64-then-rounded one by up to 5e-4 dB), and dB = 10 log10(mel + eps) [B, n_mels, T] in fp32.  Also recorded: how far
torchaudio's fp32 pipeline itself sits from a float64 evaluation of the same formula with the same tables -- that
distance is the floor of any tolerance against this golden."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend as fe  # noqa: E402

CASES = {
    "s5": dict(seed=901, B=2, L=80000, hop=512, n_mels=128),
    "s1": dict(seed=902, B=2, L=16000, hop=102, n_mels=128),
    "odd": dict(seed=903, B=1, L=4099, hop=511, n_mels=80),
}


def waveform(c):
    rng = np.random.default_rng(c["seed"])
    t = np.arange(c["L"]) / 16000.0
    x = rng.standard_normal((c["B"], c["L"])) * 0.1 + 0.05 * np.sin(2 * np.pi * 440.0 * t)[None, :]
    return x.astype(np.float32)


def main():
    import torch
    import torchaudio

    out = {}
    for name, c in CASES.items():
        x = waveform(c)
        ms = torchaudio.transforms.MelSpectrogram(16000, n_fft=1024, hop_length=c["hop"], n_mels=c["n_mels"], f_min=0.0,
                                                  f_max=8000.0, power=2.0, norm="slaney", mel_scale="slaney",
                                                  center=True, pad_mode="reflect")
        with torch.no_grad():
            mel = ms(torch.from_numpy(x))  # [B, n_mels, T]
        db = (10.0 * torch.log10(mel + fe.LOG_EPS)).numpy().astype(np.float32)
        fb = ms.mel_scale.fb.numpy().astype(np.float32)          # [513, n_mels], torchaudio's fp32 table
        win = ms.spectrogram.window.numpy().astype(np.float32)   # torch.hann_window(1024), periodic
        ref = fe.logmel_f64(x, hop=c["hop"], n_mels=c["n_mels"], fb=fb, window=win)[:, 0]
        floor = float(np.abs(db - ref).max())
        print("%s: dB %s, torchaudio fp32 vs float64 of the same formula and tables: %.2e dB" % (name, db.shape, floor))
        out[name + "_db"] = db
        out[name + "_fb"] = fb
        out[name + "_window"] = win
        out[name + "_meta"] = np.asarray([c["seed"], c["B"], c["L"], c["hop"], c["n_mels"]], dtype=np.int64)
        out[name + "_fp32_floor_db"] = np.float64(floor)
    out["torchaudio_version"] = np.asarray(torchaudio.__version__)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "logmel_torchaudio.npz"), **out)


if __name__ == "__main__":
    main()
