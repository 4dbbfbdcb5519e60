"""Episode loaders producing the batch tuple the hot path consumes.

The reference's `libfewshot_core.data` package (get_dataloader, samplers, collates, get_mean_std) is
not in the snapshot; what it must deliver is fixed by its call sites:

  * `get_dataloader(config, mode, model_type, distribute[, modality])` returns a SEQUENCE of loaders;
    callers do `zip(*loaders)`, `len(loaders[0])`, `loaders[0].sampler.set_epoch(e)`
    (trainer.py:88,159,163,265; test.py:362,367);
  * each item flattens to `[image, global_target, repeats, support_size]` (test.py:392-393):
    `image` [rows, 1, 128, 157] fp32 with rows episode-major, class-major, S support rows then the
    query windows (abstract_model.py:215-252); `repeats` [E*W*Q] = windows per query (variable-length
    clips are chopped into fixed windows and majority-voted, utils.py:436-446); `support_size` = E*W*S;
  * statistics come from `mean_std_file` = (2,1,1) [mean, std] (test.py:398-399).

Four sources share one sampler:
  SyntheticWaveformEpisodes   seeded class-tone waveforms [rows, L] for the fused log-mel front-end
  SyntheticSpectrogramEpisodes  seeded [rows,1,128,157] images (front-end already applied / bypassed)
  SpectrogramFolderEpisodes   `<root>/<class>/*.npy` log-mel arrays [128, T] (the `*_spec` folders of
                              config/headers/data.yaml:1) with a class split (Auxiliary/KOS_paper_splits.npy)

  WaveformFolderEpisodes      `<root>/<class>/*.wav` 16-bit PCM clips, served as int16 [rows, L] windows for
                              afs_logmel_fwd_pcm16 (the step the reference did offline to build its `*_spec`
                              folders), with the same support/query window protocol

Everything random is keyed by the GLOBAL episode index, so a run is invariant to the world size.
"""
import os

import numpy as np
import torch

from .dist import shard_episodes, world

N_MELS, N_FRAMES = 128, 157


class EpisodeSampler:
    """Global episode-batch indices of this rank for one epoch: batch b covers episodes
    [b*episode_size, (b+1)*episode_size); ranks take batches round-robin.  `episode_size` is per rank (the global
    batch is world_size * episode_size) and every rank iterates the same number of batches."""

    def __init__(self, n_episodes, episode_size, rank=0, world_size=1, seed=0):
        if n_episodes % episode_size:
            raise ValueError("episodes %d %% episode_size %d != 0 (trainer.py:724-754)" % (n_episodes, episode_size))
        self.n_batches = n_episodes // episode_size
        if world_size > 1 and self.n_batches % world_size:
            # every step ends in collectives (accuracy all-reduce, gradient all-reduce): ranks with different batch
            # counts would deadlock on the last one.  The reference gets equal counts by splitting episode_size over
            # the GPUs (trainer.py:724-754 asserts episode_size % n_gpu == 0); here episode_size is PER RANK and the
            # global batch is world_size * episode_size, so the batch count must divide.
            raise ValueError("%d episode batches do not divide over %d ranks: choose episodes as a multiple of "
                             "episode_size * world_size (= %d)" % (self.n_batches, world_size, episode_size * world_size))
        self.episode_size, self.rank, self.world_size, self.seed = episode_size, rank, world_size, seed
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def __len__(self):
        return len(shard_episodes(self.n_batches, self.rank, self.world_size))

    def __iter__(self):
        for b in shard_episodes(self.n_batches, self.rank, self.world_size):
            yield b

    def rng(self, batch_index, episode_in_batch):
        g = batch_index * self.episode_size + episode_in_batch
        return np.random.default_rng((self.seed, self.epoch, g))


class _EpisodeLoader:
    def __init__(self, sampler, way, shot, query, pin=False):
        self.sampler, self.way, self.shot, self.query, self.pin = sampler, way, shot, query, pin

    def __len__(self):
        return len(self.sampler)

    def _finish(self, image, target, repeats):
        image = torch.from_numpy(image)
        if self.pin and torch.cuda.is_available():
            image = image.pin_memory()
        E = self.sampler.episode_size
        return (image, torch.from_numpy(target.astype(np.int64)), torch.from_numpy(repeats.astype(np.int64)),
                E * self.way * self.shot)

    def __iter__(self):
        for b in self.sampler:
            yield self.batch(b)


class SyntheticWaveformEpisodes(_EpisodeLoader):
    """Class c of an episode is noise N(0, 0.1^2) plus the tone 0.05 sin(2 pi 200 (c+1) t)."""

    def __init__(self, sampler, way, shot, query, n_samples=80000, sample_rate=16000, pin=True):
        super().__init__(sampler, way, shot, query, pin)
        self.n_samples, self.sample_rate = n_samples, sample_rate

    def batch(self, b):
        E, W, P = self.sampler.episode_size, self.way, self.shot + self.query
        t = np.arange(self.n_samples, dtype=np.float64) / self.sample_rate
        out = np.empty((E, W, P, self.n_samples), dtype=np.float32)
        for e in range(E):
            r = self.sampler.rng(b, e)
            noise = r.standard_normal((W, P, self.n_samples)).astype(np.float32) * np.float32(0.1)
            for c in range(W):
                out[e, c] = noise[c] + (0.05 * np.sin(2.0 * np.pi * 200.0 * (c + 1) * t)).astype(np.float32)[None]
        target = np.tile(np.repeat(np.arange(W), P), E)
        return self._finish(out.reshape(E * W * P, self.n_samples), target, np.ones(E * W * self.query))


class SyntheticSpectrogramEpisodes(_EpisodeLoader):
    """Already-normalised [rows,1,128,157] images with a class-dependent band; variable window counts."""

    def __init__(self, sampler, way, shot, query, max_windows=1, pin=False):
        super().__init__(sampler, way, shot, query, pin)
        self.max_windows = max_windows

    def batch(self, b):
        E, W, S, Q = self.sampler.episode_size, self.way, self.shot, self.query
        rows, target, repeats = [], [], []
        for e in range(E):
            r = self.sampler.rng(b, e)
            for c in range(W):
                rep = r.integers(1, self.max_windows + 1, size=Q)
                n = S + int(rep.sum())
                x = r.standard_normal((n, 1, N_MELS, N_FRAMES)).astype(np.float32)
                x[:, :, 8 * c: 8 * c + 8, :] += 2.0
                rows.append(x)
                target.extend([c] * n)
                repeats.extend(rep.tolist())
        return self._finish(np.concatenate(rows, axis=0), np.asarray(target), np.asarray(repeats))


def window_spectrogram(spec, n_frames=N_FRAMES):
    """[n_mels, T] -> [k, n_mels, n_frames]: consecutive windows, the last one right-aligned (zero-padded
    when the clip is shorter than one window)."""
    n_mels, T = spec.shape
    if T <= n_frames:
        out = np.zeros((1, n_mels, n_frames), dtype=np.float32)
        out[0, :, :T] = spec
        return out
    k = -(-T // n_frames)
    starts = [min(i * n_frames, T - n_frames) for i in range(k)]
    return np.stack([spec[:, s:s + n_frames] for s in starts]).astype(np.float32)


class SpectrogramFolderEpisodes(_EpisodeLoader):
    """`root/<class>/*.npy` log-mel clips.  Supports use their first window; queries are chopped into
    all their windows and `repeats` records how many (the reference's variable-length protocol)."""

    def __init__(self, sampler, root, classes, way, shot, query, mean=0.0, std=1.0, pin=False):
        super().__init__(sampler, way, shot, query, pin)
        self.mean, self.std = float(mean), float(std)
        self.files = {}
        for c in classes:
            d = os.path.join(root, str(c))
            fs = sorted(f for f in os.listdir(d) if f.endswith(".npy")) if os.path.isdir(d) else []
            if len(fs) >= shot + query:
                self.files[str(c)] = [os.path.join(d, f) for f in fs]
        self.classes = sorted(self.files)
        if len(self.classes) < way:
            raise ValueError("only %d classes with >= %d clips under %s" % (len(self.classes), shot + query, root))
        self.class_id = {c: i for i, c in enumerate(self.classes)}

    def _load(self, path):
        spec = np.load(path).astype(np.float32)
        spec = spec.reshape(spec.shape[-2], spec.shape[-1])
        return (spec - self.mean) / self.std  # normalize_spectrogram, audio_augmentations.py:36-53

    def batch(self, b):
        E, W, S, Q = self.sampler.episode_size, self.way, self.shot, self.query
        rows, target, repeats = [], [], []
        for e in range(E):
            r = self.sampler.rng(b, e)
            for c in r.choice(len(self.classes), size=W, replace=False):
                name = self.classes[int(c)]
                picks = r.choice(len(self.files[name]), size=S + Q, replace=False)
                for i, p in enumerate(picks):
                    win = window_spectrogram(self._load(self.files[name][int(p)]))
                    if i < S:
                        win = win[:1]
                    else:
                        repeats.append(win.shape[0])
                    rows.append(win[:, None])
                    target.extend([self.class_id[name]] * win.shape[0])
        return self._finish(np.concatenate(rows, axis=0), np.asarray(target), np.asarray(repeats))


def read_wav_pcm16(path):
    """Mono int16 samples and the sample rate of a 16-bit PCM .wav file (multi-channel files: first channel)."""
    import wave

    with wave.open(path, "rb") as w:
        if w.getsampwidth() != 2:
            raise ValueError("%s: only 16-bit PCM wav files are supported (sample width %d)" % (path, w.getsampwidth()))
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        if w.getnchannels() > 1:
            pcm = pcm.reshape(-1, w.getnchannels())[:, 0]
        return np.ascontiguousarray(pcm), w.getframerate()


def window_waveform(pcm, n_samples):
    """[T] -> [k, n_samples]: consecutive windows, the last one right-aligned (zero-padded when the clip is
    shorter than one window) -- window_spectrogram's protocol in the sample domain."""
    T = pcm.shape[0]
    if T <= n_samples:
        out = np.zeros((1, n_samples), dtype=pcm.dtype)
        out[0, :T] = pcm
        return out
    k = -(-T // n_samples)
    starts = [min(i * n_samples, T - n_samples) for i in range(k)]
    return np.stack([pcm[s:s + n_samples] for s in starts])


class WaveformFolderEpisodes(_EpisodeLoader):
    """`root/<class>/*.wav` 16-bit PCM clips -> `image` = int16 waveform windows [rows, n_samples] for the fused
    log-mel front-end (LogMelFrontEnd accepts int16 and converts on load: pcm / 32768).  Supports use their first
    window; queries are chopped into all their windows and `repeats` records how many."""

    def __init__(self, sampler, root, classes, way, shot, query, n_samples=80000, sample_rate=16000, pin=True):
        super().__init__(sampler, way, shot, query, pin)
        self.n_samples, self.sample_rate = int(n_samples), int(sample_rate)
        self.files = {}
        for c in classes:
            d = os.path.join(root, str(c))
            fs = sorted(f for f in os.listdir(d) if f.lower().endswith(".wav")) if os.path.isdir(d) else []
            if len(fs) >= shot + query:
                self.files[str(c)] = [os.path.join(d, f) for f in fs]
        self.classes = sorted(self.files)
        if len(self.classes) < way:
            raise ValueError("only %d classes with >= %d clips under %s" % (len(self.classes), shot + query, root))
        self.class_id = {c: i for i, c in enumerate(self.classes)}

    def _load(self, path):
        pcm, sr = read_wav_pcm16(path)
        if sr != self.sample_rate:
            raise ValueError("%s: sample rate %d, expected %d (resample offline)" % (path, sr, self.sample_rate))
        return pcm

    def batch(self, b):
        E, W, S, Q = self.sampler.episode_size, self.way, self.shot, self.query
        rows, target, repeats = [], [], []
        for e in range(E):
            r = self.sampler.rng(b, e)
            for c in r.choice(len(self.classes), size=W, replace=False):
                name = self.classes[int(c)]
                picks = r.choice(len(self.files[name]), size=S + Q, replace=False)
                for i, p in enumerate(picks):
                    win = window_waveform(self._load(self.files[name][int(p)]), self.n_samples)
                    if i < S:
                        win = win[:1]
                    else:
                        repeats.append(win.shape[0])
                    rows.append(win)
                    target.extend([self.class_id[name]] * win.shape[0])
        return self._finish(np.concatenate(rows, axis=0), np.asarray(target), np.asarray(repeats))


def load_class_split(path, mode):
    """Auxiliary/KOS_paper_splits.npy: object array [train, val, test] of class-name lists."""
    arr = np.load(path, allow_pickle=True)
    return [str(c) for c in arr[{"train": 0, "val": 1, "test": 2}[mode]]]


def get_mean_std(config, mode="train", modality="audio"):
    """(mean, std) floats from `mean_std_file` (imported by the reference at test.py:31)."""
    from .frontend import load_mean_std
    return load_mean_std(config["mean_std_file"])


def get_dataloader(config, mode, model_type=None, distribute=False, modality="audio"):
    """Tuple of `dataloader_num` episode loaders for `mode` in {"train","val","test"}.  With a real
    `data_root` (+ `class_per_split`) it reads spectrogram folders; otherwise (`data_root` missing or
    "synthetic") it serves seeded synthetic episodes -- waveforms when `synthetic_waveform` is set."""
    rank, ws = world() if distribute else (0, 1)
    train = mode == "train"
    way = config["way_num"] if train else config["test_way"]
    shot = (config["shot_num"] if train else config["test_shot"]) * config.get("augment_times", 1)
    query = config["query_num"] if train else config["test_query"]
    n_eps = config["train_episode"] if train else config["test_episode"]
    loaders = []
    for i in range(int(config.get("dataloader_num", 1))):
        sampler = EpisodeSampler(n_eps, config.get("episode_size", 1), rank, ws,
                                 seed=int(config.get("seed", 0)) * 7 + {"train": 0, "val": 1, "test": 2}[mode] + 3 * i)
        root = config.get("data_root")
        if root and root != "synthetic" and os.path.isdir(root):
            classes = (load_class_split(config["class_per_split"], mode) if config.get("class_per_split")
                       else sorted(os.listdir(root)))
            mean, std = get_mean_std(config) if config.get("mean_std_file") else (0.0, 1.0)
            if config.get("data_format") == "wav":  # raw 16-bit clips: the log-mel front-end runs on the GPU
                loaders.append(WaveformFolderEpisodes(sampler, root, classes, way, shot, query,
                                                      n_samples=int(config.get("audio_samples", 80000)),
                                                      sample_rate=int(config.get("sample_rate", 16000))))
            else:
                loaders.append(SpectrogramFolderEpisodes(sampler, root, classes, way, shot, query, mean, std))
        elif config.get("synthetic_waveform"):
            loaders.append(SyntheticWaveformEpisodes(sampler, way, shot, query,
                                                     n_samples=int(config.get("audio_samples", 80000))))
        else:
            loaders.append(SyntheticSpectrogramEpisodes(sampler, way, shot, query,
                                                        max_windows=int(config.get("max_windows", 1))))
    return tuple(loaders)
