#!/usr/bin/env python
"""Headline benchmark: episodes/sec, 5-way 5-shot 15-query, waveform -> logits.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json `metric`; SURVEY.md 8d, config C1 = config/proto_5shot_iid.yaml): ProtoNet on
Conv64F, 5w5s15q = 100 clips per episode, each clip 5 s @ 16 kHz (L = 80 000) -> log-mel [1,128,157]
(n_fft 1024, hop 512, 128 slaney mels, KOS_0.5_alpha mean/std).  One "step" = one pass of the hot path
over `--episodes-per-step` episodes per rank:
    fused log-mel kernel -> Conv64F (tcgen05 block-1 and block-2/3 kernels, cuDNN block 4) -> prototype head
    kernel -> vote/accuracy kernel.
Episodes are independent, so ranks never exchange data inside a step ("weak" scaling: per-GPU work is
fixed).  Synthetic seeded waveforms, weights derived from parameter names (oracle.cases.perturb_bn_).

`value`  : device-resident inputs (two rotating batches, each larger than L2).
`e2e`    : the public call EpisodePipeline.stream(pinned host batches) -- every step's H2D of the waveforms
           and D2H of the logits + accuracy inside the timed region (copies overlap compute).
`e2e_pcm16`: the same call with the host waveforms as 16-bit PCM (extra key; `e2e` is the fp32-host figure).
`roofline`: the fused log-mel kernel (our dominant kernel), algorithmic bytes 4*L + 4*128*T per clip over
           its CUDA-event duration measured inside the timed region, against MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the oracle port of the reference path (torch.stft front-end spec ->
           reference Conv64F arithmetic -> ProtoLayer -> majority vote) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, S, Q = 5, 5, 15
L, HOP, N_MELS, SR = 80000, 512, 128, 16000
T_FRAMES = 1 + L // HOP
CLIPS_PER_EPISODE = W * (S + Q)
LOGMEL_BYTES_PER_CLIP = 4 * L + 4 * N_MELS * T_FRAMES  # 400 384
MEAN_STD_FILE = os.path.join(ROOT, "tests", "golden", "KOS_0.5_alpha_Mean_Std.npy")
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent
METRIC = "episodes/sec (5w5s15q, waveform->logits)"
WORKLOAD = "ProtoNet Conv64F 5w5s15q, 100 clips/episode, 5 s @ 16 kHz -> log-mel [1,128,157] (C1, shape S5)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--episodes-per-step", type=int, default=32, help="episodes per rank per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="bound of the cpu_baseline sample")
    return ap.parse_args()


def mean_std():
    m, s = np.load(MEAN_STD_FILE).flatten().tolist()
    return float(m), float(s)


def make_weights():
    """Conv64F with deterministic name-derived weights; returns the product module (CPU)."""
    import torch

    from audio_fewshot_b200 import model as arch
    from oracle import cases

    torch.manual_seed(0)
    emb = arch.Conv64F(is_flatten=True, num_channels=1)
    cases.perturb_bn_(emb)
    return emb


# ----------------------------------------------------------------------------------- CPU oracle arm
class CpuReferencePath:
    """The reference's op sequence on host cores (oracle port; /root/reference cannot travel to the box)."""

    def __init__(self):
        import torch

        from oracle import frontend as ofe

        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = {k: v.clone() for k, v in make_weights().state_dict().items()}
        self.mean, self.std = mean_std()
        self.fb = ofe.mel_filterbank(513, 0.0, SR / 2.0, N_MELS, SR)
        self.win = ofe.hann_periodic()

    def episodes(self, wav, n_episodes):
        from oracle import backbones as obb
        from oracle import frontend as ofe
        from oracle import heads

        torch = self.torch
        with torch.no_grad():
            image = ofe.logmel_torch(wav, hop=HOP, n_mels=N_MELS, sample_rate=SR, mean=self.mean, std=self.std,
                                     fb=self.fb, window=self.win)
            feat = obb.conv64f_forward(self.sd, image)
            repeats = torch.ones(n_episodes * W * Q, dtype=torch.long)
            out, acc, _ = heads.proto_forward(feat, W, S, Q, repeats, n_episodes * W * S)
        return out, acc


def cpu_sample(seconds, max_episodes=24):
    """Time the oracle path one episode at a time until `seconds` of CPU work are done."""
    from oracle import cases

    path = CpuReferencePath()
    wav = cases.synthetic_clip_batch(1234, 0, 1, W, S, Q, L)
    path.episodes(wav, 1)  # warm-up (thread pools, FFT plans)
    n, t0 = 0, time.perf_counter()
    while n < max_episodes and (n < 2 or time.perf_counter() - t0 < seconds):
        path.episodes(wav, 1)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "episodes/sec", "cores": path.cores, "kind": "port",
            "sample": "%d episodes of the same workload, one at a time, %.1f s on %d host threads "
                      "(oracle: torch.stft log-mel -> Conv64F -> ProtoLayer -> vote)" % (n, dt, path.cores)}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path, rank 0 only."""
    if rank != 0:
        return
    from oracle import cases

    path = CpuReferencePath()
    e_ref = 1  # bounded sample: one episode per step
    wav = [cases.synthetic_clip_batch(1234, i, e_ref, W, S, Q, L) for i in range(2)]
    for i in range(args.warmup):
        path.episodes(wav[i % 2], e_ref)
    t0 = time.perf_counter()
    for i in range(args.steps):
        path.episodes(wav[i % 2], e_ref)
    dt = time.perf_counter() - t0
    value = args.steps * e_ref / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "episodes/sec", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "episodes_per_step": e_ref, "note": "CPU oracle port of the reference path; "
                   "runs on rank 0's host cores only, whatever --gpus says"},
        "cpu_baseline": {"value": value, "unit": "episodes/sec", "cores": path.cores, "kind": "port",
                         "sample": "%d steps x %d episode on %d host threads" % (args.steps, e_ref, path.cores)},
        "e2e": {"value": value, "unit": "episodes/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="afs_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for row in open(self.path):
                f = [c.strip() for c in row.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    smax.append(float(f[1]))
                except ValueError:
                    continue
                for name, flag in zip(names, f[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ----------------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    from audio_fewshot_b200 import model as arch
    from audio_fewshot_b200 import ops
    from audio_fewshot_b200.frontend import LogMelFrontEnd
    from audio_fewshot_b200.pipeline import EpisodePipeline
    from oracle import cases  # synthetic inputs + weight recipe only; nothing of oracle/ is timed here

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    E = args.episodes_per_step
    n_clips = E * CLIPS_PER_EPISODE
    mean, std = mean_std()

    emb = make_weights()
    model = arch.ProtoNet(way_num=W, shot_num=S, query_num=Q, test_way=W, test_shot=S, test_query=Q,
                          emb_func=emb, device=dev).to(dev).eval()
    front = LogMelFrontEnd(sample_rate=SR, hop_length=HOP, n_mels=N_MELS, mean=mean, std=std).to(dev).eval()
    pipe = EpisodePipeline(front, model)
    repeats = torch.ones(E * W * Q, dtype=torch.long)
    support_size = E * W * S

    # two rotating batches per rank; content keyed by the GLOBAL episode index
    host = []
    for b in range(2):
        first = (b * world + rank) * E
        host.append(torch.from_numpy(cases.synthetic_clip_batch(1234, first, E, W, S, Q, L)).pin_memory())
    devb = [h.to(dev) for h in host]
    wav_bytes = n_clips * L * 4
    # the same batches as 16-bit PCM (the wav-file sample format): second end-to-end leg, half the PCIe bytes
    host_pcm = [(h * 32768.0).round_().clamp_(-32768, 32767).to(torch.int16).pin_memory() for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_device(i, ev=None):
        wav = devb[i % 2]
        if ev is not None:
            ev[0].record()
        image = front(wav, first_clip_index=0)
        if ev is not None:
            ev[1].record()
        return model.set_forward([image, None, repeats, support_size])

    def run_e2e(n, batches=host):
        """n steps through the public streaming call: every step's waveforms go pinned host -> device on the
        copy stream, its logits and accuracy come back to pinned host memory; copies overlap compute."""
        last = None
        for last in pipe.stream((batches[i % 2] for i in range(n)), repeats, support_size):
            pass
        return last

    with torch.no_grad():
        # ---- device-resident: `value`
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for i in range(args.warmup):
            step_device(i)
        barrier()
        launches0 = ops.launch_count()
        lm_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            output, acc = step_device(i, lm_events[i])
        t1.record()
        barrier()
        ms_dev = max_over_ranks(t0.elapsed_time(t1))
        launches = ops.launch_count() - launches0
        logmel_ms = float(np.mean([a.elapsed_time(b) for a, b in lm_events]))
        acc_dev = float(acc.item())

        # ---- end to end through the public call: `e2e`
        run_e2e(args.warmup)
        barrier()
        t0.record()
        out_host, acc_host = run_e2e(args.steps)
        t1.record()
        barrier()
        ms_e2e = max_over_ranks(t0.elapsed_time(t1))
        clocks = sampler.stop() if rank == 0 else None
        acc_e2e = float(acc_host.item())

        # ---- the same call fed with int16 PCM host buffers: `e2e_pcm16` (extra key, not the headline)
        run_e2e(args.warmup, host_pcm)
        barrier()
        t0.record()
        _, acc_pcm_host = run_e2e(args.steps, host_pcm)
        t1.record()
        barrier()
        ms_pcm = max_over_ranks(t0.elapsed_time(t1))
        acc_pcm = float(acc_pcm_host.item())

    total_eps = args.steps * E * world
    if world > 1:
        lm = torch.tensor([logmel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(lm, op=dist.ReduceOp.MAX)
        logmel_ms = float(lm.item())
    if rank != 0:
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    achieved = LOGMEL_BYTES_PER_CLIP * n_clips / (logmel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "logmel_traffic.json")  # written from an ncu --set full capture
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))  # captured at `clips_per_launch` clips; DRAM bytes scale with the clip count
            traffic = tj["dram_bytes_per_launch"] / tj["clips_per_launch"] * n_clips
        except Exception:
            traffic = None

    line = {
        "metric": METRIC, "value": total_eps / (ms_dev * 1e-3), "unit": "episodes/sec", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "episodes_per_step_per_gpu": E, "clips_per_step_per_gpu": n_clips,
                   "way": W, "shot": S, "query": Q, "clip_samples": L, "n_fft": 1024, "hop": HOP, "n_mels": N_MELS,
                   "backbone": "Conv64F eval path: tcgen05 TF32 kernels for block 1 (conv+BN+ReLU+pool) and blocks 2-3 "
                               "(implicit GEMM+BN+ReLU+pool), cuDNN conv+bias+ReLU for block 4 (TF32 allowed, the "
                               "reference's PyTorch default)",
                   "e2e_path": "EpisodePipeline.stream: H2D on a copy stream overlapped with compute, 2 buffers",
                   "l2_policy": "inputs larger than L2: %d MB of waveform per step, two rotating batches"
                                % (wav_bytes // 2 ** 20),
                   "parallelism": "episodes sharded over %d rank(s), no data-path collective" % world,
                   "accuracy_pct": acc_dev},
        "e2e": {"value": total_eps / (ms_e2e * 1e-3), "unit": "episodes/sec", "h2d_bytes_per_step": wav_bytes,
                "d2h_bytes_per_step": out_host.numel() * 4 + 4, "ms_per_step": ms_e2e / args.steps,
                "accuracy_pct": acc_e2e},
        "e2e_pcm16": {"value": total_eps / (ms_pcm * 1e-3), "unit": "episodes/sec", "h2d_bytes_per_step": wav_bytes // 2,
                      "d2h_bytes_per_step": out_host.numel() * 4 + 4, "ms_per_step": ms_pcm / args.steps,
                      "accuracy_pct": acc_pcm,
                      "note": "same public call, host waveforms quantised to int16 PCM (afs_logmel_fwd_pcm16 converts "
                              "on load); `e2e` above is the fp32-host-buffer figure"},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "logmel_kernel<false> (fused waveform->log-mel)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_launch": LOGMEL_BYTES_PER_CLIP * n_clips,
                     "ms_per_launch": logmel_ms, "share_of_step": logmel_ms / (ms_dev / args.steps)},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_sample(args.cpu_seconds)
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """Libraries (NCCL's version banner, cuDNN warnings) print to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    args = parse()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
