#!/usr/bin/env python
"""Per-kernel roofline table at the BASELINE configs: each head / front-end kernel timed
alone with CUDA events, inputs larger than L2 (or L2 flushed), algorithmic bytes / duration against the
measured HBM peak.  Prints one JSON line per kernel; `python bench_kernels.py > profiles/rNN_kernels.jsonl`.
bench.py is the contract benchmark; this is the supporting evidence for DESIGN.md section 3."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from audio_fewshot_b200 import ops  # noqa: E402
from audio_fewshot_b200.episode import EpisodeTable  # noqa: E402
from audio_fewshot_b200.frontend import LogMelFrontEnd  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, iters=20, warmup=3, flush=None):
    if os.environ.get("AFS_BENCH_ONCE"):  # under ncu: one warm launch + one profiled launch per kernel
        iters, warmup = 1, 1
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()  # > L2: evicts the previous iteration's lines
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times)), float(np.min(times))


def report(name, unit, units, bytes_per_unit, ms, ms_min, flops_per_unit=None, note=""):
    peak = peak_gbs()
    gbs = units * bytes_per_unit / (ms * 1e-3) / 1e9
    line = {"kernel": name, "unit": unit, "units_per_launch": units, "algorithmic_bytes_per_unit": bytes_per_unit,
            "ms_median": ms, "ms_min": ms_min, "achieved_gbs": gbs, "hbm_peak_gbs": peak, "frac_of_measured_hbm": gbs / peak,
            "units_per_s": units / (ms * 1e-3), "note": note}
    if flops_per_unit:
        line["tflops"] = units * flops_per_unit / (ms * 1e-3) / 1e12
    print(json.dumps(line), flush=True)


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    flush = torch.empty(160 * 2 ** 20 // 4, dtype=torch.float32, device=dev)  # 160 MB > 126 MB L2

    # ---- fused log-mel, shapes S5 and S1
    for tag, L, hop, B in (("S5 L=80000 hop=512", 80000, 512, 800), ("S5 L=80000 hop=512", 80000, 512, 3200),
                           ("S1 L=16000 hop=102", 16000, 102, 800)):
        for engine, kname in (("pair", "logmel_pair_kernel<false> (warp-per-frame-pair engine, default)"),
                              ("fft", "logmel_kernel<false> (radix-8 engine)"), ("tc", "logmel_tc_kernel<false> (tcgen05 DFT engine)")):
            fr = LogMelFrontEnd(hop_length=hop, n_mels=128, mean=-15.0, std=26.0, engine=engine).to(dev).eval()
            wav = torch.randn(B, L, device=dev) * 0.1
            T = 1 + L // hop
            out = torch.empty(B, 1, 128, T, device=dev)
            ms, mn = timeit(lambda: fr(wav, out=out), flush=flush)
            report("%s %s B=%d" % (kname, tag, B), "clip", B, 4 * L + 4 * 128 * T, ms, mn, note="T=%d" % T)
            del wav, out
    fr = LogMelFrontEnd(hop_length=512, n_mels=128, mean=-15.0, std=26.0, seed=7,
                        aug={"gain_db": (-6.0, 6.0), "max_shift": 1600, "noise_std": (0.0, 0.02)}).to(dev).train()
    wav = torch.randn(800, 80000, device=dev) * 0.1
    out = torch.empty(800, 1, 128, 157, device=dev)
    ms, mn = timeit(lambda: fr(wav, out=out), flush=flush)
    report("logmel_kernel<true> S5 + Philox gain/shift/noise", "clip", 800, 4 * 80000 + 4 * 128 * 157, ms, mn)

    # ---- conv1 stem
    img = torch.randn(800, 1, 128, 157, device=dev)
    w = np.random.default_rng(0).standard_normal((64, 9)).astype(np.float32)
    b = np.zeros(64, np.float32)
    ms, mn = timeit(lambda: ops.conv1_bn_act_pool3(img, w, b, 0.0), flush=flush)
    report("conv1_bn_act_pool3_kernel", "clip", 800, 4 * 128 * 157 + 4 * 64 * 42 * 52, ms, mn,
           flops_per_unit=2 * 64 * 42 * 52 * 81, note="fp32 SIMT; flops count the 9 conv positions per pooled pixel")
    ms, mn = timeit(lambda: ops.conv1_bn_act_pool3(img, w, b, 0.0, tf32=True), flush=flush)
    report("conv1_tc_kernel (tcgen05 TF32)", "clip", 800, 4 * 128 * 157 + 4 * 64 * 42 * 52, ms, mn,
           flops_per_unit=2 * 64 * 42 * 52 * 81, note="useful flops; the MMAs pad K 9 -> 16")
    ms, mn = timeit(lambda: ops.conv1_bn_act_pool3(img, w, b, 0.0, tf32=True, out_dtype=torch.bfloat16), flush=flush)
    report("conv1_tc_kernel, bf16 output (stated separately)", "clip", 800, 4 * 128 * 157 + 2 * 64 * 42 * 52, ms, mn,
           flops_per_unit=2 * 64 * 42 * 52 * 81, note="same TF32 arithmetic, rounded to bf16 at the store")
    act = torch.randn(800, 64, 42, 52, device=dev).contiguous(memory_format=torch.channels_last)
    ms, mn = timeit(lambda: ops.maxpool3_channels_last(act), flush=flush)
    report("maxpool3_nhwc_kernel [800,64,42,52]", "clip", 800, 4 * 64 * (42 * 52 + 14 * 17), ms, mn)

    # ---- Conv64F blocks 2 and 3 on the tensor cores (conv + folded BN + ReLU + max-pool) vs cuDNN + our pool
    for tag, H, Wd in (("block 2 [800,64,42,52]", 42, 52), ("block 3 [800,64,14,17]", 14, 17)):
        act = torch.randn(800, 64, H, Wd, device=dev).contiguous(memory_format=torch.channels_last)
        wt = (torch.randn(64, 64, 3, 3, device=dev) * 0.06)
        wcl = wt.contiguous(memory_format=torch.channels_last)
        bias = torch.randn(64, device=dev)
        packed = torch.from_numpy(ops.conv3x3_c64_pack_weights(wt)).to(dev)
        flops = 2 * 64 * 64 * 9 * H * Wd
        nbytes = 4 * 64 * (H * Wd + (H // 3) * (Wd // 3))
        ms, mn = timeit(lambda: ops.conv3x3_c64_bn_act(act, packed, bias, 0.0, pool=True), flush=flush)
        report("conv3x3_c64_tc_kernel " + tag, "clip", 800, nbytes, ms, mn, flops_per_unit=flops,
               note="tcgen05 TF32, pool fused; useful flops")
        act16 = act.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        packed16 = torch.from_numpy(ops.conv3x3_c64_pack_weights_bf16(wt)).to(dev).view(torch.bfloat16)
        ms, mn = timeit(lambda: ops.conv3x3_c64_bn_act_bf16(act16, packed16, bias, 0.0, pool=True), flush=flush)
        report("conv3x3_c64_tc_kernel bf16 (stated separately) " + tag, "clip", 800, nbytes // 2, ms, mn, flops_per_unit=flops,
               note="tcgen05 kind::f16 bf16 operands, K = 16, two epilogue groups, bf16 in / bf16 out; useful flops")
        torch.backends.cudnn.allow_tf32 = True
        ms, mn = timeit(lambda: ops.maxpool3_channels_last(
            torch.cudnn_convolution_relu(act, wcl, bias, (1, 1), (1, 1), (1, 1), 1)), flush=flush)
        report("cuDNN conv+bias+ReLU (TF32) + maxpool3_nhwc " + tag, "clip", 800, nbytes, ms, mn, flops_per_unit=flops,
               note="library baseline for the kernel above")

    # ---- prototype head: C1 (D=1600, 5w5s15q), C2 (D=12800, 5w1s15q), C4 vectors (D=2080, 5w5s10q)
    for tag, E, W, S, Q, D, mode in (("C1 D=1600 5w5s15q", 256, 5, 5, 15, 1600, "euclidean"),
                                     ("C1 D=1600 5w5s15q", 2048, 5, 5, 15, 1600, "euclidean"),
                                     ("C2 D=12800 5w1s15q", 64, 5, 1, 15, 12800, "euclidean"),
                                     ("C2 D=12800 5w1s15q", 512, 5, 1, 15, 12800, "euclidean"),
                                     ("C4 D=2080 5w5s10q", 256, 5, 5, 10, 2080, "euclidean"),
                                     ("C1 cosine", 256, 5, 5, 15, 1600, "cos_sim")):
        N = E * W * (S + Q)
        feat = torch.randn(N, D, device=dev)
        tab = EpisodeTable(E, W, S, Q, np.ones(E * W * Q, dtype=np.int64), dev)
        ms, mn = timeit(lambda: ops.proto_logits(feat, tab.cls_row, E, W, S, mode), flush=flush)
        report("proto_mean_kernel + proto_fwd_kernel %s E=%d" % (tag, E), "episode", E, 4 * W * (S + Q) * D + 4 * W * Q * W, ms, mn,
               flops_per_unit=3 * W * Q * W * D + W * S * D, note="two launches; small batches are launch-bound")
        if mode == "euclidean" and D % 32 == 0:
            ms, mn = timeit(lambda: ops.proto_logits(feat, tab.cls_row, E, W, S, mode, precision="tf32"), flush=flush)
            report("proto_tc_kernel (TMA + tcgen05 TF32 GEMM epilogue) %s E=%d" % (tag, E), "episode", E,
                   4 * W * (S + Q) * D + 4 * W * Q * W, ms, mn, flops_per_unit=2 * W * (S + Q) * 32 * D,
                   note="precision class tf32; flops = executed MMA flops (N = 32 columns)")
        del feat

    # ---- DN4: C3 Conv64F maps [64,4,5], 5w5s15q n_k=3; ResNet-12 maps [640,8,9] 5w5s10q
    for tag, E, W, S, Q, C, H, Wd in (("C3 map 64x4x5 5w5s15q", 128, 5, 5, 15, 64, 4, 5),
                                      ("ResNet-12 map 640x8x9 5w5s10q", 4, 5, 5, 10, 640, 8, 9)):
        N = E * W * (S + Q)
        feat = torch.rand(N, C, H, Wd, device=dev)
        tab = EpisodeTable(E, W, S, Q, np.ones(E * W * Q, dtype=np.int64), dev)
        ms, mn = timeit(lambda: ops.dn4_scores(feat, tab.cls_row, E, W, S, 3), flush=flush)
        HW = H * Wd
        report("dn4 fp32 (normalize+main+reduce) " + tag, "episode", E, 4 * W * (S + Q) * C * HW + 4 * W * Q * W, ms, mn,
               flops_per_unit=2 * (W * Q * HW) * (W * S * HW) * C)
        if C % 32 == 0:
            ms, mn = timeit(lambda: ops.dn4_scores(feat, tab.cls_row, E, W, S, 3, precision="tf32"), flush=flush)
            report("dn4 tf32 TMA+tcgen05 (prep+main+reduce) " + tag, "episode", E,
                   4 * W * (S + Q) * C * HW + 4 * W * Q * W, ms, mn, flops_per_unit=2 * (W * Q * HW) * (W * S * HW) * C)

    # ---- BDC matrix: C4 map [64,16,19]
    x = torch.relu(torch.randn(2000, 64, 16, 19, device=dev))
    t = torch.full((1, 1), float(np.log(1 / 200.0)), device=dev)
    for B_ in (2000, 8000):
        x = torch.relu(torch.randn(B_, 64, 16, 19, device=dev))
        for tc in (False, True):
            ops.bdc_set_tensor_core(tc)
            ms, mn = timeit(lambda: ops.bdc_pool(x, t), flush=flush)
            report("%s C4 map 64x16x19 B=%d" % ("bdc_tc_kernel (tcgen05, 3xTF32)" if tc else "bdc_kernel (fp32 FMA)", B_), "clip", B_,
                   4 * 64 * 304 + 4 * 2080, ms, mn, flops_per_unit=2 * 64 * 64 * 304)
        ops.bdc_set_tensor_core(True)

    # ---- vote + accuracy, 5-way, one window per query
    nq = 75 * 4096
    logits = torch.randn(nq, 5, device=dev)
    q_start = torch.arange(nq + 1, dtype=torch.int32, device=dev)
    target = torch.randint(0, 5, (nq,), dtype=torch.int32, device=dev)
    ms, mn = timeit(lambda: ops.vote_acc(logits, q_start, target), flush=flush)
    report("vote_kernel 5-way", "episode", 4096, 4 * 75 * 5 + 4 * 75 * 3, ms, mn, note="logits + q_start + target + pred")


def extra():
    """Kernels outside the evaluation step: PCM16 front-end, spectrogram augmentation, ResNet-12 block tail, and the
    fused training block 1."""
    import random

    from audio_fewshot_b200 import augment as aug
    dev = torch.device("cuda", 0)
    flush = torch.empty(160 * 2 ** 20 // 4, dtype=torch.float32, device=dev)
    fr = LogMelFrontEnd(hop_length=512, n_mels=128, mean=-15.0, std=26.0).to(dev).eval()
    pcm = (torch.randn(800, 80000, device=dev) * 0.1 * 32768).round().clamp(-32768, 32767).to(torch.int16)
    out = torch.empty(800, 1, 128, 157, device=dev)
    ms, mn = timeit(lambda: fr(pcm, out=out), flush=flush)
    report("logmel_kernel<false, int16> S5 (16-bit PCM input)", "clip", 800, 2 * 80000 + 4 * 128 * 157, ms, mn)

    img = torch.randn(800, 1, 128, 157, device=dev)
    for t in ("noise_suppression", "background_subtraction", "cutout"):
        random.seed(1)
        ms, mn = timeit(lambda: aug.augment_spectrogram(img, -15.0, 26.0, augmentation_type=t), flush=flush)
        report("spec_augment_kernel " + t + " [800,1,128,157]", "clip", 800, 2 * 4 * 128 * 157, ms, mn)

    a = torch.randn(320, 64, 128, 157, device=dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn_like(a).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(64, device=dev)
    ms, mn = timeit(lambda: ops.add_bias_act_pool(a, b, bias, 0.1, 2), flush=flush)
    report("add_bias_act_pool_nhwc_kernel k=2 [320,64,128,157] (ResNet-12 layer-1 tail)", "clip", 320,
           4 * 64 * 128 * 157 * 2 + 4 * 64 * 64 * 78, ms, mn)

    conv = torch.nn.Conv2d(1, 64, 3, padding=1).to(dev)
    bn = torch.nn.BatchNorm2d(64).to(dev).train()
    x = torch.randn(200, 1, 128, 157, device=dev)
    g = torch.randn(200, 64, 42, 52, device=dev).contiguous(memory_format=torch.channels_last)
    y = ops.conv1_bn_act_pool3_train(x, conv, bn, 0.0)
    ms, mn = timeit(lambda: ops.conv1_bn_act_pool3_train(x, conv, bn, 0.0), flush=flush)
    report("conv1 training block forward (autocorr + fwd kernels + batch-stat algebra) [200,1,128,157]", "clip", 200,
           4 * 128 * 157 * 2 + 4 * 64 * 42 * 52, ms, mn, flops_per_unit=2 * 64 * 42 * 52 * 81)

    def bwd():
        y = ops.conv1_bn_act_pool3_train(x, conv, bn, 0.0)
        y.backward(g)

    ms_fb, mn_fb = timeit(bwd, flush=flush)
    report("conv1 training block forward + backward [200,1,128,157]", "clip", 200,
           4 * 128 * 157 * 3 + 2 * 4 * 64 * 42 * 52, ms_fb, mn_fb, flops_per_unit=3 * 2 * 64 * 42 * 52 * 81)


if __name__ == "__main__":
    main()
    extra()
