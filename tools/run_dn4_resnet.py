"""DN4 head on ResNet-12 maps [640, 8, 9], 5w5s10q: fp32 path vs the K-streaming tcgen05 schedule (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
from audio_fewshot_b200.episode import EpisodeTable
dev = torch.device("cuda", 0)
for E in (4, 16):
    W, S, Q, C, H, Wd = 5, 5, 10, 640, 8, 9
    N = E * W * (S + Q)
    feat = torch.rand(N, C, H, Wd, device=dev)
    tab = EpisodeTable(E, W, S, Q, np.ones(E * W * Q, dtype=np.int64), dev)
    res = {}
    for prec in ("fp32", "tf32"):
        for _ in range(2):
            out = ops.dn4_scores(feat, tab.cls_row, E, W, S, 3, precision=prec)[0]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            out = ops.dn4_scores(feat, tab.cls_row, E, W, S, 3, precision=prec)[0]
        t1.record(); torch.cuda.synchronize()
        res[prec] = (t0.elapsed_time(t1) / 5, out)
    flops = 2.0 * (W * Q * H * Wd) * (W * S * H * Wd) * C * E
    err = (res["fp32"][1] - res["tf32"][1]).abs().max().item() / res["fp32"][1].abs().max().item()
    print("E=%d  fp32 %.3f ms (%.0f TFLOP/s)   tf32 %.3f ms (%.0f TFLOP/s)   rel diff %.1e" % (
        E, res["fp32"][0], flops / res["fp32"][0] / 1e9, res["tf32"][0], flops / res["tf32"][0] / 1e9, err), flush=True)
