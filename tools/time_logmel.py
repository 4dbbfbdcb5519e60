"""Median time of the fused log-mel kernel at the bench shape, L2 flushed between launches (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200.frontend import LogMelFrontEnd
dev = torch.device("cuda", 0)
fr = LogMelFrontEnd(hop_length=512, n_mels=128, mean=-15.0, std=26.0).to(dev).eval()
for B in (800, 3200):
    wav = torch.randn(B, 80000, device=dev) * 0.1
    out = torch.empty(B, 1, 128, 157, device=dev)
    flush = torch.empty(40 * 2 ** 20, device=dev)
    for _ in range(3):
        fr(wav, out=out)
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fr(wav, out=out); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("logmel B=%d: median %.4f ms, min %.4f ms, %.0f GB/s" % (B, ts[10], ts[0], B * 400384 / ts[10] / 1e6))
