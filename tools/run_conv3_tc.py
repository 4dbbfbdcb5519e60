"""Quick GPU check + timing of csrc/conv3_tc.cu (development helper)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
for (N, H, Wd, pool) in [(2, 42, 52, True), (5, 14, 17, True), (3, 14, 17, False), (300, 42, 52, True), (800, 42, 52, True)]:
    torch.manual_seed(1)
    x = torch.randn(N, 64, H, Wd, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 64, 3, 3, device=dev) * 0.06
    b = torch.randn(64, device=dev)
    packed = torch.from_numpy(ops.conv3x3_c64_pack_weights(w)).to(dev)
    got = ops.conv3x3_c64_bn_act(x, packed, b, 0.0, pool=pool)
    torch.cuda.synchronize()
    want = torch.relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1))
    if pool:
        want = torch.nn.functional.max_pool2d(want, 3, 3)
    err = (got.double() - want).abs().max().item()
    print(N, H, Wd, pool, "err", err, "range", want.abs().max().item(), flush=True)
    if N >= 300:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            ops.conv3x3_c64_bn_act(x, packed, b, 0.0, pool=pool)
        t0.record()
        for _ in range(10):
            ops.conv3x3_c64_bn_act(x, packed, b, 0.0, pool=pool)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print("  ms", ms, "TFLOP/s", 2 * 64 * 64 * 9 * H * Wd * N / ms / 1e9, flush=True)
