"""Correctness (vs the exact-fp32 stem) and timing of csrc/conv1_tc.cu at the bench shape (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
for N in (800, 3200):
    x = torch.randn(N, 1, 128, 157, device=dev)
    w = np.random.default_rng(0).standard_normal((64, 9)).astype(np.float32) * 0.3
    b = np.random.default_rng(1).standard_normal(64).astype(np.float32)
    ref = ops.conv1_bn_act_pool3(x[:64], w, b, 0.0)
    got = ops.conv1_bn_act_pool3(x[:64], w, b, 0.0, tf32=True)
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    for _ in range(3):
        ops.conv1_bn_act_pool3(x, w, b, 0.0, tf32=True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        ops.conv1_bn_act_pool3(x, w, b, 0.0, tf32=True)
    t1.record(); torch.cuda.synchronize()
    print("N=%d rel err %.2e  %.4f ms" % (N, err, t0.elapsed_time(t1) / 10), flush=True)
