"""Timing of the bf16 backbone path against the TF32 path at the bench shape (development helper; B200).

stem (csrc/conv1_tc.cu) with fp32 / bf16 output, block 2 and block 3 (csrc/conv3_tc.cu) as TF32 / bf16 kernels, and the
whole Conv64F inference forward in both modes, 3 200 clips, CUDA events, L2 flushed between iterations."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from audio_fewshot_b200 import ops
from audio_fewshot_b200.backbone.conv_four import Conv64F

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1))
    return float(np.median(ts))


N = int(os.environ.get("N", "3200"))
rng = np.random.default_rng(0)
x = torch.randn(N, 1, 128, 157, device=dev)
w1 = rng.standard_normal((64, 9)).astype(np.float32) * 0.3
b1 = rng.standard_normal(64).astype(np.float32)
print("stem fp32 out  %.4f ms" % timeit(lambda: ops.conv1_bn_act_pool3(x, w1, b1, 0.0, tf32=True)), flush=True)
print("stem bf16 out  %.4f ms" % timeit(lambda: ops.conv1_bn_act_pool3(x, w1, b1, 0.0, tf32=True, out_dtype=torch.bfloat16)), flush=True)
h32 = ops.conv1_bn_act_pool3(x, w1, b1, 0.0, tf32=True)
h16 = ops.conv1_bn_act_pool3(x, w1, b1, 0.0, tf32=True, out_dtype=torch.bfloat16)
w = torch.from_numpy((rng.standard_normal((64, 64, 3, 3)) * 0.06).astype(np.float32)).to(dev)
b = torch.from_numpy(rng.standard_normal(64).astype(np.float32)).to(dev)
p32 = torch.from_numpy(ops.conv3x3_c64_pack_weights(w)).to(dev)
p16 = torch.from_numpy(ops.conv3x3_c64_pack_weights_bf16(w)).to(dev).view(torch.bfloat16)
print("block 2 tf32   %.4f ms" % timeit(lambda: ops.conv3x3_c64_bn_act(h32, p32, b, 0.0, pool=True)), flush=True)
print("block 2 bf16   %.4f ms" % timeit(lambda: ops.conv3x3_c64_bn_act_bf16(h16, p16, b, 0.0, pool=True)), flush=True)
g32 = ops.conv3x3_c64_bn_act(h32, p32, b, 0.0, pool=True)
g16 = ops.conv3x3_c64_bn_act_bf16(h16, p16, b, 0.0, pool=True)
print("block 2 bf16 vs tf32: max |diff| / max |ref| = %.3e" % ((g16.float() - g32).abs().max() / g32.abs().max()).item())
print("block 3 tf32   %.4f ms" % timeit(lambda: ops.conv3x3_c64_bn_act(g32, p32, b, 0.0, pool=True)), flush=True)
print("block 3 bf16   %.4f ms" % timeit(lambda: ops.conv3x3_c64_bn_act_bf16(g16, p16, b, 0.0, pool=True, out_dtype=torch.float32)), flush=True)
torch.manual_seed(0)
net = Conv64F(is_flatten=True, num_channels=1).to(dev).eval()
for m in net.modules():
    if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
        m.running_mean.normal_(0, 0.1)
        m.running_var.uniform_(0.5, 1.5)
with torch.no_grad():
    ref = net(x[:256])
    print("Conv64F tf32   %.4f ms" % timeit(lambda: net(x)), flush=True)
    net.precision = "bf16"
    fast = net(x[:256])
    print("Conv64F bf16   %.4f ms" % timeit(lambda: net(x)), flush=True)
print("features bf16 vs tf32: max |diff| / max |ref| = %.3e" % ((fast - ref).abs().max() / ref.abs().max()).item())
