"""BDC pooling: tensor-core Gram (csrc/bdc_tc.cu) vs the fp32 FMA kernel (csrc/bdc.cu), agreement and time at the
C4 shape (64 x 16 x 19 maps).  Development helper."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
lt = torch.tensor([float(np.log(1.0 / 200.0))], device=dev)
x = torch.relu(torch.randn(7, 64, 16, 19, device=dev))
ops.bdc_set_tensor_core(False); a = ops.bdc_pool(x, lt)
ops.bdc_set_tensor_core(True); b = ops.bdc_pool(x, lt)
print("max abs diff tc vs fp32:", (a - b).abs().max().item(), "max |out|", a.abs().max().item(), "nan", int(torch.isnan(b).sum()))
flush = torch.empty(64 * 2 ** 20, device=dev)
for B in (2000, 8000):
    x = torch.relu(torch.randn(B, 64, 16, 19, device=dev))
    for tc in (False, True):
        ops.bdc_set_tensor_core(tc)
        for _ in range(3): ops.bdc_pool(x, lt)
        ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.bdc_pool(x, lt); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[7]
        nbytes = B * (4 * 64 * 304 + 4 * 2080)
        print("bdc %s B=%d: %.4f ms  %.0f GB/s  frac %.3f" % ("tcgen05" if tc else "fp32   ", B, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 6555.5))
