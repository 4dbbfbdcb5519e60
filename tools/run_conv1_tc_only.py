"""Launch the tensor-core stem a few times at the bench shape (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
x = torch.randn(800, 1, 128, 157, device=dev)
w = np.random.default_rng(0).standard_normal((64, 9)).astype(np.float32) * 0.3
b = np.random.default_rng(1).standard_normal(64).astype(np.float32)
for _ in range(5):
    out = ops.conv1_bn_act_pool3(x, w, b, 0.0, tf32=True)
torch.cuda.synchronize()
print("ok", float(out.mean()))
