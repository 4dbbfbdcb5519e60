"""Launch the tensor-core BDC kernel a few times at the C4 shape (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
lt = torch.tensor([float(np.log(1.0 / 200.0))], device=dev)
x = torch.relu(torch.randn(4000, 64, 16, 19, device=dev))
for _ in range(5):
    out = ops.bdc_pool(x, lt)
torch.cuda.synchronize()
print("ok", float(out.mean()))
