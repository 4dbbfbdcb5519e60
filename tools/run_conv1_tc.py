import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
x = torch.randn(800,1,128,157,device=dev)
w = np.random.default_rng(0).standard_normal((64,9)).astype(np.float32); b = np.zeros(64,np.float32)
for _ in range(3): ops.conv1_bn_act_pool3(x,w,b,0.0,tf32=True)
torch.cuda.synchronize(); print("ok")
