import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for (N,H,Wd,slope) in [(3,128,157,0.0),(2,7,11,0.2),(1,3,3,0.0),(130,9,10,0.1),(37,128,157,0.0)]:
    x = torch.from_numpy(rng.standard_normal((N,1,H,Wd)).astype(np.float32)).to(dev)
    w = rng.standard_normal((64,9)).astype(np.float32)*0.3; w[5]*=-1
    b = rng.standard_normal(64).astype(np.float32)
    a = ops.conv1_bn_act_pool3(x,w,b,slope)
    t = ops.conv1_bn_act_pool3(x,w,b,slope,tf32=True)
    torch.cuda.synchronize()
    print((N,H,Wd,slope), "max abs diff %.3e  (max |a| %.2f)" % ((a-t).abs().max().item(), a.abs().max().item()), flush=True)
x = torch.randn(800,1,128,157,device=dev)
w = rng.standard_normal((64,9)).astype(np.float32); b = np.zeros(64,np.float32)
for tf in (False, True):
    for _ in range(3): ops.conv1_bn_act_pool3(x,w,b,0.0,tf32=tf)
    torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.conv1_bn_act_pool3(x,w,b,0.0,tf32=tf)
    e1.record(); e1.synchronize()
    print("tf32" if tf else "fp32", "ms per 800 clips", e0.elapsed_time(e1)/10)
