import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
from audio_fewshot_b200.episode import EpisodeTable
dev = torch.device("cuda", 0)
E,W,S,Q,C,HW,nk = 128,5,5,15,64,20,3
feat = torch.rand(E*W*(S+Q),C,HW,device=dev)
tab = EpisodeTable(E,W,S,Q,np.ones(E*W*Q,dtype=np.int64),dev)
for prec in ("fp32","tf32_staged","tf32"):
    for _ in range(2): ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, precision=prec)
torch.cuda.synchronize(); print("ok")
