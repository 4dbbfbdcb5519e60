import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
from audio_fewshot_b200.episode import EpisodeTable
dev = torch.device("cuda", 0)
E,W,S,Q,C,HW,nk = 128,5,5,15,64,20,3
feat = torch.rand(E*W*(S+Q),C,HW,device=dev)
tab = EpisodeTable(E,W,S,Q,np.ones(E*W*Q,dtype=np.int64),dev)
for _ in range(3): ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, precision="tf32")
torch.cuda.synchronize(); print("ok")
