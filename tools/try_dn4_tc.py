import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from audio_fewshot_b200 import ops
from audio_fewshot_b200.episode import EpisodeTable
dev = torch.device("cuda", 0)
cases = [(1,5,5,15,64,20,3,None), (1,5,1,10,64,20,3,None), (2,5,5,10,64,20,1,None), (2,3,2,3,40,21,5,None),
         (2,5,8,4,64,20,3,None), (3,5,5,4,64,20,3,"ragged"), (1,5,5,2,128,12,3,None)]
for (E,W,S,Q,C,HW,nk,rag) in cases:
    rng = np.random.default_rng(E*100+C+HW)
    rep = rng.integers(1,4,size=E*W*Q) if rag else np.ones(E*W*Q, dtype=np.int64)
    N = E*W*S + int(rep.sum())
    x = np.maximum(rng.standard_normal((N,C,HW)),0).astype(np.float32) + 0.01*np.abs(rng.standard_normal((N,C,HW))).astype(np.float32)
    feat = torch.from_numpy(x).to(dev)
    tab = EpisodeTable(E,W,S,Q,rep,dev)
    s0,i0,p0 = ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, want_topk=True, want_pred=True)
    torch.cuda.synchronize()
    s1,i1,p1 = ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, want_topk=True, want_pred=True, precision="tf32")
    torch.cuda.synchronize()
    rel = ((s1-s0).abs().max()/s0.abs().max()).item()
    agree = (i0==i1).float().mean().item()
    s2,i2,p2 = ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, want_topk=True, want_pred=True, precision="tf32_staged")
    torch.cuda.synchronize()
    print("case",(E,W,S,Q,C,HW,nk,rag),"rel err %.2e"%rel,"idx agree %.4f"%agree,"pred equal",bool((p0==p1).all()),
          "| tma vs staged: max diff %.2e idx equal %s" % ((s1-s2).abs().max().item(), bool((i1==i2).all())), flush=True)
# timing at C3 x128 episodes
E,W,S,Q,C,HW,nk = 128,5,5,15,64,20,3
feat = torch.rand(E*W*(S+Q),C,HW,device=dev)
tab = EpisodeTable(E,W,S,Q,np.ones(E*W*Q,dtype=np.int64),dev)
for prec in ("fp32","tf32_staged","tf32"):
    for _ in range(3): ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, precision=prec)
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.dn4_scores(feat, tab.cls_row, E,W,S,nk, precision=prec)
    b.record(); b.synchronize()
    print(prec, "ms per call (128 episodes)", a.elapsed_time(b)/10)
