"""Experiment: MAML train step with channels_last parameters/inputs and cudnn.benchmark (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200 import model as arch
dev = torch.device("cuda", 0)
E, W, S, Q = 2, 5, 5, 10
target = torch.arange(W).repeat_interleave(S + Q).repeat(E)
def run(tag, cl, bench):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    emb = arch.Conv64F(is_flatten=True, num_channels=1)
    m = arch.MAML(inner_param={"lr": 0.01, "train_iter": 5, "test_iter": 10}, feat_dim=1600, way_num=5, shot_num=S,
                  query_num=Q, test_way=5, test_shot=S, test_query=Q, emb_func=emb, device=dev).to(dev).train()
    x = torch.randn(E * W * (S + Q), 1, 128, 157, device=dev)
    if cl:
        m = m.to(memory_format=torch.channels_last)
        n, c, h, w = x.shape
        x = x.as_strided((n, c, h, w), (h * w, 1, w, 1))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    def step():
        opt.zero_grad(set_to_none=True)
        out, acc, loss = m([x, target])
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        step()
    t1.record(); torch.cuda.synchronize()
    print(tag, "%.1f ms/step" % (t0.elapsed_time(t1) / 5), flush=True)
run("nchw", False, False)
run("nchw + cudnn.benchmark", False, True)
run("channels_last", True, False)
run("channels_last + cudnn.benchmark", True, True)
