"""Kernel-time summary of one ProtoNet/Conv64F training step (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200 import model as arch
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
torch.manual_seed(0)
W, S, Q, E = 5, 5, 15, 2
emb = arch.Conv64F(is_flatten=True, num_channels=1)
m = arch.ProtoNet(way_num=W, shot_num=S, query_num=Q, test_way=W, test_shot=S, test_query=Q, emb_func=emb, device=dev).to(dev).train()
if len(sys.argv) > 1 and sys.argv[1] == "cl":
    m = m.to(memory_format=torch.channels_last)
x = torch.randn(E * W * (S + Q), 1, 128, 157, device=dev)
target = torch.arange(W).repeat_interleave(S + Q).repeat(E)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
def step():
    opt.zero_grad(set_to_none=True)
    out, acc, loss = m([x, target])
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in evs)
for e in evs[:22]:
    print("%8.1f us %5.1f%% x%-3d %s" % (e.self_device_time_total, 100 * e.self_device_time_total / tot, e.count, e.key[:110]))
print("total %.1f us" % tot)
