"""Kernel-time summary of one eager MAML train step (development helper)."""
import sys
import torch
sys.path.insert(0, ".")
from audio_fewshot_b200 import model as arch
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
torch.manual_seed(0)
emb = arch.Conv64F(is_flatten=True, num_channels=1)
m = arch.MAML(inner_param={"lr": 0.01, "train_iter": 5, "test_iter": 10}, feat_dim=1600, way_num=5, shot_num=5,
              query_num=10, test_way=5, test_shot=5, test_query=10, emb_func=emb, device=dev).to(dev).train()
E, W, S, Q = 2, 5, 5, 10
x = torch.randn(E * W * (S + Q), 1, 128, 157, device=dev)
target = torch.arange(W).repeat_interleave(S + Q).repeat(E)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
def step():
    opt.zero_grad(set_to_none=True)
    out, acc, loss = m([x, target])
    loss.backward()
    opt.step()
for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
