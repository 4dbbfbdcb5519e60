"""How far apart are two evaluations of the same MAML training step?  Eager vs eager and eager vs CUDA graph, with and
without cudnn.deterministic, all starting each step from the same weights (development helper; calibrates
tests/test_gpu_models.py::test_graphed_train_step_matches_eager_maml)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200 import model as arch
from audio_fewshot_b200.graph_step import GraphedTrainStep

cuda = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def flat(m):
    return torch.cat([p.detach().reshape(-1) for p in m.parameters()])


def run(deterministic):
    torch.backends.cudnn.deterministic = deterministic
    torch.manual_seed(3)
    emb = arch.Conv64F(is_flatten=True, num_channels=1)
    emb.logits[0].p = 0.0
    kw = dict(way_num=3, shot_num=2, query_num=3, test_way=3, test_shot=2, test_query=3, device=cuda)
    m1 = arch.MAML(inner_param={"lr": 0.01, "train_iter": 2, "test_iter": 2}, feat_dim=1600, emb_func=emb, **kw).to(cuda).train()
    m2, m3 = copy.deepcopy(m1), copy.deepcopy(m1)
    E, W, S, Q = 2, 3, 2, 3
    target = torch.arange(W).repeat_interleave(S + Q).repeat(E)
    batches = [torch.randn(E * W * (S + Q), 1, 128, 157, device=cuda) * 0.7 for _ in range(3)]
    ref = copy.deepcopy(m1.state_dict())
    opts = [torch.optim.SGD(m.parameters(), lr=1e-2) for m in (m1, m2, m3)]
    step = GraphedTrainStep(m2, opts[1], batches[0].shape, target=target, warmup=1)
    state = ref
    for i, b in enumerate(batches):
        for m in (m1, m2, m3):
            m.load_state_dict(state)
        w0 = flat(m1).clone()
        res = []
        for m, o in ((m1, opts[0]), (m3, opts[2])):
            o.zero_grad(set_to_none=True)
            _, _, loss = m([b, target])
            loss.backward()
            o.step()
            res.append((float(loss.detach()), flat(m) - w0))
        _, _, l2 = step(b)
        res.append((float(l2.detach()), flat(m2) - w0))
        (la, da), (lb, db), (lg, dg) = res
        n = da.norm().item()
        print("det=%d step %d: loss eager %.7f eager' %.7f graph %.7f | |update| %.3e  rel diff eager' %.3e graph %.3e  "
              "max|diff| eager' %.3e graph %.3e  cos graph %.6f" % (
                  deterministic, i, la, lb, lg, n, (db - da).norm().item() / n, (dg - da).norm().item() / n,
                  (db - da).abs().max().item(), (dg - da).abs().max().item(),
                  torch.nn.functional.cosine_similarity(da, dg, dim=0).item()), flush=True)
        state = copy.deepcopy(m1.state_dict())


for det in (False, True):
    try:
        run(det)
    except Exception as e:  # noqa: BLE001
        print("det=%d failed: %r" % (det, e), flush=True)
