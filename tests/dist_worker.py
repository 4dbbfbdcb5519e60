"""Worker of tests/test_gpu_dist.py: one process per GPU (NCCL).  Evaluates 8 synthetic episodes sharded round-robin
over the ranks and runs one data-parallel training step; rank 0 writes the results as JSON to argv[1]."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from audio_fewshot_b200 import dist as afs_dist
from audio_fewshot_b200 import model as arch
from audio_fewshot_b200.frontend import LogMelFrontEnd
from audio_fewshot_b200.synthetic import name_seeded_weights_, synthetic_clip_batch


def main():
    out_path = sys.argv[1]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, S, Q, L, N_EP = 5, 1, 3, 16000, 8
    torch.manual_seed(0)
    emb = name_seeded_weights_(arch.Conv64F(is_flatten=True, num_channels=1))
    model = arch.ProtoNet(way_num=W, shot_num=S, query_num=Q, test_way=W, test_shot=S, test_query=Q, emb_func=emb,
                          device=dev).to(dev).eval()
    front = LogMelFrontEnd(hop_length=102, n_mels=128, mean=-15.1, std=26.2).to(dev).eval()
    repeats = torch.ones(W * Q, dtype=torch.long)
    target = torch.arange(W, device=dev).repeat_interleave(Q)

    # ---- evaluation: per-episode accuracies, ONE all_gather at the end (reference test.py:210)
    mine = afs_dist.shard_episodes(N_EP, rank, world)
    accs, logits = [], {}
    with torch.no_grad():
        for g in mine:
            wav = torch.from_numpy(synthetic_clip_batch(5, g, 1, W, S, Q, L)).to(dev)
            output, acc = model.set_forward([front(wav, first_clip_index=g * W * (S + Q)), None, repeats, W * S])
            accs.append((output.argmax(1) == target).float().mean() * 100.0)
            logits[g] = output.double().sum().item()
    full = afs_dist.gather_episode_accuracies(torch.stack(accs), N_EP)
    mean, half = afs_dist.mean_confidence_interval(full.tolist())

    # ---- training: the reference's 1-float accuracy all-reduce (utils.py:116-118) + ONE flat gradient all-reduce
    model.train()
    model.acc_on_device = True
    wav = torch.from_numpy(synthetic_clip_batch(6, rank, 1, W, S, Q, L)).to(dev)  # a different episode per rank
    tgt = torch.arange(W).repeat_interleave(S + Q)
    with torch.no_grad():
        image = front(wav)
    output, acc, loss = model.set_forward_loss([image, tgt])
    local_correct = (output.argmax(1) == target).float().sum().item()
    model.zero_grad(set_to_none=True)
    loss.backward()
    local_grad = torch.cat([p.grad.reshape(-1).double() for p in model.parameters() if p.grad is not None])
    local_sum = local_grad.sum().clone()
    n = afs_dist.all_reduce_gradients(model.parameters())
    red = torch.cat([p.grad.reshape(-1).double() for p in model.parameters() if p.grad is not None])
    red_sum = red.sum()
    stats = torch.stack([local_sum, red_sum, torch.tensor(local_correct, dtype=torch.float64, device=dev)])
    if world > 1:
        gathered = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
    else:
        gathered = [stats]
    if rank == 0:
        json.dump({"world": world, "acc": full.tolist(), "mean": float(mean), "half": float(half),
                   "train_acc_allreduced": float(acc.item()), "grad_elements": int(n),
                   "per_rank": [[float(v) for v in g.tolist()] for g in gathered],
                   "logit_sums": {str(k): v for k, v in logits.items()}}, open(out_path, "w"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
