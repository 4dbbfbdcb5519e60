"""Episode sharding and the (only) collectives of the path.

Episodes are independent: rank r of `world` takes the global episode indices
g with g % world == r, all randomness is keyed by g, so per-episode results do not depend on
the world size.  Evaluation needs ONE collective at the end -- an all_gather of the per-rank
accuracy vectors -- instead of the reference's 4-byte all_reduce per batch
(libfewshot_core/utils/utils.py:116-118).  Training uses DDP's gradient all-reduce
(trainer.py:504-509) unchanged: the heads are autograd Functions, so DDP sees ordinary grads.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_episodes(n_episodes, rank, world_size):
    """Global episode indices owned by `rank` (round-robin)."""
    return list(range(rank, n_episodes, world_size))


def gather_episode_accuracies(local_acc, n_episodes):
    """local_acc: 1-D float tensor, entry i = accuracy of global episode rank + i*world.
    Returns the full [n_episodes] vector (on every rank), in global episode order."""
    rank, ws = world()
    if ws == 1:
        return local_acc.detach().float().cpu()
    per_rank = (n_episodes + ws - 1) // ws
    buf = torch.full((per_rank,), float("nan"), dtype=torch.float32, device=local_acc.device)
    buf[: local_acc.numel()] = local_acc.detach().float()
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    full = torch.stack(out, dim=1).reshape(-1)[:n_episodes]  # [per_rank, ws] -> global order
    return full.cpu()


def mean_confidence_interval(data, confidence=0.95):
    """mean and half-width of the 95 % CI over per-episode accuracies
    (libfewshot_core/utils/utils.py:148-159: scipy sem * t.ppf((1+c)/2, n-1))."""
    import scipy.stats

    a = np.asarray([1.0 * np.array(d) for d in data], dtype=np.float64)
    n = len(a)
    m, se = np.mean(a), scipy.stats.sem(a)
    h = se * scipy.stats.t.ppf((1 + confidence) / 2.0, n - 1)
    return m, h


def all_reduce_gradients(parameters, average=True):
    """Data-parallel gradient reduction as ONE collective: flatten every .grad into a single fp32 buffer,
    all_reduce it (NCCL over NVLink on the GPU box), scatter it back.  Replaces the bucketed hooks of
    DistributedDataParallel(find_unused_parameters=True) (reference trainer.py:504-509) for the episodic
    models, whose whole gradient is small (Conv64F 0.86 MB, ResNet-12 49.7 MB: latency-, not
    bandwidth-bound on NVSwitch).  Parameters without a gradient contribute zeros, so ranks whose episode
    left a parameter unused stay in step.  Returns the number of reduced elements."""
    params = [p for p in parameters if p.requires_grad]
    if not params:
        return 0
    rank, ws = world()
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(ws)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off
