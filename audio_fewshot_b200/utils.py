"""The helpers of `libfewshot_core.utils` that the hot path's callers import by name
(libfewshot_core/utils/utils.py:20-35, 84-121, 148-159, 432-471), with the reference's signatures and return types,
backed by the device kernels: one launch per call, no per-query host synchronisation.

    from audio_fewshot_b200.utils import accuracy, majority_vote, vote_catagorical_acc, average_logits, \\
        mean_confidence_interval, get_instance
"""
import numpy as np
import torch

from . import ops
from .dist import mean_confidence_interval  # noqa: F401  (utils.py:148-159)
from .model import get_instance  # noqa: F401  (utils.py:20-35)
from .model.proto_net import accuracy_percent


def accuracy(output, target, topk=1):
    """Top-1 accuracy in percent as a Python float, summed over ranks when torch.distributed is initialised
    (utils.py:84-121; the hot path only ever asks for topk=1)."""
    if topk != 1:
        k = int(topk)
        with torch.no_grad():
            hit = (output.topk(k, dim=1).indices == target.view(-1, 1)).any(dim=1).float().sum(0, keepdim=True)
            n = target.size(0)
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(hit, op=dist.ReduceOp.SUM)
                n *= dist.get_world_size()
            return hit.mul_(100.0 / n).item()
    return accuracy_percent(output, target)


def _groups(query_nums, device):
    nums = query_nums.detach().cpu().numpy() if isinstance(query_nums, torch.Tensor) else np.asarray(query_nums)
    nums = nums.astype(np.int64).reshape(-1)
    q_start = np.concatenate([[0], np.cumsum(nums)]).astype(np.int32)
    return nums, torch.from_numpy(q_start).to(device)


def majority_vote(soft_logits, query_nums, tie_rule="torch_cuda"):
    """Per-query mode of the window arg-maxes (utils.py:436-446): returns a float32 CPU tensor [len(query_nums)] as the
    reference does.  tie_rule "torch_cuda" reproduces torch.mode on a CUDA slice (the reference's live path),
    "smallest" the CPU rule."""
    nums, q_start = _groups(query_nums, soft_logits.device)
    dummy = torch.zeros(len(nums), dtype=torch.int32, device=soft_logits.device)
    pred, _, _ = ops.vote_acc(soft_logits.float(), q_start, dummy, tie_rule=tie_rule)
    return pred.float().cpu()


def vote_catagorical_acc(targets, predictions):
    """utils.py:432-433 (the reference's spelling)."""
    return (predictions == targets).sum().float() / targets.size(0) * 100.0


def average_logits(soft_logits, query_nums):
    """Mean of each query's window logits (utils.py:449-471), [len(query_nums), W]; empty groups give zeros."""
    nums, _ = _groups(query_nums, soft_logits.device)
    n_q, W = len(nums), soft_logits.size(1)
    used = int(nums.sum())
    seg = torch.from_numpy(np.repeat(np.arange(n_q), nums)).to(soft_logits.device)
    out = torch.zeros((n_q, W), dtype=soft_logits.dtype, device=soft_logits.device)
    out.index_add_(0, seg, soft_logits[:used])
    cnt = torch.from_numpy(np.maximum(nums, 1)).to(soft_logits.device, soft_logits.dtype)
    return out / cnt.view(-1, 1)
