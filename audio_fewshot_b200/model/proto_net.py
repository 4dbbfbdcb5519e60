"""ProtoNet behind the reference's API (libfewshot_core/model/metric/proto_net.py:67-154).

set_forward:  image -> emb_func -> [fused prototype + distance kernel over all episodes]
              -> [vote + accuracy kernel] -> (raw logits [sum R, W], acc 0-dim tensor in %).
The reference loops over episodes in Python (proto_net.py:107-112), materialises the
broadcast difference, and syncs once per query inside majority_vote (utils.py:441-445);
here the head is two kernel launches and no host sync.
"""
import torch
from torch import nn

from .. import ops
from .abstract_model import MetricModel


def accuracy_percent(output, target, as_tensor=False):
    """utils.accuracy (libfewshot_core/utils/utils.py:84-121), top-1: percent as a Python float,
    summed over ranks when torch.distributed is initialised.  as_tensor=True keeps the value on the
    device (a 1-element tensor, no host sync) -- what a CUDA-graph-captured train step needs."""
    import torch.distributed as dist

    with torch.no_grad():
        n = target.size(0)
        correct = (output.argmax(dim=1) == target).float().sum(0, keepdim=True)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(correct, op=dist.ReduceOp.SUM)
            n *= dist.get_world_size()
        correct.mul_(100.0 / n)
        return correct if as_tensor else correct.item()


class ProtoNet(MetricModel):
    def __init__(self, distance="euclidean", precision="fp32", **kwargs):
        """precision (not a reference kwarg): "fp32" is the parity path; "tf32" evaluates the logits as
        -(|q|^2 - 2 q.p + |p|^2) with q.p on the tcgen05 tensor cores (euclidean, evaluation only; a separate
        precision class, see csrc/proto_tc.cu)."""
        super().__init__(**kwargs)
        self.distance = distance
        self.precision = precision
        self.loss_func = nn.CrossEntropyLoss()
        self.is_clap = kwargs.get("is_clap", False)

    def set_forward(self, batch, update_threshold=False, enhance_classification_via_energy=False):
        image, repeats, support_size = self._unpack(batch)
        feat = self.emb_func(image)
        tab = self._table(feat.shape[0], repeats, support_size)
        output = ops.proto_logits(feat, tab.cls_row, tab.E, tab.W, tab.S, self.distance, precision=self.precision)
        _, acc, _ = ops.vote_acc(output, tab.q_start, tab.q_target)
        return output, acc

    def set_forward_loss(self, batch):
        image, _, _ = self._unpack(batch)
        emb = image if self.is_clap else self.emb_func(image)
        # the reference ignores `repeats` on this path (proto_net.py:144-146): fixed layout
        tab = self._table(emb.shape[0], None, 0)
        output = ops.proto_logits(emb, tab.cls_row, tab.E, tab.W, tab.S, self.distance)
        target = tab.q_target_long
        loss = self.loss_func(output, target)
        acc = accuracy_percent(output, target, as_tensor=getattr(self, "acc_on_device", False))
        return output, acc, loss
