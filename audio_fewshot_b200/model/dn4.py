"""DN4 behind the reference's API (libfewshot_core/model/metric/dn4.py:78-155)."""
import torch
from torch import nn

from .. import ops
from .abstract_model import MetricModel
from .proto_net import accuracy_percent


class DN4(MetricModel):
    def __init__(self, n_k=3, **kwargs):
        super().__init__(**kwargs)
        self.n_k = n_k
        self.loss_func = nn.CrossEntropyLoss()

    def set_forward(self, batch, update_threshold=False, enhance_classification_via_energy=False):
        image, repeats, support_size = self._unpack(batch)
        feat = self.emb_func(image)  # [N, C, H, W]
        tab = self._table(feat.shape[0], repeats, support_size)
        output, _, _ = ops.dn4_scores(feat, tab.cls_row, tab.E, tab.W, tab.S, self.n_k)
        _, acc, _ = ops.vote_acc(output, tab.q_start, tab.q_target)
        return output, acc

    def set_forward_loss(self, batch):
        # The DN4 kernel has no backward yet: training this head is outside the built path.
        raise NotImplementedError("DN4.set_forward_loss: backward of the DN4 kernel is not built")
