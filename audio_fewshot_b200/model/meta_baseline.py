"""MetaBaseline behind the reference's API (libfewshot_core/model/metric/meta_baseline.py:49-332):
cosine similarity to the class prototypes times a learnable temperature (`temp`, initialised to 10).

The cosine head is the prototype kernel in AFS_PROTO_COSINE mode (csrc/proto.cu: F.normalize on query and
prototype, eps 1e-12) and its backward afs_proto_bwd_cos; the temperature stays a torch parameter so that
autograd delivers its gradient.  The reference's plotting helper (visualize_features, :56-260) is out of
scope."""
import torch
from torch import nn

from .. import ops
from .abstract_model import MetricModel
from .proto_net import accuracy_percent


class MetaBaseline(MetricModel):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.loss_func = nn.CrossEntropyLoss()
        self.temp = nn.Parameter(torch.tensor(10.0))

    def set_forward(self, batch, update_threshold=False, enhance_classification_via_energy=False):
        image, repeats, support_size = self._unpack(batch)
        feat = self.emb_func(image)
        tab = self._table(feat.shape[0], repeats, support_size)
        output = ops.proto_logits(feat, tab.cls_row, tab.E, tab.W, tab.S, "cos_sim") * self.temp  # :292-294
        _, acc, _ = ops.vote_acc(output, tab.q_start, tab.q_target)
        return output, acc

    def set_forward_loss(self, batch):
        image, _, _ = self._unpack(batch)
        emb = self.emb_func(image)
        tab = self._table(emb.shape[0], None, 0)  # the reference ignores repeats here (:322-324)
        output = ops.proto_logits(emb, tab.cls_row, tab.E, tab.W, tab.S, "cos_sim") * self.temp
        target = tab.q_target_long
        loss = self.loss_func(output, target)
        acc = accuracy_percent(output, target, as_tensor=getattr(self, "acc_on_device", False))
        return output, acc, loss
