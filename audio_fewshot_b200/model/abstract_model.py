"""Base classes of the drop-in boundary.

Mirrors the caller-facing contract of the reference's AbstractModel / MetricModel
(libfewshot_core/model/abstract_model.py:125-162,414-429; metric/metric_model.py:8-18):
constructor kwargs become attributes, `forward` dispatches on `self.training`,
`reverse_setting_info` swaps train/test episode settings, `model_type` tells the
trainer whether gradients are needed (trainer.py:259).  `split_by_episode` is NOT
reproduced as tensor slicing: heads consume an EpisodeTable (episode.py).
"""
from enum import Enum

import torch
from torch import nn

from ..episode import EpisodeTableCache


class ModelType(Enum):  # reference libfewshot_core/utils/enum_type.py:5-9
    ABSTRACT = 0
    FINETUNING = 1
    METRIC = 2
    META = 3


class AbstractModel(nn.Module):
    def __init__(self, init_type="normal", model_type=ModelType.ABSTRACT, **kwargs):
        super().__init__()
        self.init_type = init_type
        self.model_type = model_type
        for key, value in kwargs.items():  # abstract_model.py:131-132
            setattr(self, key, value)
        self._tables = EpisodeTableCache()

    def set_forward(self, *args, **kwargs):
        raise NotImplementedError

    def set_forward_loss(self, *args, **kwargs):
        raise NotImplementedError

    def get_uncertainty_threshold(self, policy="mean"):
        return None

    def forward(self, x, update_threshold=False, enhance_classification_via_energy=False):
        # abstract_model.py:149-153 (the reference drops both kwargs; so DeepBDC's flags only
        # take effect when set_forward is called directly -- kept, and documented in DESIGN.md)
        if self.training:
            return self.set_forward_loss(x)
        return self.set_forward(x)

    def train(self, mode=True):
        # abstract_model.py:155-159 returns None; nn.Module convention (return self) is a superset
        super().train(mode)
        if hasattr(self, "distill_layer"):
            self.distill_layer.train(False)
        return self

    def reverse_setting_info(self):  # abstract_model.py:414-429
        (self.way_num, self.shot_num, self.query_num, self.test_way, self.test_shot, self.test_query) = (
            self.test_way, self.test_shot, self.test_query, self.way_num, self.shot_num, self.query_num)

    # ------------------------------------------------------------------ helpers for subclasses
    def _unpack(self, batch):
        """(image, target[, repeats, support_size]) -> image on device, repeats, support_size."""
        if len(batch) == 2:
            image, _ = batch
            repeats, support_size = None, 0
        else:
            image, _, repeats, support_size = batch
        image = image.to(self.device, non_blocking=True)
        return image, repeats, support_size

    def _table(self, n_rows, repeats, support_size):
        W, S, Q = self.way_num, self.shot_num, self.query_num
        if repeats is not None:
            E = (len(repeats) + support_size) // (W * (S + Q))  # abstract_model.py:185
        else:
            E = n_rows // (W * (S + Q))  # :189-191
        tab = self._tables.get(E, W, S, Q, repeats, torch.device(self.device))
        if tab.N != n_rows:
            raise ValueError("batch has %d rows but the episode layout implies %d" % (n_rows, tab.N))
        return tab


class MetricModel(AbstractModel):
    def __init__(self, init_type="normal", **kwargs):
        super().__init__(init_type, ModelType.METRIC, **kwargs)


class FinetuningModel(AbstractModel):
    """The third model family of the boundary (libfewshot_core/model/finetuning/finetuning_model.py:10-31): a
    backbone pre-trained with a plain classifier, adapted per episode at test time.  The contract is three
    methods -- `set_forward` (episodic evaluation), `set_forward_loss` (pre-training step on a flat batch) and
    `set_forward_adaptation` (the per-episode fine-tuning loop) -- and `sub_optimizer`, which builds the optimiser of
    that loop from a `{"name": ..., "kwargs": ...}` config block exactly as the reference does.  Concrete fine-tuning
    classifiers are outside SURVEY 8's scope; the class exists so that user code written against the reference's
    base class binds here unchanged."""

    def __init__(self, init_type="normal", **kwargs):
        super().__init__(init_type, ModelType.FINETUNING, **kwargs)

    def set_forward_adaptation(self, *args, **kwargs):
        raise NotImplementedError

    def sub_optimizer(self, model, config):  # finetuning_model.py:26-31
        kwargs = dict()
        if config["kwargs"] is not None:
            kwargs.update(config["kwargs"])
        return getattr(torch.optim, config["name"])(model.parameters(), **kwargs)

