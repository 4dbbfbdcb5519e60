"""DeepBDC behind the reference's API (libfewshot_core/model/metric/deepbdc.py:56-441).

The BDC matrix itself lives in the backbone (resnet12Bdc.bdc_pool -> ops.bdc_pool); this
class is the prototype head on the 2080-d BDC vectors: -||q-p||^2 when shot_num > 1, the raw
dot product otherwise (deepbdc.py:37-53), plus the energy-score bookkeeping (:318-351).
"""
import os

import numpy as np
import torch
from torch import nn

from .. import ops
from .abstract_model import MetricModel
from .proto_net import accuracy_percent


class DeepBDC(MetricModel):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.loss_func = nn.CrossEntropyLoss()
        self.uncertainty_threshold = []

    def _mode(self):
        return "euclidean" if self.shot_num > 1 else "dot"

    def set_forward(self, batch, update_threshold=False, enhance_classification_via_energy=False):
        image, repeats, support_size = self._unpack(batch)
        feat = self.emb_func(image)
        tab = self._table(feat.shape[0], repeats, support_size)
        output = ops.proto_logits(feat, tab.cls_row, tab.E, tab.W, tab.S, self._mode())
        q_pred, acc, _ = ops.vote_acc(output, tab.q_start, tab.q_target)
        if not (update_threshold or enhance_classification_via_energy):
            return output, acc
        uncertains = ops.energy_score(output, tab.q_start, tab.nq)  # deepbdc.py:318-319
        if update_threshold:  # :320-322
            is_correct = (q_pred == tab.q_target).cpu().numpy()
            self.uncertainty_threshold.append([uncertains.detach().cpu().numpy(), is_correct])
        if not enhance_classification_via_energy:
            return output, acc
        # :326-351 -- side effect kept: append this batch's scores to test_uncertainty.npy
        fname = "test_uncertainty.npy"
        cur = uncertains.detach().cpu().numpy().ravel()
        try:
            if os.path.exists(fname):
                np.save(fname, np.concatenate([np.asarray(np.load(fname)).ravel(), cur]))
            else:
                np.save(fname, cur)
        except Exception as exc:  # same tolerance as the reference
            print("Could not save uncertainties to %s: %s" % (fname, exc))
        ood_query_mask = np.zeros(cur.shape[0], dtype=bool)
        ood_query_mask[np.argsort(-cur)[: int(0.2 * len(cur))]] = True
        query_mask = np.zeros(tab.N, dtype=bool)  # abstract_model.py:232,243-252
        for g in range(tab.E * tab.W):
            query_mask[tab.cls_row_host[g] + tab.S : tab.cls_row_host[g + 1]] = True
        return output, acc, uncertains, ood_query_mask, query_mask

    def set_forward_loss(self, batch):
        image, _, _ = self._unpack(batch)
        feat = self.emb_func(image)
        tab = self._table(feat.shape[0], None, 0)
        # the reference uses only query_feat[0] (deepbdc.py:372-374), i.e. it is only defined for
        # episode_size == 1; we compute every episode, which coincides there.
        output = ops.proto_logits(feat, tab.cls_row, tab.E, tab.W, tab.S, self._mode())
        target = tab.q_target_long
        loss = self.loss_func(output, target)
        acc = accuracy_percent(output, target, as_tensor=getattr(self, "acc_on_device", False))
        return output, acc, loss

    def get_uncertainty_threshold(self, policy="mean", normalize=False):
        """deepbdc.py:381-441: 95 % quantile of the energy of correctly classified queries
        ('overall': pooled; otherwise mean over calibration batches)."""
        if len(self.uncertainty_threshold) == 0:
            return None, None
        if policy == "overall":
            u = np.concatenate([np.asarray(i[0]).ravel() for i in self.uncertainty_threshold])
            ok = np.concatenate([np.asarray(i[1]).ravel() for i in self.uncertainty_threshold])
            self.uncertains_mean = np.mean(u[ok])
            self.uncertains_std = np.std(u[ok])
            self.uncertain_global_threshold = np.quantile(u[ok], 0.95)
            return self.uncertain_global_threshold, None
        thresholds = []
        for u, ok in self.uncertainty_threshold:
            u, ok = np.asarray(u), np.asarray(ok)
            if ok.sum() == 0:
                continue
            thresholds.append(np.quantile(u[ok], 0.95))
        self.uncertain_global_threshold = float(np.mean(thresholds)) if thresholds else None
        return self.uncertain_global_threshold, thresholds
