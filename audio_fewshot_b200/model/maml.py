"""MAML behind the reference's API (libfewshot_core/model/meta/maml.py:30-161 and
libfewshot_core/model/backbone/utils/maml_module.py:11-146).

What stays PyTorch: the second-order inner loop.  Every inner step differentiates through the
backbone with create_graph=True (maml.py:141) -- autograd through cuDNN; no kernel of ours sits
inside it.  What this class changes around it:

  * fast weights are a name -> tensor dict fed to torch.func.functional_call instead of `.fast`
    attributes patched onto Parameters and three *_fw module classes;
  * episodes are cut out of the batch with the EpisodeTable row indices (one index_select per
    episode) instead of the E*W Python slicing loop of split_by_episode mode 2;
  * evaluation votes and scores on the device (afs_vote_acc), no per-query host sync;
  * data-parallel training can reduce gradients with one flat NCCL all-reduce
    (audio_fewshot_b200.dist.all_reduce_gradients) -- MAML is excluded from SyncBN in the
    reference as well (trainer.py:489-502).

Semantics kept, quirks included: BatchNorm2d always normalises with batch statistics and never
updates running stats (BatchNorm2d_fw, maml_module.py:78-108); only Linear / Conv2d / BatchNorm2d
parameters are adapted -- BatchNorm1d of Conv64F.logits is left slow because convert_maml_module
(maml_module.py:111-146) never converts it; adaptation puts emb_func and the classifier in train()
mode and leaves them there, so Dropout(0.3) is live on the query pass, at test time too
(maml.py:131-132); test-time adaptation needs grad mode enabled (trainer.py:259).
"""
import torch
import torch.nn.functional as F
from torch import nn
from torch.func import functional_call

from .. import ops
from .abstract_model import AbstractModel, ModelType
from .proto_net import accuracy_percent


class MetaModel(AbstractModel):
    """libfewshot_core/model/meta/meta_model.py:10-31."""

    def __init__(self, init_type="normal", **kwargs):
        super().__init__(init_type, ModelType.META, **kwargs)

    def sub_optimizer(self, parameters, config):
        kwargs = dict()
        if config["kwargs"] is not None:
            kwargs.update(config["kwargs"])
        return getattr(torch.optim, config["name"])(parameters, **kwargs)


class BatchStatNorm2d(nn.BatchNorm2d):
    """BatchNorm2d_fw (maml_module.py:78-108): batch statistics always, throw-away running buffers.
    Same parameter / buffer names as nn.BatchNorm2d, so reference checkpoints load."""

    def forward(self, x):
        c = x.shape[1]
        return F.batch_norm(x, x.new_zeros(c), x.new_ones(c), self.weight, self.bias, training=True, momentum=1)


def convert_maml_module(module):
    """Swap every BatchNorm2d for BatchStatNorm2d in place (Linear / Conv2d need no subclass here: fast
    weights arrive through functional_call).  Returns the module."""
    for name, child in list(module.named_children()):
        if isinstance(child, nn.BatchNorm2d) and not isinstance(child, BatchStatNorm2d):
            new = BatchStatNorm2d(child.num_features)
            new.load_state_dict(child.state_dict())
            setattr(module, name, new)
        else:
            convert_maml_module(child)
    return module


class MAMLLayer(nn.Module):  # maml.py:30-36
    def __init__(self, feat_dim=64, way_num=5):
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(feat_dim, way_num))

    def forward(self, x):
        return self.layers(x)


class _Net(nn.Module):
    """emb_func + classifier as one functional_call target (parameter names 'emb_func.*', 'classifier.*')."""

    def __init__(self, emb_func, classifier):
        super().__init__()
        self.emb_func = emb_func
        self.classifier = classifier

    def forward(self, x):
        return self.classifier(self.emb_func(x))


class MAML(MetaModel):
    def __init__(self, inner_param, feat_dim, **kwargs):
        super().__init__(**kwargs)
        self.feat_dim = feat_dim
        self.loss_func = nn.CrossEntropyLoss()
        self.classifier = MAMLLayer(feat_dim, way_num=self.way_num)
        self.inner_param = inner_param
        convert_maml_module(self)
        fast_types = (nn.Linear, nn.Conv2d, nn.BatchNorm2d)
        self._fast_names = [
            "%s.%s" % (mname, pname) if mname else pname
            for mname, mod in self._net().named_modules() if isinstance(mod, fast_types)
            for pname, _ in mod.named_parameters(recurse=False)]

    def _net(self):
        # not registered as a sub-module: state_dict keys stay 'emb_func.*' / 'classifier.*' only
        net = self.__dict__.get("_net_view")
        if net is None:
            net = _Net(self.emb_func, self.classifier)
            self.__dict__["_net_view"] = net
        return net

    def forward_output(self, x, fast=None):
        if fast is None:
            return self.classifier(self.emb_func(x))
        return functional_call(self._net(), fast, (x,))

    # ------------------------------------------------------------------ inner loop (maml.py:125-161)
    def set_forward_adaptation(self, support_set, support_target):
        lr = self.inner_param["lr"]
        named = dict(self._net().named_parameters())
        fast = {n: named[n] for n in self._fast_names}
        self.emb_func.train()
        self.classifier.train()
        n_iter = self.inner_param["train_iter"] if self.training else self.inner_param["test_iter"]
        for _ in range(n_iter):
            loss = self.loss_func(functional_call(self._net(), fast, (support_set,)), support_target)
            names = list(fast.keys())
            grads = torch.autograd.grad(loss, [fast[n] for n in names], create_graph=True, allow_unused=True)
            fast = {n: (fast[n] if g is None else fast[n] - lr * g) for n, g in zip(names, grads)}
        return fast

    def _episodes(self, batch):
        image, repeats, support_size = self._unpack(batch)
        tab = self._table(image.shape[0], repeats, support_size)
        sup_idx, qry_idx = tab.episode_rows()
        return image, tab, sup_idx, qry_idx

    def _adapt_all(self, image, tab, sup_idx, qry_idx):
        support_target = torch.arange(tab.W, device=image.device).repeat_interleave(tab.S)  # abstract_model.py:167-174
        outs = []
        for e in range(tab.E):
            fast = self.set_forward_adaptation(image.index_select(0, sup_idx[e]), support_target)
            outs.append(self.forward_output(image.index_select(0, qry_idx[e]), fast))
        return torch.cat(outs, dim=0)

    def set_forward(self, batch, update_threshold=False, enhance_classification_via_energy=False):
        image, tab, sup_idx, qry_idx = self._episodes(batch)
        with torch.enable_grad():  # the inner loop differentiates even at test time
            output = self._adapt_all(image, tab, sup_idx, qry_idx).detach()
        _, acc, _ = ops.vote_acc(output, tab.q_start, tab.q_target)
        return output, acc

    def set_forward_loss(self, batch):
        image, tab, sup_idx, qry_idx = self._episodes(batch)
        if tab.NQ != tab.nq:
            raise ValueError("MAML.set_forward_loss needs one window per query (the reference's loss compares "
                             "per-window outputs with per-query targets, maml.py:120-121)")
        output = self._adapt_all(image, tab, sup_idx, qry_idx)
        target = tab.q_target_long
        loss = self.loss_func(output, target)
        acc = accuracy_percent(output, target, as_tensor=getattr(self, "acc_on_device", False))
        return output, acc, loss
