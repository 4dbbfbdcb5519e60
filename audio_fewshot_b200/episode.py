"""Episode row tables: AbstractModel.split_by_episode as index math.

The reference slices and re-stacks the feature tensor on the host for every
episode (libfewshot_core/model/abstract_model.py:176-332: numpy cumsum,
E*W Python iterations, vstack copies).  The kernels here never move rows: they
read the backbone output in place through three small int32 tables.

Row layout (abstract_model.py:215-252): episode-major,
class-major; block g = e*W + w holds S support rows followed by the windows of
that class's Q queries, query k having repeats[k] windows.

    cls_row  [E*W+1]  first feature row of block g (cls_row[E*W] = N)
    q_start  [nq+1]   first output (window) row of query k; output rows are the
                      query windows in feature order, which is the order
                      torch.cat produces at proto_net.py:106-113
    q_target [nq]     local label of query k = its class index w
                      (abstract_model.py:167-174, :264-269)
"""
from collections import OrderedDict

import numpy as np
import torch


class EpisodeTable:
    __slots__ = ("E", "W", "S", "Q", "N", "NQ", "nq", "cls_row", "q_start", "q_target", "q_target_long",
                 "cls_row_host", "q_start_host", "_rows")

    def __init__(self, E, W, S, Q, repeats_host, device):
        rep = np.asarray(repeats_host, dtype=np.int64).reshape(-1)
        if rep.size != E * W * Q:
            raise ValueError("repeats must have E*W*Q = %d entries, got %d" % (E * W * Q, rep.size))
        if (rep < 0).any():
            raise ValueError("repeats must be non-negative")
        per_block = rep.reshape(E * W, Q).sum(axis=1)
        cum = np.concatenate([[0], np.cumsum(per_block)])
        g = np.arange(E * W + 1, dtype=np.int64)
        cls_row = g * S + cum
        q_start = np.concatenate([[0], np.cumsum(rep)])
        q_target = np.repeat(np.tile(np.arange(W, dtype=np.int64), E), Q)
        self.E, self.W, self.S, self.Q = E, W, S, Q
        self.N = int(cls_row[-1])
        self.NQ = int(q_start[-1])
        self.nq = int(rep.size)
        self.cls_row_host = cls_row.astype(np.int32)
        self.q_start_host = q_start.astype(np.int32)
        self.cls_row = torch.from_numpy(self.cls_row_host).to(device)
        self.q_start = torch.from_numpy(self.q_start_host).to(device)
        self.q_target = torch.from_numpy(q_target.astype(np.int32)).to(device)
        self.q_target_long = self.q_target.long()
        self._rows = None

    def episode_rows(self):
        """(support_rows[e], query_rows[e]) int64 device index tensors per episode, in the order the
        reference stacks them (abstract_model.py:274-332, mode 2): class-major, supports before queries."""
        if self._rows is None:
            dev = self.cls_row.device
            sup, qry = [], []
            for e in range(self.E):
                blocks = range(e * self.W, (e + 1) * self.W)
                s_idx = np.concatenate([np.arange(self.cls_row_host[g], self.cls_row_host[g] + self.S) for g in blocks])
                q_idx = np.concatenate([np.arange(self.cls_row_host[g] + self.S, self.cls_row_host[g + 1])
                                        for g in blocks])
                sup.append(torch.from_numpy(s_idx.astype(np.int64)).to(dev))
                qry.append(torch.from_numpy(q_idx.astype(np.int64)).to(dev))
            self._rows = (sup, qry)
        return self._rows


class EpisodeTableCache:
    """Tables depend only on (E, W, S, Q, repeats); evaluation loops reuse a handful of them."""

    def __init__(self, max_entries=64):
        self._cache = OrderedDict()
        self._max = max_entries

    def get(self, E, W, S, Q, repeats, device):
        if repeats is None:
            rep = None
            key = (E, W, S, Q, None, str(device))
        else:
            rep = repeats.detach().cpu().numpy() if isinstance(repeats, torch.Tensor) else np.asarray(repeats)
            rep = rep.astype(np.int64).reshape(-1)
            key = (E, W, S, Q, rep.tobytes(), str(device))
        tab = self._cache.get(key)
        if tab is None:
            if rep is None:
                rep = np.ones(E * W * Q, dtype=np.int64)
            tab = EpisodeTable(E, W, S, Q, rep, device)
            self._cache[key] = tab
            if len(self._cache) > self._max:
                self._cache.popitem(last=False)
        else:
            self._cache.move_to_end(key)
        return tab


def episode_size_from_batch(n_rows, W, S, Q, repeats=None, support_size=0):
    """abstract_model.py:184-191."""
    if repeats is not None:
        return (len(repeats) + support_size) // (W * (S + Q))
    return n_rows // (W * (S + Q))
