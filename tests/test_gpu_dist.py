"""The path's collectives on real GPUs over NCCL: per-episode results do not depend on the world size,
the accuracy gather + CI, the reference's 1-float accuracy all-reduce and the flat gradient all-reduce.
Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped on a single-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(world, out):
    worker = os.path.join(ROOT, "tests", "dist_worker.py")
    if world == 1:
        cmd = [sys.executable, worker, out]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", "29541", worker, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.load(open(out))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_world_size_invariance_and_collectives_on_nccl(tmp_path):
    one = _run(1, str(tmp_path / "w1.json"))
    two = _run(2, str(tmp_path / "w2.json"))
    # evaluation: identical per-episode accuracies in global episode order, whatever the sharding
    assert one["acc"] == two["acc"] and len(two["acc"]) == 8
    assert abs(one["mean"] - two["mean"]) < 1e-9 and abs(one["half"] - two["half"]) < 1e-9
    # training accuracy: the all-reduced percentage is the mean of the ranks' own percentages (utils.py:116-118)
    nq = 15
    per_rank = two["per_rank"]
    assert len(per_rank) == 2
    want = 100.0 * sum(r[2] for r in per_rank) / (2 * nq)
    assert abs(two["train_acc_allreduced"] - want) < 1e-3
    # flat gradient all-reduce: every rank ends with the same gradient = the mean of the local ones
    local_sums = [r[0] for r in per_rank]
    reduced = [r[1] for r in per_rank]
    assert abs(reduced[0] - reduced[1]) <= 1e-9 * max(1.0, abs(reduced[0]))
    assert abs(reduced[0] - sum(local_sums) / 2) <= 1e-6 * max(1.0, abs(reduced[0]))
    assert two["grad_elements"] == one["grad_elements"] > 0
