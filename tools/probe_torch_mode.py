"""Measure what torch.mode returns on a CUDA slice when several labels tie (the reference runs
torch.mode on CUDA slices: libfewshot_core/utils/utils.py:443 via proto_net.py:116).  Run on the GPU
box; writes gpurun_out/torch_mode_cuda.npz, which is committed as tests/golden/torch_mode_cuda.npz and
pins oracle.heads.torch_mode_cuda and the AFS_VOTE_TIE_TORCH_CUDA rule of csrc/vote.cu."""
import numpy as np
import torch

rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
MAXN = 320
labels = np.full((6000, MAXN), -1, dtype=np.int8)
n_arr = np.zeros(6000, dtype=np.int32)
got = np.zeros(6000, dtype=np.int8)
for i in range(6000):
    n = int(rng.integers(1, 14)) if i < 4000 else int(rng.integers(14, MAXN + 1))
    W = int(rng.integers(2, 8))
    if i >= 4000 and i % 2 == 0:  # force exact ties between two or three labels in long slices
        k = int(rng.integers(2, min(W, 3) + 1))
        y = np.repeat(rng.permutation(W)[:k], n // k)
        y = np.concatenate([y, rng.integers(0, W, size=0)])
        rng.shuffle(y)
        n = len(y)
    else:
        y = rng.integers(0, W, size=n)
    labels[i, :n] = y
    n_arr[i] = n
    got[i] = torch.mode(torch.from_numpy(y.astype(np.int64)).to(dev))[0].item()
np.savez_compressed("gpurun_out/torch_mode_cuda.npz", labels=labels, n=n_arr, mode=got,
                    torch_version=np.array(torch.__version__), gpu=np.array(torch.cuda.get_device_name(0)))
print("wrote", len(n_arr), "slices; torch", torch.__version__)
