true
timeout 120 python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from audio_fewshot_b200 import ops
from audio_fewshot_b200.episode import EpisodeTable
dev = torch.device("cuda", 0)
flush = torch.empty(64 * 2**20, device=dev)
for tag, E, W, S, Q, D, mode in (("C1 D=1600 5w5s15q", 256, 5, 5, 15, 1600, "euclidean"), ("C1 D=1600 E=2048", 2048, 5, 5, 15, 1600, "euclidean"), ("C2 D=12800 5w1s15q", 64, 5, 1, 15, 12800, "euclidean"), ("C2 D=12800 E=512", 512, 5, 1, 15, 12800, "euclidean"),
                                 ("C4 D=2080 5w5s10q", 256, 5, 5, 10, 2080, "euclidean"), ("C1 cosine", 256, 5, 5, 15, 1600, "cos_sim"), ("C1 E=32", 32, 5, 5, 15, 1600, "euclidean")):
    N = E * W * (S + Q)
    feat = torch.randn(N, D, device=dev)
    tab = EpisodeTable(E, W, S, Q, np.ones(E * W * Q, dtype=np.int64), dev)
    for _ in range(3): ops.proto_logits(feat, tab.cls_row, E, W, S, mode)
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.proto_logits(feat, tab.cls_row, E, W, S, mode); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort(); ms = ts[10]
    nbytes = E * (4 * W * (S + Q) * D + 4 * W * Q * W)
    print("proto %s: %.4f ms  %.0f GB/s  frac %.3f" % (tag, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 6555.5))
PY
