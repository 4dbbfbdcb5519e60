"""Per-role wait / work cycles of the tensor-core log-mel kernel (CTA 0).  Development helper: needs a library built
with AFS_TC_PROFILE=1 (`AFS_TC_PROFILE=1 python -m audio_fewshot_b200.build --force`)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200 import _lib
from audio_fewshot_b200.frontend import LogMelFrontEnd
dev = torch.device("cuda", 0)
hop, L = (102, 16000) if "s1" in sys.argv else (512, 80000)
fr = LogMelFrontEnd(hop_length=hop, n_mels=128, mean=-15.0, std=26.0, engine="tc").to(dev).eval()
wav = torch.randn(3200, L, device=dev) * 0.1
out = torch.empty(3200, 1, 128, 157, device=dev)
for _ in range(3):
    fr(wav, out=out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); fr(wav, out=out); b.record(); torch.cuda.synchronize()
raw = C.CDLL(_lib.lib_path())
buf = (C.c_longlong * 96)()
assert raw.afs_logmel_tc_profile_read(buf) == 0
p = list(buf)
print("kernel %.3f ms; cycles PER CHUNK (wait on mbarriers / everything else):" % a.elapsed_time(b))
for name, s in (("LD ", 0), ("MMA", 8), ("E1 ", 16), ("E3 ", 24)) + tuple(("MEL%d" % w, 32 + 4 * w) for w in range(8)):
    n = max(p[s + 2], 1)
    print("  %s: wait %6.0f  work %6.0f   (%d chunks)" % (name, p[s] / n, p[s + 1] / n, n))
n = max(p[2], 1)
print("  LD segments : load+window+max %6.0f | role barrier %6.0f | scale+wait+split+store %6.0f | fence+arrive %6.0f" % tuple(p[64 + k] / n for k in range(4)))
for name, s in (("MEL warp 0 (short filters)", 68), ("MEL warp 7 (long filters)", 72)):
    print("  %s: dot loop %6.0f | shuffle+rescale %6.0f | role barriers+tile %6.0f | store+loop %6.0f" % ((name,) + tuple(p[s + k] / n for k in range(4))))
