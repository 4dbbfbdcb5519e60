"""Launch the fused log-mel kernel a few times at the bench shape (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_fewshot_b200.frontend import LogMelFrontEnd
dev = torch.device("cuda", 0)
eng = sys.argv[1] if len(sys.argv) > 1 else None  # "fft" | "tc" | default
fr = LogMelFrontEnd(hop_length=512, n_mels=128, mean=-15.0, std=26.0, engine=eng).to(dev).eval()
wav = torch.randn(800, 80000, device=dev) * 0.1
out = torch.empty(800, 1, 128, 157, device=dev)
for _ in range(5):
    fr(wav, out=out)
torch.cuda.synchronize()
print("ok", float(out.mean()))
