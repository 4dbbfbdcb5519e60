"""Launch one log-mel engine a few times (ncu target).  usage: python tools/run_logmel.py <engine> [B] [hop] [L]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_fewshot_b200.frontend import LogMelFrontEnd

engine = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 800
hop = int(sys.argv[3]) if len(sys.argv) > 3 else 512
L = int(sys.argv[4]) if len(sys.argv) > 4 else 80000
dev = torch.device("cuda", 0)
fr = LogMelFrontEnd(hop_length=hop, n_mels=128, mean=-15.0, std=26.0, engine=engine).to(dev).eval()
wav = torch.randn(B, L, device=dev) * 0.1
out = torch.empty(B, 1, 128, 1 + L // hop, device=dev)
for _ in range(4):
    fr(wav, out=out)
torch.cuda.synchronize()
