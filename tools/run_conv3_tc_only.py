"""Two launches of the block-2 shape for ncu (development helper): the TF32 kernel, then the bf16 kernel."""
import sys
import torch
sys.path.insert(0, ".")
from audio_fewshot_b200 import ops
dev = torch.device("cuda", 0)
x = torch.randn(800, 64, 42, 52, device=dev).contiguous(memory_format=torch.channels_last)
w = torch.randn(64, 64, 3, 3, device=dev) * 0.06
b = torch.randn(64, device=dev)
packed = torch.from_numpy(ops.conv3x3_c64_pack_weights(w)).to(dev)
x16 = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
packed16 = torch.from_numpy(ops.conv3x3_c64_pack_weights_bf16(w)).to(dev).view(torch.bfloat16)
for _ in range(2):
    ops.conv3x3_c64_bn_act(x, packed, b, 0.0, pool=True)
    ops.conv3x3_c64_bn_act_bf16(x16, packed16, b, 0.0, pool=True)
torch.cuda.synchronize()
