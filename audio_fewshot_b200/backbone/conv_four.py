"""Conv64F feature extractor, audio variant (3x3/stride-3 max-pools, 64->1600 head).

Same architecture, constructor kwargs and state_dict names as the reference's
libfewshot_core/model/backbone/conv_four.py:28-128, so `*_best.pth` checkpoints load
unchanged (layer{1..4}.0 = Conv2d with bias, layer{1..4}.1 = BatchNorm2d,
logits.1 = BatchNorm1d(64), logits.2 = Linear(64, 1600)).

Two execution paths with the same parameters:
  * training / autograd / CPU-constructed checks: the plain module graph (cuDNN + ATen), exactly the
    reference's op sequence -- except block 1 in training mode on CUDA, which runs as three fused sm_100a
    kernels (csrc/conv1_train.cu: batch statistics from the input's 9-tap autocorrelation, forward without
    the [N,64,128,157] activation, backward by recomputation);
  * inference on CUDA (eval mode, grad disabled, 1 input channel): `_forward_inference` --
    block 1 is ONE sm_100a kernel (csrc/conv1.cu / conv1_tc.cu: conv + BatchNorm + activation + max-pool,
    the [N,64,128,157] activation never reaches HBM); blocks 2-3 are ONE tcgen05 kernel each
    (csrc/conv3_tc.cu: TF32 implicit GEMM over shifted shared-memory views + folded BatchNorm + activation
    + max-pool) when TF32 convolutions are allowed, otherwise -- and block 4 always -- cuDNN's fused
    conv+bias+ReLU on channels-last tensors with BatchNorm folded into the weights; the last max-pool, the
    flatten and the final Linear (BatchNorm1d folded in) are ONE kernel (csrc/tail.cu).  In the reference's
    eager sequence those elementwise and layout kernels are 88 % of an evaluation step on B200
    (profiles/r01_bench_launches.csv).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


def pooled_extent(n, times, k=3):
    for _ in range(times):
        n = n // k
    return n


def _conv_block(c_in, c_out, act, pool, track):
    layers = [nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1),
              nn.BatchNorm2d(c_out, track_running_stats=track), act]
    if pool:
        layers.append(nn.MaxPool2d(kernel_size=3, stride=3))
    return nn.Sequential(*layers)


class Conv64F(nn.Module):
    IN_MELS, IN_FRAMES = 128, 157  # the [1,128,157] log-mel image (conv_four.py:87)

    def __init__(self, is_flatten=False, is_feature=False, leaky_relu=False, negative_slope=0.2,
                 last_pool=True, maxpool_last2=True, use_running_statistics=True, num_channels=3):
        super().__init__()
        self.is_flatten, self.is_feature = is_flatten, is_feature
        self.last_pool, self.maxpool_last2 = last_pool, maxpool_last2
        self.stem_tf32 = None  # None: follow torch.backends.cudnn.allow_tf32
        self.block_tc = None   # tcgen05 kernel for blocks 2-3; None: follow torch.backends.cudnn.allow_tf32
        # "bf16": the separately stated reduced-precision inference path -- the stem writes bf16, blocks 2-3 run as
        # bf16 tcgen05 MMAs (K = 16) with fp32 accumulation; block 4, the Linear and the head stay as they are.
        # Not the parity path (logits move by ~1e-2 relative); None / "tf32": the reference's precision class.
        self.precision = None
        self.fused_train_stem = True  # training: block 1 forward + backward as fused kernels (csrc/conv1_train.cu)
        act = nn.LeakyReLU(negative_slope=negative_slope, inplace=True) if leaky_relu else nn.ReLU(inplace=False)
        trk = use_running_statistics
        self.layer1 = _conv_block(num_channels, 64, act, True, trk)
        self.layer2 = _conv_block(64, 64, act, True, trk)
        self.layer3 = _conv_block(64, 64, act, False, trk)
        self.layer3_maxpool = nn.MaxPool2d(kernel_size=3, stride=3)
        self.layer4 = _conv_block(64, 64, act, False, trk)
        self.layer4_pool = nn.MaxPool2d(kernel_size=3, stride=3)
        flat = 64 * pooled_extent(self.IN_MELS, 4) * pooled_extent(self.IN_FRAMES, 4)
        self.logits = nn.Sequential(nn.Dropout(p=0.3), nn.BatchNorm1d(flat, eps=1e-05, momentum=0.1, affine=True),
                                    nn.Linear(in_features=flat, out_features=1600))

    # ------------------------------------------------------------------ inference path
    def _inference_ok(self, x):
        return (not self.training and not torch.is_grad_enabled() and x.is_cuda and not self.is_feature
                and x.dim() == 4 and x.shape[1] == 1 and x.dtype == torch.float32
                and self.layer1[1].track_running_stats and self.layer1[0].out_channels == 64)

    def _state_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    @staticmethod
    def _fold(conv, bn):
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = conv.weight * scale.view(-1, 1, 1, 1)
        b = (conv.bias - bn.running_mean) * scale + bn.bias
        return w, b

    def _folded(self):
        key = self._state_key()
        cache = getattr(self, "_fold_cache", None)
        if cache is not None and cache["key"] == key:
            return cache
        with torch.no_grad():
            w1, b1 = self._fold(self.layer1[0], self.layer1[1])
            cache = {"key": key,
                     "w1": w1.reshape(64, 9).float().cpu().numpy(), "b1": b1.float().cpu().numpy(),
                     "slope": float(getattr(self.layer1[2], "negative_slope", 0.0))}
            for i, layer in ((2, self.layer2), (3, self.layer3), (4, self.layer4)):
                w, b = self._fold(layer[0], layer[1])
                cache["w%d" % i] = w.contiguous(memory_format=torch.channels_last)
                cache["b%d" % i] = b.contiguous()
                if i < 4 and w.is_cuda and tuple(w.shape) == (64, 64, 3, 3):
                    cache["p%d" % i] = torch.from_numpy(ops.conv3x3_c64_pack_weights(w)).to(w.device)
            bn, lin = self.logits[1], self.logits[2]
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            cache["wl"] = (lin.weight * s.view(1, -1)).contiguous()
            cache["bl"] = (lin.bias + lin.weight @ (bn.bias - bn.running_mean * s)).contiguous()
        self._fold_cache = cache
        return cache

    @staticmethod
    def _conv_act(h, w, b, slope):
        if slope == 0.0:
            return torch.cudnn_convolution_relu(h, w, b, (1, 1), (1, 1), (1, 1), 1)
        return F.leaky_relu(F.conv2d(h, w, b, padding=1), slope, inplace=True)

    def _forward_inference(self, x):
        c = self._folded()
        # tcgen05 TF32 stem (csrc/conv1_tc.cu, 0.28 ms per 800 clips) when the caller allows TF32 convolutions --
        # PyTorch's and the reference's default, and the switch that governs the cuDNN blocks below -- otherwise
        # the exact-fp32 SIMT stem (csrc/conv1.cu, 0.36 ms).  stem_tf32 = True/False forces one of them.
        if self.precision == "bf16" and "p2" in c and "p3" in c:
            h = self._forward_blocks_bf16(x, c)
            if h is not None:
                return self._forward_tail(h, c)
        elif self.precision not in (None, "tf32", "bf16"):
            raise ValueError("Conv64F.precision must be None, 'tf32' or 'bf16'")
        tf32 = torch.backends.cudnn.allow_tf32 if self.stem_tf32 is None else self.stem_tf32
        h = ops.conv1_bn_act_pool3(x, c["w1"], c["b1"], c["slope"], tf32=tf32)  # [N,64,H/3,W/3] channels_last
        tc = torch.backends.cudnn.allow_tf32 if self.block_tc is None else self.block_tc
        if tc and "p2" in c and ops.conv3x3_c64_supported(h):
            h = ops.conv3x3_c64_bn_act(h, c["p2"], c["b2"], c["slope"], pool=True)
        else:
            h = ops.maxpool3_channels_last(self._conv_act(h, c["w2"], c["b2"], c["slope"]))
        if tc and "p3" in c and ops.conv3x3_c64_supported(h):
            h = ops.conv3x3_c64_bn_act(h, c["p3"], c["b3"], c["slope"], pool=self.maxpool_last2)
        else:
            h = self._conv_act(h, c["w3"], c["b3"], c["slope"])
            if self.maxpool_last2:
                h = ops.maxpool3_channels_last(h)
        return self._forward_tail(h, c)

    def _forward_blocks_bf16(self, x, c):
        """Blocks 1-3 with bf16 activations between them (csrc/conv1_tc.cu bf16 output, csrc/conv3_tc.cu bf16 MMAs);
        returns the fp32 channels_last input of block 4, or None when a shape is outside what the kernels are built for."""
        need = 27 if self.maxpool_last2 else 9  # every pooled block needs a 3x3 input at least
        if x.shape[2] < need or x.shape[3] < need or x.shape[3] // 3 > 61:
            return None
        for i in (2, 3):
            if "q%d" % i not in c:
                w = c["w%d" % i]
                c["q%d" % i] = torch.from_numpy(ops.conv3x3_c64_pack_weights_bf16(w)).to(w.device).view(torch.bfloat16)
        h = ops.conv1_bn_act_pool3(x, c["w1"], c["b1"], c["slope"], tf32=True, out_dtype=torch.bfloat16)
        h = ops.conv3x3_c64_bn_act_bf16(h, c["q2"], c["b2"], c["slope"], pool=True, out_dtype=torch.bfloat16)
        return ops.conv3x3_c64_bn_act_bf16(h, c["q3"], c["b3"], c["slope"], pool=self.maxpool_last2, out_dtype=torch.float32)

    def _forward_tail(self, h, c):
        h = self._conv_act(h, c["w4"], c["b4"], c["slope"])
        if self.last_pool and self.is_flatten and ops.pool3_linear_supported(h, c["wl"].shape[0]):
            return ops.pool3_linear(h, c["wl"], c["bl"])  # last max-pool + flatten + BatchNorm1d + Linear: one kernel
        if self.last_pool:
            h = ops.maxpool3_channels_last(h)
        if self.is_flatten:
            h = h.contiguous().view(h.size(0), -1)  # NCHW flatten order, as out4.view(N, -1)
            h = torch.addmm(c["bl"], h, c["wl"].t())
        return h

    def _train_stem_ok(self, x):
        """Training with batch statistics on CUDA, plain nn.BatchNorm2d (MAML's BatchStatNorm2d needs double
        backward through this block and keeps the module graph), 1-channel data input without grad."""
        conv, bn = self.layer1[0], self.layer1[1]
        return (self.fused_train_stem and self.training and torch.is_grad_enabled() and x.is_cuda and x.dim() == 4
                and x.shape[1] == 1 and x.dtype == torch.float32 and not x.requires_grad
                and type(bn) is nn.BatchNorm2d and bn.affine and bn.momentum is not None
                and type(conv) is nn.Conv2d and conv.out_channels == 64 and conv.bias is not None
                and conv.weight.dtype == torch.float32 and len(self.layer1) == 4)

    def forward(self, x):
        if self._inference_ok(x):
            return self._forward_inference(x)
        if self._train_stem_ok(x):
            out1 = ops.conv1_bn_act_pool3_train(x, self.layer1[0], self.layer1[1],
                                                float(getattr(self.layer1[2], "negative_slope", 0.0)))
        else:
            out1 = self.layer1(x)
        out2 = self.layer2(out1)
        out3 = self.layer3(out2)
        if self.maxpool_last2:
            out3 = self.layer3_maxpool(out3)
        out4 = self.layer4(out3)
        if self.last_pool:
            out4 = self.layer4_pool(out4)
        if self.is_flatten:
            out4 = self.logits(out4.view(out4.size(0), -1))
        if self.is_feature:
            return out1, out2, out3, out4
        return out4
