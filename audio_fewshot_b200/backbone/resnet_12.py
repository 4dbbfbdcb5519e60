"""ResNet-12 (64/160/320/640, LeakyReLU 0.1) with the reference's parameter names.

Architecture and state_dict layout of libfewshot_core/model/backbone/resnet_12.py:26-301
(layerK.0.conv{1,2,3}, layerK.0.bn{1,2,3}, layerK.0.downsample.{0,1}).  DropBlock is the
reference's regulariser for training (resnet_12.py:83-99); in eval it is the identity.
Here it is implemented with device-agnostic torch ops (the reference's version calls
.cuda() unconditionally, dropblock.py:28-29).  Convolutions stay on cuDNN.

Inference path (eval mode, grad disabled, CUDA): BatchNorms are folded into the convolution weights, tensors stay
channels-last, the bias + LeakyReLU after conv1/conv2 is one in-place sm_100a kernel and the block tail
(bn3 -> += residual -> LeakyReLU -> MaxPool2d) is ONE kernel (csrc/pool.cu: add_bias_act_pool) that reads the two
convolution outputs once and writes the pooled map -- instead of the module graph's BatchNorm, add, activation
and pooling passes over the un-pooled activation.  Training / autograd keeps the reference's op sequence.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


def fold_conv_bn(conv, bn):
    """Eval-mode BatchNorm folded into the preceding bias-free convolution -> (weight channels_last, bias)."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = (conv.weight * scale.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
    b = bn.bias - bn.running_mean * scale
    if conv.bias is not None:
        b = b + conv.bias * scale
    return w, b.contiguous()


class FoldedTrunkMixin:
    """Shared by ResNet and ResNetBdc: the folded channels-last evaluation of layer1..layer4."""

    fast_eval = True
    # "bf16": the separately stated reduced-precision inference trunk: bf16 channels-last cuDNN
    # convolutions with the BatchNorms folded in, the block tails as csrc/pool.cu's bf16 kernel (fp32 arithmetic), the
    # last block's output in fp32 for the heads.  None / "tf32": the reference's precision class (the parity path).
    precision = None

    def _inference_ok(self, x):
        return (self.fast_eval and not self.training and not torch.is_grad_enabled() and x.is_cuda
                and x.dim() == 4 and x.dtype == torch.float32)

    def _trunk_state_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _folded_trunk(self):
        key = self._trunk_state_key()
        cache = getattr(self, "_trunk_cache", None)
        if cache is not None and cache["key"] == key:
            return cache["blocks"]
        blocks = []
        with torch.no_grad():
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                blk = layer[0]
                w1, b1 = fold_conv_bn(blk.conv1, blk.bn1)
                w2, b2 = fold_conv_bn(blk.conv2, blk.bn2)
                w3, b3 = fold_conv_bn(blk.conv3, blk.bn3)
                wd = None
                if blk.downsample is not None:
                    wd, bd = fold_conv_bn(blk.downsample[0], blk.downsample[1])
                    b3 = (b3 + bd).contiguous()
                blocks.append((w1, b1, w2, b2, w3, b3, wd, float(blk.relu.negative_slope),
                               int(blk.stride) if blk.use_pool else 1))
        self._trunk_cache = {"key": key, "blocks": blocks}
        return blocks

    def _folded_trunk_bf16(self):
        blocks = self._folded_trunk()
        cache = self._trunk_cache
        if "bf16" not in cache:
            to16 = lambda w: None if w is None else w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            cache["bf16"] = [(to16(w1), b1, to16(w2), b2, to16(w3), b3, to16(wd), slope, k)
                             for (w1, b1, w2, b2, w3, b3, wd, slope, k) in blocks]
        return cache["bf16"]

    def _trunk_inference_bf16(self, x):
        x = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        if x.shape[1] == 1:
            n, c, h, w = x.shape
            x = x.as_strided((n, c, h, w), (h * w, 1, w, 1))
        layers = (self.layer1, self.layer2, self.layer3, self.layer4)
        for i, (layer, (w1, b1, w2, b2, w3, b3, wd, slope, k)) in enumerate(zip(layers, self._folded_trunk_bf16())):
            layer[0].num_batches_tracked += 1
            o = ops.add_bias_act_pool_bf16(F.conv2d(x, w1, padding=1), None, b1, slope, 1, inplace=True)
            o = ops.add_bias_act_pool_bf16(F.conv2d(o, w2, padding=1), None, b2, slope, 1, inplace=True)
            o = F.conv2d(o, w3, padding=1)
            r = x if wd is None else F.conv2d(x, wd)
            x = ops.add_bias_act_pool_bf16(o, r, b3, slope, k,
                                           out_dtype=torch.float32 if i == len(layers) - 1 else torch.bfloat16)
        return x

    def _trunk_inference(self, x):
        if self.precision == "bf16":
            return self._trunk_inference_bf16(x)
        if self.precision not in (None, "tf32"):
            raise ValueError("precision must be None, 'tf32' or 'bf16'")
        x = x.contiguous(memory_format=torch.channels_last)
        if x.shape[1] == 1:
            # a 1-channel tensor is both NCHW- and NHWC-contiguous; give it unambiguous channels-last strides so
            # cuDNN's first convolution writes channels-last output (PyTorch picks the format from the strides)
            n, c, h, w = x.shape
            x = x.as_strided((n, c, h, w), (h * w, 1, w, 1))
        for layer, (w1, b1, w2, b2, w3, b3, wd, slope, k) in zip((self.layer1, self.layer2, self.layer3, self.layer4),
                                                                  self._folded_trunk()):
            layer[0].num_batches_tracked += 1  # the reference counts every forward (resnet_12.py:73)
            o = ops.add_bias_act_pool(F.conv2d(x, w1, padding=1), None, b1, slope, 1, inplace=True)
            o = ops.add_bias_act_pool(F.conv2d(o, w2, padding=1), None, b2, slope, 1, inplace=True)
            o = F.conv2d(o, w3, padding=1)
            r = x if wd is None else F.conv2d(x, wd)
            x = ops.add_bias_act_pool(o, r, b3, slope, k)
        return x


def drop_block(x, gamma, block_size, training):
    """DropBlock (reference backbone/utils/dropblock.py:14-43): sample block centres with
    probability gamma, zero block_size x block_size squares, renormalise by the kept fraction."""
    if not training:
        return x
    b, c, h, w = x.shape
    mh, mw = h - (block_size - 1), w - (block_size - 1)
    if mh <= 0 or mw <= 0:
        return x
    seeds = (torch.rand(b, c, mh, mw, device=x.device) < gamma).to(x.dtype)
    # a seed at (i, j) zeroes rows i..i+bs-1, cols j..j+bs-1 of the h x w map (dropblock.py:45-80)
    pad = block_size - 1
    dropped = F.max_pool2d(F.pad(seeds, (pad, pad, pad, pad)), kernel_size=block_size, stride=1)
    mask = 1.0 - dropped
    kept = mask.sum()
    return mask * x * (mask.numel() / kept.clamp_min(1.0))


def conv3x3(c_in, c_out):
    return nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, drop_rate=0.0, drop_block=False,
                 block_size=1, use_pool=True):
        super().__init__()
        self.conv1, self.bn1 = conv3x3(inplanes, planes), nn.BatchNorm2d(planes)
        self.relu = nn.LeakyReLU(0.1)
        self.conv2, self.bn2 = conv3x3(planes, planes), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = conv3x3(planes, planes), nn.BatchNorm2d(planes)
        self.maxpool = nn.MaxPool2d(stride)
        self.downsample = downsample
        self.stride, self.drop_rate = stride, drop_rate
        self.num_batches_tracked = 0
        self.drop_block, self.block_size, self.use_pool = drop_block, block_size, use_pool

    def forward(self, x):
        self.num_batches_tracked += 1
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        residual = x if self.downsample is None else self.downsample(x)
        out = self.relu(out + residual)
        if self.use_pool:
            out = self.maxpool(out)
        if self.drop_rate > 0:
            if self.drop_block:
                feat = out.size(2)
                keep = max(1.0 - self.drop_rate / (20 * 2000) * self.num_batches_tracked, 1.0 - self.drop_rate)
                gamma = (1 - keep) / self.block_size ** 2 * feat ** 2 / (feat - self.block_size + 1) ** 2
                out = drop_block(out, gamma, self.block_size, self.training)
            else:
                out = F.dropout(out, p=self.drop_rate, training=self.training, inplace=True)
        return out


def make_stage(block, inplanes, planes, stride, drop_rate, drop_block=False, block_size=1, **extra):
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, kernel_size=1, stride=1, bias=False),
                                   nn.BatchNorm2d(planes * block.expansion))
    return nn.Sequential(block(inplanes, planes, stride, downsample, drop_rate, drop_block, block_size, **extra))


def init_resnet(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="leaky_relu")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class ResNet(FoldedTrunkMixin, nn.Module):
    def __init__(self, planes=(64, 160, 320, 640), keep_prob=1.0, avg_pool=True, drop_rate=0.1,
                 dropblock_size=5, is_flatten=True, maxpool_last2=True, num_channels=3):
        super().__init__()
        self.layer1 = make_stage(BasicBlock, num_channels, planes[0], 2, drop_rate)
        self.layer2 = make_stage(BasicBlock, planes[0], planes[1], 2, drop_rate)
        self.layer3 = make_stage(BasicBlock, planes[1], planes[2], 2, drop_rate, True, dropblock_size,
                                 use_pool=maxpool_last2)
        self.layer4 = make_stage(BasicBlock, planes[2], planes[3], 2, drop_rate, True, dropblock_size,
                                 use_pool=maxpool_last2)
        if avg_pool:
            self.avgpool = nn.AvgPool2d(5, stride=1)
        self.keep_prob, self.keep_avg_pool = keep_prob, avg_pool
        self.dropout = nn.Dropout(p=1 - keep_prob, inplace=False)
        self.drop_rate, self.is_flatten = drop_rate, is_flatten
        init_resnet(self)

    def forward(self, x):
        if self._inference_ok(x):
            x = self._trunk_inference(x)
        else:
            x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if self.keep_avg_pool:
            x = self.avgpool(x)
        if self.is_flatten:
            x = x.contiguous().view(x.size(0), -1)  # NCHW flatten order (resnet_12.py:285)
        return x


def resnet12(keep_prob=1.0, avg_pool=True, is_flatten=True, maxpool_last2=True, **kwargs):
    return ResNet(keep_prob=keep_prob, avg_pool=avg_pool, is_flatten=is_flatten, maxpool_last2=maxpool_last2,
                  **kwargs)
