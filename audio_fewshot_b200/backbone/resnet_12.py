"""ResNet-12 (64/160/320/640, LeakyReLU 0.1) with the reference's parameter names.

Architecture and state_dict layout of libfewshot_core/model/backbone/resnet_12.py:26-301
(layerK.0.conv{1,2,3}, layerK.0.bn{1,2,3}, layerK.0.downsample.{0,1}).  DropBlock is the
reference's regulariser for training (resnet_12.py:83-99); in eval it is the identity.
Here it is implemented with device-agnostic torch ops (the reference's version calls
.cuda() unconditionally, dropblock.py:28-29).  Convolutions stay on cuDNN.
"""
import torch
import torch.nn.functional as F
from torch import nn


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


class ResNet(nn.Module):
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
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if self.keep_avg_pool:
            x = self.avgpool(x)
        if self.is_flatten:
            x = x.view(x.size(0), -1)
        return x


def resnet12(keep_prob=1.0, avg_pool=True, is_flatten=True, maxpool_last2=True, **kwargs):
    return ResNet(keep_prob=keep_prob, avg_pool=avg_pool, is_flatten=is_flatten, maxpool_last2=maxpool_last2,
                  **kwargs)
