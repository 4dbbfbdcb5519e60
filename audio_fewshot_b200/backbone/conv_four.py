"""Conv64F feature extractor, audio variant (3x3/stride-3 max-pools, 64->1600 head).

Same architecture, constructor kwargs and state_dict names as the reference's
libfewshot_core/model/backbone/conv_four.py:28-128, so `*_best.pth` checkpoints load
unchanged (layer{1..4}.0 = Conv2d with bias, layer{1..4}.1 = BatchNorm2d,
logits.1 = BatchNorm1d(64), logits.2 = Linear(64, 1600)).  The convolutions stay on
cuDNN: the north star names no backbone kernel.
"""
import torch
from torch import nn


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

    def forward(self, x):
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
