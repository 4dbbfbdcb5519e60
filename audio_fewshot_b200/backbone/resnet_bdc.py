"""resnet12Bdc: ResNet-12 trunk (layer4 stride 1) + BdcPool, reference parameter names.

libfewshot_core/model/backbone/resnet_bdc.py:283-358 and backbone/utils/bdc_pool.py:19-66.
The trunk stays on cuDNN; the BDC matrix (BDCovpool + Triuvec, bdc_pool.py:69-93: 7 bmm +
a CPU-built gather index per call) is one sm_100a kernel (csrc/bdc.cu) on CUDA tensors.
"""
import math

import torch
from torch import nn

from .. import ops
from .resnet_12 import BasicBlock, FoldedTrunkMixin, fold_conv_bn, init_resnet, make_stage


class BdcPool(nn.Module):
    def __init__(self, is_vec=True, input_dim=(640, 10, 10), dimension_reduction=None, activate="relu"):
        super().__init__()
        self.is_vec, self.dr, self.activate = is_vec, dimension_reduction, activate
        self.input_dim = input_dim[0]
        if self.dr is not None and self.dr != self.input_dim:
            self.act = nn.LeakyReLU(0.1) if activate == "leaky_relu" else nn.ReLU(inplace=True)
            self.conv_dr_block = nn.Sequential(
                nn.Conv2d(self.input_dim, self.dr, kernel_size=1, stride=1, bias=False),
                nn.BatchNorm2d(self.dr), self.act)
        out = self.dr if self.dr else self.input_dim
        self.output_dim = out * (out + 1) // 2 if is_vec else out * out
        # bdc_pool.py:46 -- log(1 / (2 h w)) from the DECLARED feat_dim, not the real map
        self.temperature = nn.Parameter(
            torch.log((1.0 / (2 * input_dim[1] * input_dim[2])) * torch.ones(1, 1)), requires_grad=True)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, a=0, mode="fan_out", nonlinearity="leaky_relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        if self.dr is not None and self.dr != self.input_dim:
            if (not self.training and not torch.is_grad_enabled() and x.is_cuda
                    and x.is_contiguous(memory_format=torch.channels_last) and self.dr % 4 == 0):
                # inference: 1x1 conv with the BatchNorm folded in, bias + activation in one in-place kernel
                w, b = fold_conv_bn(self.conv_dr_block[0], self.conv_dr_block[1])
                slope = float(getattr(self.act, "negative_slope", 0.0))
                x = ops.add_bias_act_pool(torch.nn.functional.conv2d(x, w), None, b, slope, 1, inplace=True)
            else:
                x = self.conv_dr_block(x)
        return ops.bdc_pool(x, self.temperature, triu=self.is_vec)


class ResNetBdc(FoldedTrunkMixin, nn.Module):
    def __init__(self, keep_prob=1.0, avg_pool=False, drop_rate=0.0, dropblock_size=5, num_classes=-1,
                 use_se=False, reduce_dim=640, num_channels=3):
        super().__init__()
        if use_se:
            raise NotImplementedError("squeeze-excitation variant is outside the BASELINE configs")
        self.layer1 = make_stage(BasicBlock, num_channels, 64, 2, drop_rate)
        self.layer2 = make_stage(BasicBlock, 64, 160, 2, drop_rate)
        self.layer3 = make_stage(BasicBlock, 160, 320, 2, drop_rate, True, dropblock_size)
        self.layer4 = make_stage(BasicBlock, 320, 640, 1, drop_rate, True, dropblock_size)
        self.keep_prob, self.keep_avg_pool = keep_prob, avg_pool
        self.dropout = nn.Dropout(p=1 - keep_prob, inplace=False)
        self.drop_rate = drop_rate
        self.feat_dim = [640, 10, 10]
        init_resnet(self)
        self.bdc_pool = BdcPool(is_vec=True, input_dim=self.feat_dim, dimension_reduction=reduce_dim)
        self.num_classes = num_classes
        if num_classes > 0:
            self.classifier = nn.Linear(640, num_classes)

    def forward(self, x, is_feat=False):
        if self._inference_ok(x):
            x = self._trunk_inference(x)
        else:
            x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.bdc_pool(x)


def resnet12Bdc(keep_prob=1.0, avg_pool=True, **kwargs):
    return ResNetBdc(keep_prob=keep_prob, avg_pool=avg_pool, **kwargs)
