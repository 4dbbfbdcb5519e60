"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Functional CPU restatement (eval mode, from a reference-format state_dict) of the
three backbones on the BASELINE configs.  Used (a) to check that the product's
nn.Modules load the reference's parameter names and compute the same features,
and (b) as the `--impl reference` / cpu_baseline arm of bench.py on the GPU box,
where /root/reference does not exist.  Validated against the real reference
classes in tests/test_oracle.py when /root/reference is present.
"""
import torch
import torch.nn.functional as F


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.1, 1e-5)


def conv64f_forward(sd, x, is_flatten=True, last_pool=True, maxpool_last2=True,
                    leaky_relu=False, negative_slope=0.2):
    """reference: Conv64F.forward, libfewshot_core/model/backbone/conv_four.py:99-128
    (layers :61-92; eval mode, so Dropout(0.3) at :88 is the identity)."""
    act = (lambda t: F.leaky_relu(t, negative_slope)) if leaky_relu else F.relu
    out = x
    for i in (1, 2):  # :61-72  conv-bn-act-maxpool(3,3)
        out = F.conv2d(out, sd["layer%d.0.weight" % i], sd["layer%d.0.bias" % i], padding=1)
        out = F.max_pool2d(act(_bn(out, sd, "layer%d.1" % i)), 3, 3)
    out = act(_bn(F.conv2d(out, sd["layer3.0.weight"], sd["layer3.0.bias"], padding=1), sd, "layer3.1"))
    if maxpool_last2:  # :112-114
        out = F.max_pool2d(out, 3, 3)
    out = act(_bn(F.conv2d(out, sd["layer4.0.weight"], sd["layer4.0.bias"], padding=1), sd, "layer4.1"))
    if last_pool:  # :116-118
        out = F.max_pool2d(out, 3, 3)
    if is_flatten:  # :120-122  view -> Dropout -> BatchNorm1d -> Linear
        out = out.view(out.size(0), -1)
        out = F.batch_norm(out, sd["logits.1.running_mean"], sd["logits.1.running_var"],
                           sd["logits.1.weight"], sd["logits.1.bias"], False, 0.1, 1e-5)
        out = F.linear(out, sd["logits.2.weight"], sd["logits.2.bias"])
    return out


def _basic_block(sd, p, x, stride, use_pool=True):
    """reference: BasicBlock.forward, libfewshot_core/model/backbone/resnet_12.py:58-101
    (eval: DropBlock / dropout are identities)."""
    out = F.leaky_relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], padding=1), sd, p + ".bn1"), 0.1)
    out = F.leaky_relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2"), 0.1)
    out = _bn(F.conv2d(out, sd[p + ".conv3.weight"], padding=1), sd, p + ".bn3")
    residual = x
    if (p + ".downsample.0.weight") in sd:
        residual = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"]), sd, p + ".downsample.1")
    out = F.leaky_relu(out + residual, 0.1)
    if use_pool:
        out = F.max_pool2d(out, stride)
    return out


def resnet12_forward(sd, x, avg_pool=True, is_flatten=True, maxpool_last2=True):
    """reference: ResNet.forward, resnet_12.py:276-286 (layers :190-215)."""
    x = _basic_block(sd, "layer1.0", x, 2)
    x = _basic_block(sd, "layer2.0", x, 2)
    x = _basic_block(sd, "layer3.0", x, 2, use_pool=maxpool_last2)
    x = _basic_block(sd, "layer4.0", x, 2, use_pool=maxpool_last2)
    if avg_pool:
        x = F.avg_pool2d(x, 5, stride=1)
    if is_flatten:
        x = x.view(x.size(0), -1)
    return x


def resnet12bdc_trunk(sd, x):
    """reference: resnet.forward up to bdc_pool, libfewshot_core/model/backbone/resnet_bdc.py:345-350
    (BasicBlockVariant :224-281; layer4 has stride 1, i.e. MaxPool2d(1))."""
    x = _basic_block(sd, "layer1.0", x, 2)
    x = _basic_block(sd, "layer2.0", x, 2)
    x = _basic_block(sd, "layer3.0", x, 2)
    x = _basic_block(sd, "layer4.0", x, 1)
    return x


def bdc_reduce(sd, x):
    """reference: BdcPool.conv_dr_block, bdc_pool.py:34-38,59-60 (1x1 conv, BN, ReLU)."""
    if "bdc_pool.conv_dr_block.0.weight" in sd:
        x = F.relu(_bn(F.conv2d(x, sd["bdc_pool.conv_dr_block.0.weight"]), sd, "bdc_pool.conv_dr_block.1"))
    return x


def resnet12bdc_forward(sd, x):
    from .heads import bdcovpool, triuvec

    x = bdc_reduce(sd, resnet12bdc_trunk(sd, x))
    return triuvec(bdcovpool(x, sd["bdc_pool.temperature"]))
