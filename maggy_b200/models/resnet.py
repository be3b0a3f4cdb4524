"""ResNet-50 (He et al. 2015) for the ASHA HPO benchmark config ("ASHA HPO ResNet-50 synthetic, 8 concurrent
trials on 8xB200, 1 GPU/trial", BASELINE.json). channels_last bf16; every convolution runs on the maggy_b200
tcgen05 kernels (``ops.conv.Conv2d``: implicit GEMM for the 3x3 stride-1 layers, the plain GEMM for 1x1, im2col +
GEMM for the strided ones); ``conv_impl="torch"`` builds the cuDNN model used as the comparison arm of
``bench/hpo_asha_resnet.py``. With the native convolutions every BatchNorm is the fused NHWC kernel set of
``ops.batchnorm`` (statistics -> scale/shift -> apply + ReLU (+ residual add) in one pass; three streaming kernels in
backward) — ATen's channels-last BatchNorm kernels were 56 % of the step (profiles/resnet_step_r2.md);
``MAGGY_B200_BN=aten`` keeps them for A/B runs. The 3x3/2 max pool is ``ops.pool`` (gather backward, no atomics);
global average pooling and the 2048 -> classes head stay on ATen.
"""
from __future__ import annotations

import torch
import torch.nn as nn

import os

from ..ops.batchnorm import BatchNorm2d as _NativeBatchNorm2d
from ..ops.conv import Conv2d as _NativeConv2d
from ..ops.pool import MaxPool2d as _NativeMaxPool2d


def _conv(impl, cin, cout, k, stride=1, padding=0):
    if impl == "torch":
        return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=False)
    return _NativeConv2d(cin, cout, k, stride=stride, padding=padding)


def _fused_norm(impl) -> bool:
    return impl != "torch" and os.environ.get("MAGGY_B200_BN", "native") != "aten"


def _bn(impl, planes, relu=False):
    """(module, fused): the fused module applies ReLU (and the residual add passed to forward) itself."""
    if _fused_norm(impl):
        return _NativeBatchNorm2d(planes, relu=relu)
    return nn.BatchNorm2d(planes)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, conv_impl="native"):
        super().__init__()
        self.fused = _fused_norm(conv_impl)
        self.conv1 = _conv(conv_impl, inplanes, planes, 1)
        self.bn1 = _bn(conv_impl, planes, relu=True)
        self.conv2 = _conv(conv_impl, planes, planes, 3, stride=stride, padding=1)
        self.bn2 = _bn(conv_impl, planes, relu=True)
        self.conv3 = _conv(conv_impl, planes, planes * 4, 1)
        self.bn3 = _bn(conv_impl, planes * 4, relu=True)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        if self.fused:
            # x_skip aliases x: the skip / downsample gradient comes back through conv1's backward and is added in its
            # dgrad epilogue instead of by an autograd add kernel
            out, x_skip = self.conv1(x, with_skip=True)
            idt = x_skip if self.downsample is None else self.downsample(x_skip)
            out = self.bn1(out)                      # norm + ReLU in one kernel
            out = self.bn2(self.conv2(out))
            return self.bn3(self.conv3(out), idt)    # norm + skip add + ReLU in one kernel
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + idt)


class ResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000, width=64, dropout=0.0, conv_impl="native"):
        super().__init__()
        self.inplanes = width
        self.conv_impl = conv_impl
        self.fused = _fused_norm(conv_impl)
        self.conv1 = _conv(conv_impl, 3, width, 7, stride=2, padding=3)
        self.bn1 = _bn(conv_impl, width, relu=True)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = (_NativeMaxPool2d if self.fused else nn.MaxPool2d)(3, stride=2, padding=1)
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.drop = nn.Dropout(dropout) if dropout > 0 else nn.Identity()
        self.fc = nn.Linear(width * 8 * 4, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, planes, blocks, stride=1):
        down = None
        if stride != 1 or self.inplanes != planes * 4:
            down = nn.Sequential(_conv(self.conv_impl, self.inplanes, planes * 4, 1, stride=stride),
                                 _bn(self.conv_impl, planes * 4))
        layers = [Bottleneck(self.inplanes, planes, stride, down, self.conv_impl)]
        self.inplanes = planes * 4
        layers += [Bottleneck(self.inplanes, planes, conv_impl=self.conv_impl) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.bn1(self.conv1(x))
        x = self.maxpool(x if self.fused else self.relu(x))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.avgpool(x), 1)
        return self.fc(self.drop(x))


def resnet50(num_classes=1000, dropout=0.0, conv_impl="native") -> ResNet:
    return ResNet((3, 4, 6, 3), num_classes, dropout=dropout, conv_impl=conv_impl)


def resnet18ish(num_classes=10, conv_impl="native") -> ResNet:
    """A small bottleneck ResNet for CPU tests."""
    return ResNet((1, 1, 1, 1), num_classes, width=16, conv_impl=conv_impl)
