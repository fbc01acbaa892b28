"""Parameter-compatible mirrors of the other nn.Modules on the head path:

  SKNet / SKBlock   lib/model/modules/blocks_coatt_transformer_sk.py:915-998
  Bottleneck, make_layer4 (= `RCNN_top`)   lib/model/faster_rcnn/resnet_coatt_transformer_sk.py:73-109,416

Like system/Models.py they only own parameters under the reference's state_dict keys; their
forward runs in libaitb200 (grouped / strided / 3x3 convolutions as TMA-shifted tcgen05 GEMMs
with frozen BatchNorm folded into the packed weights).
"""
import math

import torch
import torch.nn as nn

from . import packing


class SKBlock(nn.Module):
    """relu(conv1x1_g8(x))**2 + relu(conv3x3_g8(x))**2 -- the reference's forward squares the branch
    outputs and discards the selective-kernel attention (blocks_coatt_transformer_sk.py:973-984);
    `fc` / `sk` exist only so checkpoints load."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        kernels = [1, 3]
        self.n_state = len(kernels)
        self.convs = nn.ModuleList([
            nn.Sequential(nn.Conv2d(channels, channels, kernel_size=k, stride=1, padding=k // 2, groups=8),
                          nn.ReLU(inplace=True)) for k in kernels])
        self.fc = nn.Linear(channels, channels // reduction)
        self.sk = nn.Linear(channels // reduction, channels * self.n_state)
        for m in self.modules():  # reset_params(): kaiming fan_out on convs, zero bias
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)


class SKNet(nn.Module):
    def __init__(self, channels, reduction=16, compute_dtype=torch.float32):
        super().__init__()
        if channels != 1024:
            raise NotImplementedError("ait_b200.SKNet supports channels=1024 (ResNet-50 C4) only")
        self.sk_props = SKBlock(channels, reduction)
        self.sk_query = SKBlock(channels, reduction)
        self.compute_dtype = compute_dtype
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._engine = None

    def forward(self, x_props, x_query):
        """x_props [bp,1024,8,8], x_query [bs,1024,8,8] -> same shapes (blocks_...sk.py:993-998)."""
        fp = packing.fingerprint(self)          # in-place parameter updates since the last packing?
        if self._engine is None or self._engine_fp != fp:
            self._engine_fp = fp
            self._engine = packing.HeadEngine(sk=self, dtype=self.compute_dtype)
        return self._engine.sk_forward(x_props, x_query)


class Bottleneck(nn.Module):
    """Caffe-style bottleneck: the stride sits on the first 1x1 conv (resnet_coatt...:78)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride


def make_layer4():
    """ResNet-50 layer4 with the reference's ResNet init (resnet_coatt...:130-136)."""
    down = nn.Sequential(nn.Conv2d(1024, 2048, kernel_size=1, stride=2, bias=False), nn.BatchNorm2d(2048))
    layer = nn.Sequential(Bottleneck(1024, 512, 2, down), Bottleneck(2048, 512), Bottleneck(2048, 512))
    for m in layer.modules():
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()
    return layer
