"""Parameter trees of the three backbones and the BPBreID head.

These nn.Module trees are *containers only*: they own the tensors, expose exactly the reference's
``state_dict`` key names (reference: torchreid/models/hrnet.py:299-378,
torchreid/models/resnet.py:160-243, torchreid/models/resnet_fastreid.py:243-263,
torchreid/models/bpbreid.py:18-114) so published checkpoints load unchanged, and initialise them with the
reference's distributions.  They have no ``forward``: the arithmetic runs in the sm_100a kernels driven by
``bpbreid_b200.plan`` (the container raises if somebody calls it).
"""
from __future__ import annotations

import math

import torch
from torch import nn

BN_MOMENTUM = 0.1


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the forward pass runs in bpbreid_b200 CUDA kernels")


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride, k // 2, bias=bias)


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


def _seq(*mods):
    return nn.Sequential(*mods)


class BasicUnit(_Holder):
    """two 3x3 convs + identity skip (hrnet.py:67-96)"""

    def __init__(self, c):
        super().__init__()
        self.conv1, self.bn1 = _conv(c, c, 3), _bn(c)
        self.conv2, self.bn2 = _conv(c, c, 3), _bn(c)


class BottleneckUnit(_Holder):
    """1x1 -> 3x3(stride) -> 1x1(x4) (+ 1x1 projection skip)  (hrnet.py:99-137, resnet.py:105-154)"""

    def __init__(self, cin, planes, stride=1, ibn=False):
        super().__init__()
        cout = planes * 4
        self.conv1 = _conv(cin, planes, 1)
        self.bn1 = IBNPair(planes) if ibn else _bn(planes)
        self.conv2, self.bn2 = _conv(planes, planes, 3, stride), _bn(planes)
        self.conv3, self.bn3 = _conv(planes, cout, 1), _bn(cout)
        self.stride = stride
        if stride != 1 or cin != cout:
            self.downsample = _seq(_conv(cin, cout, 1, stride), _bn(cout))
        else:
            self.downsample = None


class IBNPair(_Holder):
    """InstanceNorm(affine) on the first half of the channels, BatchNorm on the rest (resnet_fastreid.py:108-122)"""

    def __init__(self, planes):
        super().__init__()
        half = planes // 2
        self.half = half
        self.IN = nn.InstanceNorm2d(half, affine=True)
        self.BN = nn.BatchNorm2d(planes - half)


HR_WIDTHS = (32, 64, 128, 256)
HR_STAGES = ((2, 1), (3, 4), (4, 3))       # (branches, modules) of stage2/3/4 (hrnet.py:26-56)


class HRModuleParams(_Holder):
    """One HighResolutionModule: 4 BasicUnits per branch + the cross-resolution fuse convs (hrnet.py:140-279)."""

    def __init__(self, nb):
        super().__init__()
        w = HR_WIDTHS[:nb]
        self.branches = nn.ModuleList([_seq(*[BasicUnit(w[b]) for _ in range(4)]) for b in range(nb)])
        rows = []
        for i in range(nb):
            row = []
            for j in range(nb):
                if j == i:
                    row.append(None)
                elif j > i:      # 1x1 conv + BN, nearest-upsampled by 2^(j-i) at run time
                    row.append(_seq(_conv(w[j], w[i], 1), _bn(w[i])))
                else:            # (i-j) stride-2 3x3 convs; channels switch to w[i] on the last one
                    chain = []
                    for k in range(i - j):
                        co = w[i] if k == i - j - 1 else w[j]
                        chain.append(_seq(_conv(w[j], co, 3, 2), _bn(co)))
                    row.append(_seq(*chain))
            rows.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(rows)


class HRNet32Params(_Holder):
    feature_dim = 1920

    def __init__(self, dim_reduction_channels=512, **_):
        super().__init__()
        self.conv1, self.bn1 = _conv(3, 64, 3, 2), _bn(64)
        self.conv2, self.bn2 = _conv(64, 64, 3, 2), _bn(64)
        self.layer1 = _seq(BottleneckUnit(64, 64), *[BottleneckUnit(256, 64) for _ in range(3)])
        self.transition1 = nn.ModuleList([_seq(_conv(256, 32, 3), _bn(32)),
                                          _seq(_seq(_conv(256, 64, 3, 2), _bn(64)))])
        self.stage2 = _seq(*[HRModuleParams(2) for _ in range(1)])
        self.transition2 = nn.ModuleList([None, None, _seq(_seq(_conv(64, 128, 3, 2), _bn(128)))])
        self.stage3 = _seq(*[HRModuleParams(3) for _ in range(4)])
        self.transition3 = nn.ModuleList([None, None, None, _seq(_seq(_conv(128, 256, 3, 2), _bn(256)))])
        self.stage4 = _seq(*[HRModuleParams(4) for _ in range(3)])
        self.incre_modules = nn.ModuleList([_seq(BottleneckUnit(HR_WIDTHS[i], HR_WIDTHS[i])) for i in range(4)])
        # built by the reference but only run for dim_reduce='before_pooling' (hrnet.py:362-371,574-575)
        self.cls_head = _seq(nn.Conv2d(1920, dim_reduction_channels, 1), _bn(dim_reduction_channels))
        for m in self.modules():                                  # hrnet.py:578-586
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class ResNet50Params(_Holder):
    feature_dim = 2048

    def __init__(self, num_classes=751, fastreid_ibn=False, **_):
        super().__init__()
        self.fastreid = fastreid_ibn
        self.conv1, self.bn1 = _conv(3, 64, 7, 2), nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, n, stride) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 1))):
            ibn = fastreid_ibn and li < 3                            # resnet_fastreid.py:252-255
            blocks = [BottleneckUnit(cin, planes, stride, ibn)]
            blocks += [BottleneckUnit(planes * 4, planes, 1, ibn) for _ in range(n - 1)]
            setattr(self, f"layer{li + 1}", _seq(*blocks))
            cin = planes * 4
        if not fastreid_ibn:
            self.classifier = nn.Linear(2048, num_classes)          # unused by the part-based path (resnet.py:243)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):                               # resnet.py:328-333 / resnet_fastreid.py:347-351
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                m.momentum = 0.1


class DimReduceParams(_Holder):
    """Linear(C->D)+BN1d(+ReLU)  (bpbreid.py:324-350)"""

    def __init__(self, cin, cout):
        super().__init__()
        self.layers = _seq(nn.Linear(cin, cout), nn.BatchNorm1d(cout))
        nn.init.normal_(self.layers[0].weight, 0, 0.01)
        nn.init.constant_(self.layers[0].bias, 0)


class BNClassifierParams(_Holder):
    """BN1d (bias frozen) + bias-free Linear (bpbreid.py:398-415)"""

    def __init__(self, cin, ncls):
        super().__init__()
        self.bn = nn.BatchNorm1d(cin)
        self.bn.bias.requires_grad_(False)
        self.classifier = nn.Linear(cin, ncls, bias=False)
        nn.init.normal_(self.classifier.weight, 0, 0.001)


class PixelClassifierParams(_Holder):
    """BN2d + 1x1 conv C -> K+1 (bpbreid.py:376-396)"""

    def __init__(self, cin, parts_num):
        super().__init__()
        self.bn = nn.BatchNorm2d(cin)
        self.classifier = nn.Conv2d(cin, parts_num + 1, 1)
        nn.init.normal_(self.classifier.weight, 0, 0.001)
        nn.init.constant_(self.classifier.bias, 0)


BACKBONES = {
    "hrnet32": lambda num_classes, **kw: HRNet32Params(**kw),
    "resnet50": lambda num_classes, **kw: ResNet50Params(num_classes, False),
    "fastreid_resnet_ibn": lambda num_classes, **kw: ResNet50Params(num_classes, True),
}


def seeded_state_dict(module: nn.Module, seed: int):
    """Deterministic re-draw of every floating tensor of `module` (used by tests/goldens so that the same
    weights can be regenerated on the GPU box without shipping 160 MB fixtures): conv/linear weights keep
    their init scale, BN affine/running stats are perturbed so that eval-mode BN is exercised too."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    out = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            out[k] = v.clone()
            continue
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "running_mean":
            t = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf == "running_var":
            t = 0.5 + torch.rand(v.shape, generator=g)
        elif v.dim() == 1 and leaf == "weight":        # norm scale
            t = 0.75 + 0.5 * torch.rand(v.shape, generator=g)
        elif v.dim() == 1:                              # biases (norm shift, linear/conv bias)
            t = 0.1 * torch.randn(v.shape, generator=g)
        else:
            fan = v[0].numel() if v.dim() > 1 else v.numel()
            std = float(v.std()) if v.numel() > 1 and float(v.std()) > 0 else math.sqrt(2.0 / fan)
            t = std * torch.randn(v.shape, generator=g)
        out[k] = t.to(v.dtype)
    return out
