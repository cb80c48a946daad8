"""ORACLE (test infrastructure, never shipped or measured as the product).

Pure-PyTorch restatements of the two timm trunks the reference scorer wraps, with
timm-compatible *module names* so that the reference's hook lists resolve:

* ``resnet50``  (ImageNet ResNet-50 v1.5) -- taps ``layer{k}.2.act3``
  (/root/reference/models/global_eval_models.py:701).  Architecture follows the
  torchvision definition vendored in the reference at
  /root/reference/additional_approaches/src/transalnet/utils/resnet.py:106-161
  (Bottleneck) and :164-283 (ResNet); state_dict keys are identical to it.
* ``resnet50_clip.openai`` (OpenAI CLIP ModifiedResNet-50 as wrapped by timm's ByobNet)
  -- taps ``stages.{s}.2.act`` (/root/reference/models/global_eval_models.py:327).
  timm is NOT vendored, pinned or installed, so this follows the public CLIP
  ModifiedResNet design: 3-conv stem + avg-pool, anti-aliased (avg-pool) strides,
  avg-pool + 1x1 shortcut, attention-pool head.  **parity unpinned** at the timm
  boundary (parameter key names are from memory of timm >= 1.0).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# ImageNet ResNet-50 (timm naming: act1/act2/act3 per block, downsample = Sequential)
# --------------------------------------------------------------------------------------


class TimmBottleneck(nn.Module):
    """resnet.py:106-161 of the vendored torchvision file, renamed the timm way."""

    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: nn.Module | None = None):
        super().__init__()
        width = planes
        outplanes = planes * self.expansion
        self.conv1 = nn.Conv2d(inplanes, width, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.act1 = nn.ReLU(inplace=True)
        # v1.5: the stride sits on the 3x3 (resnet.py:107-111, :135)
        self.conv2 = nn.Conv2d(width, width, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.act2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(width, outplanes, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(outplanes)
        self.act3 = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        shortcut = x
        x = self.act1(self.bn1(self.conv1(x)))
        x = self.act2(self.bn2(self.conv2(x)))
        x = self.bn3(self.conv3(x))
        if self.downsample is not None:
            shortcut = self.downsample(shortcut)
        x = x + shortcut
        return self.act3(x)


class TimmResNet50(nn.Module):
    """ImageNet ResNet-50; resnet.py:164-283 (layers [3,4,6,3], :326-334)."""

    default_cfg = {
        "input_size": (3, 224, 224),
        "interpolation": "bicubic",
        "mean": (0.485, 0.456, 0.406),
        "std": (0.229, 0.224, 0.225),
        "crop_pct": 0.95,
        "crop_mode": "center",
    }

    def __init__(self, num_classes: int = 1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.act1 = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(64, 3, 1)
        self.layer2 = self._make_layer(128, 4, 2)
        self.layer3 = self._make_layer(256, 6, 2)
        self.layer4 = self._make_layer(512, 3, 2)
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(2048, num_classes)

    def _make_layer(self, planes: int, blocks: int, stride: int) -> nn.Sequential:
        downsample = None
        if stride != 1 or self.inplanes != planes * 4:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * 4, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * 4),
            )
        layers = [TimmBottleneck(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * 4
        for _ in range(1, blocks):
            layers.append(TimmBottleneck(self.inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(self.act1(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.global_pool(x), 1)
        return self.fc(x)  # dead for the scorer (models/global_eval_models.py:726 output unused)


# --------------------------------------------------------------------------------------
# CLIP ModifiedResNet-50 under timm ByobNet naming
# --------------------------------------------------------------------------------------


class ConvNormAct(nn.Module):
    """timm ConvNormAct: .conv, .bn (+ optional activation, + optional anti-alias pool .aa)."""

    def __init__(self, cin, cout, k, stride=1, apply_act=True, aa_stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.apply_act = apply_act
        self.aa = nn.AvgPool2d(aa_stride) if aa_stride > 1 else nn.Identity()

    def forward(self, x):
        x = self.bn(self.conv(x))
        if self.apply_act:
            x = F.relu(x)
        return self.aa(x)


class DownsampleAvg(nn.Module):
    """CLIP shortcut: AvgPool(stride) -> 1x1 conv -> BN."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.pool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv = ConvNormAct(cin, cout, 1, apply_act=False)

    def forward(self, x):
        return self.conv(self.pool(x))


class ClipBottleneck(nn.Module):
    def __init__(self, cin, planes, stride):
        super().__init__()
        cout = planes * 4
        self.shortcut = DownsampleAvg(cin, cout, stride) if (stride > 1 or cin != cout) else None
        self.conv1_1x1 = ConvNormAct(cin, planes, 1)
        # 3x3 always runs at stride 1; the stride is an average pool after its activation
        self.conv2_kxk = ConvNormAct(planes, planes, 3, aa_stride=stride)
        self.conv3_1x1 = ConvNormAct(planes, cout, 1, apply_act=False)
        self.act = nn.ReLU(inplace=True)

    def forward(self, x):
        shortcut = x if self.shortcut is None else self.shortcut(x)
        x = self.conv3_1x1(self.conv2_kxk(self.conv1_1x1(x)))
        return self.act(x + shortcut)


class ClipStem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = ConvNormAct(3, 32, 3, stride=2)
        self.conv2 = ConvNormAct(32, 32, 3)
        self.conv3 = ConvNormAct(32, 64, 3)
        self.pool = nn.AvgPool2d(2)

    def forward(self, x):
        return self.pool(self.conv3(self.conv2(self.conv1(x))))


class ClipAttentionPool(nn.Module):
    """CLIP AttentionPool2d (49+1 tokens, 32 heads, 2048 -> 1024).  Dead for the score."""

    def __init__(self, feat=7, dim=2048, heads=32, out=1024):
        super().__init__()
        self.pos_embed = nn.Parameter(torch.randn(feat * feat + 1, dim) / dim ** 0.5)
        self.q = nn.Linear(dim, dim)
        self.k = nn.Linear(dim, dim)
        self.v = nn.Linear(dim, dim)
        self.proj = nn.Linear(dim, out)
        self.heads = heads

    def forward(self, x):
        n, c, h, w = x.shape
        x = x.flatten(2).permute(0, 2, 1)
        x = torch.cat([x.mean(1, keepdim=True), x], 1)
        if x.shape[1] == self.pos_embed.shape[0]:
            x = x + self.pos_embed[None]
        q = self.q(x[:, :1]).view(n, 1, self.heads, -1).transpose(1, 2)
        k = self.k(x).view(n, -1, self.heads, c // self.heads).transpose(1, 2)
        v = self.v(x).view(n, -1, self.heads, c // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, c)
        return self.proj(o)


class TimmClipResNet50(nn.Module):
    default_cfg = {
        "input_size": (3, 224, 224),
        "interpolation": "bicubic",
        "mean": (0.48145466, 0.4578275, 0.40821073),
        "std": (0.26862954, 0.26130258, 0.27577711),
        "crop_pct": 1.0,
        "crop_mode": "center",
    }

    def __init__(self):
        super().__init__()
        self.stem = ClipStem()
        stages, cin = [], 64
        for s, (planes, blocks) in enumerate(zip((64, 128, 256, 512), (3, 4, 6, 3))):
            blks = []
            for b in range(blocks):
                blks.append(ClipBottleneck(cin, planes, stride=2 if (b == 0 and s > 0) else 1))
                cin = planes * 4
            stages.append(nn.Sequential(*blks))
        self.stages = nn.Sequential(*stages)
        self.head = ClipAttentionPool()

    def forward(self, x):
        return self.head(self.stages(self.stem(x)))


# --------------------------------------------------------------------------------------
# Seeded synthetic weights (there are no checkpoints offline)
# --------------------------------------------------------------------------------------

TRUNKS = {"resnet50": TimmResNet50, "resnet50_clip.openai": TimmClipResNet50}


def seeded_init_(model: nn.Module, seed: int = 0, calibrate_bn: bool = True) -> nn.Module:
    """Deterministic random init; BN gets non-trivial affine params and (optionally)
    running stats calibrated on one batch so that BN folding is a real test and
    activations stay O(1) through 50 layers."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
        elif isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
        elif isinstance(m, nn.Linear):
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) / m.in_features ** 0.5)
                m.bias.zero_()
    for name, p in model.named_parameters():
        if name.endswith("pos_embed"):
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) / p.shape[-1] ** 0.5)
    if calibrate_bn:
        bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
        for m in bns:
            m.reset_running_stats()
            m.momentum = None  # cumulative average -> running stats == batch stats of this pass
        model.train()
        with torch.no_grad():
            model(torch.randn(8, 3, 224, 224, generator=g))
        for m in bns:
            m.momentum = 0.1
        model.eval()
    return model


def build_trunk(name: str, seed: int = 0, calibrate_bn: bool = True) -> nn.Module:
    if name not in TRUNKS:
        raise ValueError(f"oracle trunk {name!r} not restated (have {sorted(TRUNKS)})")
    with torch.random.fork_rng(devices=[]):  # leave the caller's global RNG stream untouched
        model = TRUNKS[name]()
        seeded_init_(model, seed, calibrate_bn)
    model.pretrained_cfg = dict(model.default_cfg)
    return model.eval()
