"""ORACLE (test infrastructure): from-scratch restatement of the reference scorer so that the
check can run where /root/reference does not exist (the GPU box).  Validated against the
unmodified reference file in tests/test_oracle.py (bit-identical on CPU fp32).

Follows /root/reference/models/global_eval_models.py:
  ctor      :309-339 / :683-715   (tap list, w_layers = Conv2d(256*2**s, 1, 1))
  forward   :341-397 / :717-773   ((a-b)**2 -> w_layers[j] -> mean W -> mean H -> mean layers -> ReLU)
  save/load :419-429 / :795-805
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .trunks import build_trunk


def tap_names(trunk: str, depth: int, variant: str = "stages"):
    if variant == "wperlay":                                            # :832-833
        return [f"stages.{s}.{lay}.act" for s in range(4) for lay in range(3)][11 - depth:]
    if trunk == "resnet50":
        return [f"layer{s}.2.act3" for s in range(4 - depth, 5)]      # :701
    return [f"stages.{s}.2.act" for s in range(3 - depth, 4)]          # :327


class RestatedScorer(nn.Module):
    def __init__(self, trunk: str, depth: int, seed: int = 0, calibrate_bn: bool = True, variant: str = "stages"):
        super().__init__()
        self.clip = build_trunk(trunk, seed=seed, calibrate_bn=calibrate_bn)
        self.depth = depth
        self.trunk_name = trunk
        self.wanted_layers = tap_names(trunk, depth, variant)
        torch.manual_seed(seed + 1000)  # same stream as reference_loader.build_reference_scorer
        if variant == "wperlay":   # :841-847: one Conv2d(256 * 2**stage, 1, 1) per hooked block
            chans = [256 * (2 ** int(name.split(".")[1])) for name in self.wanted_layers]
        else:                      # :336
            chans = [256 * (2 ** s) for s in range(3 - depth, 4)]
        self.w_layers = nn.ModuleList([nn.Conv2d(c, 1, kernel_size=1, stride=1) for c in chans])
        self._taps = {}
        mods = dict(self.clip.named_modules())
        for name in self.wanted_layers:
            mods[name].register_forward_hook(self._hook(name))
        self.eval()

    def _hook(self, name):
        def hook(module, inp, out):
            self._taps[name] = out
        return hook

    def features(self, x):
        self._taps = {}
        self.clip(x)
        return [self._taps[n] for n in self.wanted_layers]

    @torch.no_grad()
    def forward(self, a, b, pre_relu: bool = False):
        fa, fb = self.features(a), self.features(b)
        per_layer = []
        for j, (xa, xb) in enumerate(zip(fa, fb)):
            d = (xa - xb) ** 2                                            # :379
            w = self.w_layers[j](d).squeeze(1)                            # :381
            per_layer.append(torch.mean(torch.mean(w, dim=-1), dim=-1))   # :384
        s = per_layer[0] if len(per_layer) == 1 else torch.mean(torch.stack(per_layer), dim=0)  # :385-392
        return s if pre_relu else torch.relu(s)                           # :395


# ---------------------------------------------------------------------------------------------
# local maps: /root/reference/models/local_eval_models.py:7-171 (CLIP trunk) and :175-339 (ImageNet trunk)
# ---------------------------------------------------------------------------------------------
def unet_tap_names(trunk: str):
    if trunk == "resnet50":
        return ["conv1"] + [f"layer{s}.2.act3" for s in range(1, 5)]        # :196
    return ["stem.conv3"] + [f"stages.{s}.2.act" for s in range(4)]          # :27


def unet_decoder() -> nn.ModuleList:
    """:38-82 / :207-251"""
    def level(cin, cout):
        return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding="same"), nn.BatchNorm2d(cout), nn.ReLU(),
                             nn.Conv2d(cout, cout, kernel_size=3, padding="same"), nn.BatchNorm2d(cout), nn.ReLU())
    head = nn.Sequential(nn.Conv2d(256 + 64, 64, kernel_size=3, padding="same"), nn.BatchNorm2d(64), nn.ReLU(),
                         nn.Conv2d(64, 1, kernel_size=1, padding="same"), nn.ReLU())
    return nn.ModuleList([head, level(256 + 512, 256), level(512 + 1024, 512), level(1024 + 2048, 1024), level(2048, 2048)])


class RestatedUnet(nn.Module):
    def __init__(self, trunk: str, seed: int = 0, calibrate_bn: bool = True):
        super().__init__()
        self.clip = build_trunk(trunk, seed=seed, calibrate_bn=calibrate_bn)
        self.trunk_name = trunk
        self.wanted_layers = unet_tap_names(trunk)
        torch.manual_seed(seed + 2000)   # same stream as reference_loader.build_reference_unet
        self.decoder = unet_decoder()
        self.upscaler = nn.UpsamplingBilinear2d(scale_factor=2)             # :84
        self.final_sigmoid = nn.Sigmoid()
        for m in self.decoder.modules():                                     # init_weights, :144-157
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._taps = {}
        mods = dict(self.clip.named_modules())
        for name in self.wanted_layers:
            mods[name].register_forward_hook(self._hook(name))
        self.eval()

    def _hook(self, name):
        def hook(module, inp, out):
            self._taps[name] = out
        return hook

    def features(self, x):
        self._taps = {}
        self.clip(x)
        return [self._taps[n] for n in self.wanted_layers]

    def decode(self, diff, logits: bool = False):
        x = self.upscaler(self.decoder[-1](diff[-1]))                        # :117-118
        for j in range(2, len(diff) + 1):                                    # :119-123
            x = self.upscaler(self.decoder[-j](torch.concat((diff[-j], x), dim=1)))
        return x if logits else self.final_sigmoid(x)                       # :125

    @torch.no_grad()
    def forward(self, a, b, logits: bool = False):
        fa, fb = self.features(a), self.features(b)
        return self.decode([(xa - xb) ** 2 for xa, xb in zip(fa, fb)], logits)   # :115


def calibrate_unet_decoder(model, seed: int = 0):
    """Give the decoder non-trivial BatchNorm statistics, affine parameters and conv biases (its default init is identity
    BatchNorm + zero biases, under which the map saturates): one train-mode pass over seeded pairs with momentum=None sets
    the running statistics, then seeded affine / bias values.  Works on the reference's module and on RestatedUnet alike
    (both expose .clip hooks via forward and .decoder)."""
    from .synth import make_pairs

    g = torch.Generator().manual_seed(seed + 3000)
    with torch.no_grad():
        for m in model.decoder.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, nn.Conv2d):
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    gt, sr = make_pairs(4, seed=seed + 77)
    bns = [m for m in model.decoder.modules() if isinstance(m, nn.BatchNorm2d)]
    for m in bns:
        m.momentum = None
        m.reset_running_stats()
    model.decoder.train()
    with torch.no_grad():
        model(gt, sr)
    model.decoder.eval()
    for m in bns:
        m.momentum = 0.1
    with torch.no_grad():   # keep the 1-channel head in the sigmoid's sensitive range
        head = list(model.decoder[0].children())[3]
        head.weight.mul_(0.1)
        head.bias.fill_(0.3)
    return model
