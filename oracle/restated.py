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
