"""ORACLE / test helper: seeded synthetic GT/SR pairs and head-weight modes.
Pure torch-CPU generators so any host regenerates the same tensors."""
from __future__ import annotations

import math

import torch


def make_pairs(n: int, seed: int = 0, size: int = 224, sigma_lo: float = 0.02, sigma_hi: float = 2.0):
    """gt ~ N(0,1); sr = (gt + sigma_i * noise) / sqrt(1 + sigma_i^2) with sigma_i log-uniform in
    [lo, hi], one per pair (variance-preserving: an SR output has the dynamic range of its GT, and
    sigma -> inf is the "independent image" case).  Per-pair generator seeded by (seed, i) so a
    shard can regenerate just its own pairs."""
    gts, srs = [], []
    for i in range(n):
        g = torch.Generator().manual_seed(1_000_003 * seed + i)
        gt = torch.randn(3, size, size, generator=g)
        u = torch.rand((), generator=g).item()
        sigma = math.exp(math.log(sigma_lo) + u * (math.log(sigma_hi) - math.log(sigma_lo)))
        srs.append((gt + sigma * torch.randn(3, size, size, generator=g)) / math.sqrt(1.0 + sigma * sigma))
        gts.append(gt)
    return torch.stack(gts), torch.stack(srs)


def set_head(model, mode: str):
    """'signed' = the reference's default Conv2d init (global_eval_models.py:336) as drawn;
    'abs' = |w|, |b| -- a trained LPIPS-style head is non-negative, and it avoids the catastrophic
    cancellation that makes relative error meaningless."""
    assert mode in ("signed", "abs")
    if mode == "abs":
        with torch.no_grad():
            for m in model.w_layers:
                m.weight.abs_()
                m.bias.abs_()
    return model
