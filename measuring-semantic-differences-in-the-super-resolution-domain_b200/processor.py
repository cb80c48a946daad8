"""Eval-time image transform, the stand-in for `timm.data.create_transform(**data_config, is_training=False)`
(/root/reference/models/global_eval_models.py:333-334), used by the datasets as `model.processor(PIL)`
(/root/reference/datasets/global_eval_torch_ds.py:20-21).  It runs in DataLoader workers on the CPU, outside
forward(): resize (shorter side -> floor(size / crop_pct), bicubic) -> center crop -> [0,1] tensor -> normalise."""
from __future__ import annotations

import math


def make_processor(cfg: dict):
    try:
        import timm  # noqa: PLC0415
    except ImportError:     # no timm in this environment: the restated transform below (a timm that is installed but
        timm = None         # fails must fail loudly, not silently change the preprocessing)
    if timm is not None and getattr(timm, "__version__", None):
        return timm.data.create_transform(**timm.data.resolve_data_config(cfg), is_training=False)
    from torchvision import transforms as T  # noqa: PLC0415

    size = cfg["input_size"][-1]
    mode = {"bicubic": T.InterpolationMode.BICUBIC, "bilinear": T.InterpolationMode.BILINEAR}[cfg.get("interpolation", "bicubic")]
    return T.Compose([
        T.Resize(int(math.floor(size / cfg.get("crop_pct", 1.0))), interpolation=mode),
        T.CenterCrop(size),
        T.ToTensor(),
        T.Normalize(mean=cfg["mean"], std=cfg["std"]),
    ])
