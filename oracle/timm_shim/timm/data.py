"""ORACLE shim for timm.data (see timm/__init__.py).  Restates timm's eval transform:
resize(shorter side = floor(size / crop_pct), interpolation) -> center crop -> to tensor ->
normalise.  crop_pct / interpolation per tag are from memory of timm's pretrained cfgs
(parity unpinned)."""
import math


def resolve_model_data_config(model, **kwargs):
    cfg = dict(getattr(model, "pretrained_cfg", None) or model.default_cfg)
    return {k: cfg[k] for k in ("input_size", "interpolation", "mean", "std", "crop_pct", "crop_mode")}


def create_transform(input_size, interpolation="bicubic", mean=None, std=None, crop_pct=1.0,
                     crop_mode="center", is_training=False, **kwargs):
    from torchvision import transforms as T

    assert not is_training
    size = input_size[-1]
    interp = {"bicubic": T.InterpolationMode.BICUBIC, "bilinear": T.InterpolationMode.BILINEAR}[interpolation]
    return T.Compose([
        T.Resize(int(math.floor(size / crop_pct)), interpolation=interp),
        T.CenterCrop(size),
        T.ToTensor(),
        T.Normalize(mean=mean, std=std),
    ])
