"""ORACLE (test infrastructure): import the reference scorer file UNMODIFIED.

/root/reference/models/global_eval_models.py does `import timm` (:3) and
`from icecream import ic` (:919); neither is installed, so oracle/timm_shim is put on
sys.path for the duration of the import.  Everything the reference itself wrote (hooks,
squared diff, w_layers, means, ReLU, save/load) then runs verbatim; only the third-party
trunk is restated (oracle/trunks.py).  /root/reference exists in the build container only;
on the GPU box this loader reports unavailable and tests fall back to oracle/restated.py +
tests/golden/.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys

REFERENCE_FILE = "/root/reference/models/global_eval_models.py"
REFERENCE_LOCAL_FILE = "/root/reference/models/local_eval_models.py"
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "timm_shim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cached = None
_cached_local = None


def available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


def load_reference_module():
    """Returns the reference's `global_eval_models` module object."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(REFERENCE_FILE)
    for p in (_REPO, _SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    spec = importlib.util.spec_from_file_location("_reference_global_eval_models", REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cached = mod
    return mod


def build_reference_scorer(trunk: str, depth: int, seed: int = 0, calibrate_bn: bool = True, quiet: bool = True,
                           variant: str = "stages"):
    """Instantiate the reference class for `trunk` on CPU with the seeded oracle trunk.
    resnet50 -> CLIP_lpips_stages_cnn_clsbckb (:682), resnet50_clip.openai -> CLIP_lpips_stages_cnn (:308)."""
    import torch

    mod = load_reference_module()
    import timm  # the shim (on sys.path after load_reference_module)
    timm.SEED, timm.CALIBRATE_BN = seed, calibrate_bn
    cls = mod.CLIP_lpips_stages_cnn_clsbckb if trunk == "resnet50" else mod.CLIP_lpips_stages_cnn
    if variant == "wperlay":
        cls = mod.CLIP_lpips_wperlay_cnn  # :815
    torch.manual_seed(seed + 1000)  # seeds the reference's own default init of w_layers (:336)
    with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
        model = cls(clip_name=trunk, depth=depth, device="cpu")
    return model.eval()


def load_reference_local_module():
    """The reference's `local_eval_models` module object (imports timm and pytora: both shimmed)."""
    global _cached_local
    if _cached_local is not None:
        return _cached_local
    if not os.path.isfile(REFERENCE_LOCAL_FILE):
        raise FileNotFoundError(REFERENCE_LOCAL_FILE)
    for p in (_REPO, _SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    # The file does not compile as a whole (stray token + broken indentation at :624-625, inside CLIP_lpips_Unet_v3), so the
    # loader executes its text verbatim up to the third class: lines 1-339 = imports, CLIP_lpips_Unet, CLIP_lpips_Unet_clsbckbn.
    import types

    with open(REFERENCE_LOCAL_FILE) as f:
        src = f.read()
    cut = src.index("class CLIP_lpips_Unet_v2(")
    mod = types.ModuleType("_reference_local_eval_models")
    mod.__file__ = REFERENCE_LOCAL_FILE
    exec(compile(src[:cut], REFERENCE_LOCAL_FILE, "exec"), mod.__dict__)
    _cached_local = mod
    return mod


def build_reference_unet(trunk: str, seed: int = 0, calibrate_bn: bool = True):
    """resnet50 -> CLIP_lpips_Unet_clsbckbn (local_eval_models.py:175), resnet50_clip.openai -> CLIP_lpips_Unet (:7),
    on CPU with the seeded oracle trunk and the reference's own decoder init (:144-157), lora_rank=None."""
    import torch

    mod = load_reference_local_module()
    import timm  # the shim
    timm.SEED, timm.CALIBRATE_BN = seed, calibrate_bn
    cls = mod.CLIP_lpips_Unet_clsbckbn if trunk == "resnet50" else mod.CLIP_lpips_Unet
    torch.manual_seed(seed + 2000)   # seeds the reference's kaiming init of the decoder
    return cls(clip_name=trunk, device="cpu").eval()
