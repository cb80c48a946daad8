"""B200-native global semantic-fidelity scorer (CLIP-LPIPS regressor): drop-in for
/root/reference/models/global_eval_models.py's CLIP_lpips_stages_cnn / CLIP_lpips_stages_cnn_clsbckb / CLIP_lpips_wperlay_cnn,
plus the inference path of the local-map U-Nets of /root/reference/models/local_eval_models.py (CLIP_lpips_Unet,
CLIP_lpips_Unet_clsbckbn).  Import as `semdiff_b200` (alias module at the repo root; this directory's name is not a
Python identifier)."""
from . import _lib, trunks  # noqa: F401
from .global_eval_models import CLIP_lpips_stages_cnn, CLIP_lpips_stages_cnn_clsbckb, CLIP_lpips_wperlay_cnn  # noqa: F401
from .local_eval_models import CLIP_lpips_Unet, CLIP_lpips_Unet_clsbckbn  # noqa: F401

__all__ = ["CLIP_lpips_stages_cnn", "CLIP_lpips_stages_cnn_clsbckb", "CLIP_lpips_wperlay_cnn", "CLIP_lpips_Unet",
           "CLIP_lpips_Unet_clsbckbn", "trunks"]
