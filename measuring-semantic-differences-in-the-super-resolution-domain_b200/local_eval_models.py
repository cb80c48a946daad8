"""B200 inference path of the reference's LOCAL semantic-difference maps.

    reference                                                     here
    models/local_eval_models.py:7    CLIP_lpips_Unet               CLIP_lpips_Unet            (timm resnet50_clip.openai)
    models/local_eval_models.py:175  CLIP_lpips_Unet_clsbckbn      CLIP_lpips_Unet_clsbckbn   (timm resnet50)

Same constructor `(clip_name, device, lora_rank=None)`, same attributes (`clip`, `decoder`, `upscaler`, `final_sigmoid`,
`wanted_layers`, `processor`), same `state_dict()` keys and `save_model` / `load_model` files (decoder.state_dict() when
lora_rank is None, :160-171), same `forward(a, b) -> [N, 1, H, W]` sigmoid map.  Underneath, forward() is ONE call into
libsemdiff_b200.so (semdiff_score_map): both trunk passes, the squared differences of the five taps (:115), the U-Net
decoder (:117-123; 3x3 convs on the tcgen05 kernels with BatchNorm and bias folded, channel concat, bilinear x2
upsampling with align_corners) and the final sigmoid.

Scope: INFERENCE.  The decoder's BatchNorm runs with its running statistics (module.eval() semantics) and there is no
backward pass: calling forward() with autograd enabled on trainable decoder parameters raises - train the decoder with
the reference module, load the checkpoint here with load_model().  LoRA trunks (`lora_rank` not None) are not supported.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, trunks
from .global_eval_models import _Plan
from .preprocess import GpuProcessor
from .processor import make_processor


def _decoder() -> nn.ModuleList:
    """The reference's decoder, module for module (:38-82 / :207-251), so that state_dict keys and checkpoints match."""
    def level(cin, cout):
        return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding="same"), nn.BatchNorm2d(cout), nn.ReLU(),
                             nn.Conv2d(cout, cout, kernel_size=3, padding="same"), nn.BatchNorm2d(cout), nn.ReLU())
    head = nn.Sequential(nn.Conv2d(256 + 64, 64, kernel_size=3, padding="same"), nn.BatchNorm2d(64), nn.ReLU(),
                         nn.Conv2d(64, 1, kernel_size=1, padding="same"), nn.ReLU())
    return nn.ModuleList([head, level(256 + 512, 256), level(512 + 1024, 512), level(1024 + 2048, 1024), level(2048, 2048)])


class _B200Unet(nn.Module):
    _FAMILY = None

    def __init__(self, clip_name: str, device: str, lora_rank: int | None = None, *, precision: str = "fp16x3",
                 microbatch: int | None = None, pretrained: bool | None = None):
        super().__init__()
        if lora_rank is not None:
            raise NotImplementedError("lora_rank: LoRA / full fine-tuning of the trunk is not supported by the B200 path "
                                      "(inference-only kernel program with folded BatchNorm)")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"device={device!r}: the B200 path has no CPU fallback; use the reference module on CPU")
        _lib.load()
        self.family = trunks.trunk_family(clip_name)
        if self.family != self._FAMILY:
            raise ValueError(f"{type(self).__name__} hooks the modules of a {self._FAMILY} trunk (got {clip_name!r})")
        self.clip = trunks.create_trunk(clip_name, pretrained=pretrained)
        self.lora_rank = lora_rank
        self.clip.eval()
        self.clip.to(dev)
        self.wanted_layers = self._tap_names()
        cfg = getattr(self.clip, "pretrained_cfg", None) or {}
        self.processor = make_processor(dict(cfg))
        self.gpu_processor = GpuProcessor(dict(cfg), dev)
        self.decoder = _decoder()
        self.upscaler = nn.UpsamplingBilinear2d(scale_factor=2)
        self.final_sigmoid = nn.Sigmoid()
        self.init_weights()
        self.decoder.to(dev)
        self.precision, self.microbatch = precision, microbatch
        self._device = dev
        self._plan = {}

    def init_weights(self):
        """:144-157"""
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def save_model(self, path: str):
        torch.save(self.decoder.state_dict(), path)       # lora_rank is None here (:165)

    def load_model(self, path: str):
        self.decoder.load_state_dict(torch.load(path, weights_only=True))   # (:171)
        self._plan = {}

    # ---- keep the native plan in sync with the parameters ----
    def train(self, mode: bool = True):
        super().train(mode)
        self.clip.eval()
        return self

    def _apply(self, fn, *args, **kwargs):
        self._plan = {}
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._plan = {}
        return super().load_state_dict(*args, **kwargs)

    def refresh(self):
        """Re-fold trunk and decoder after their parameters / BatchNorm statistics were modified in place."""
        self._plan = {}

    def plan(self, stem=True) -> _Plan:
        if stem not in self._plan:
            program = trunks.lower_unet(self.clip, self.decoder, self.family, s2d_stem=stem)
            self._plan[stem] = _Plan(self.clip, self.family, 3, self.precision, self._device, program=program)
        return self._plan[stem]

    def default_microbatch(self, H: int, W: int) -> int:
        """Pairs per pass: the decoder's widest tensors (320 channels at H/2 x W/2) bound the workspace at a few GB."""
        if self.microbatch:
            return int(self.microbatch)
        return max(1, min(64, (64 * 224 * 224) // max(H * W, 1)))

    def forward(self, a, b):
        if a.shape != b.shape or a.dim() != 4 or a.shape[1] != 3:
            raise ValueError(f"expected two [N,3,H,W] tensors of the same shape, got {tuple(a.shape)} and {tuple(b.shape)}")
        if a.device.type != "cuda" or b.device.type != "cuda":
            raise RuntimeError("inputs must be CUDA tensors (no CPU fallback)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.decoder.parameters()):
            raise NotImplementedError("the B200 local-map path is inference-only: call it under torch.no_grad() (train the decoder "
                                      "with the reference module and load the checkpoint with load_model())")
        n, _, H, W = a.shape
        if H % 32 or W % 32:
            raise ValueError(f"the U-Net needs image sizes that are multiples of 32 (five x2 levels), got {H}x{W}")
        plan = self.plan(True)
        in_dt = a.dtype if a.dtype in (torch.bfloat16, torch.float16) and b.dtype == a.dtype else torch.float32
        a, b = a.detach().contiguous().to(in_dt), b.detach().contiguous().to(in_dt)
        in_prec = {torch.float32: _lib.FP32, torch.bfloat16: _lib.BF16, torch.float16: _lib.FP16}[in_dt]
        out = torch.empty(n, 1, H, W, dtype=torch.float32, device=a.device)
        if n == 0:
            return out
        mb = min(self.default_microbatch(H, W), n)
        ws = plan.workspace(mb, H, W)
        with torch.cuda.device(a.device):
            rc = plan.lib.semdiff_score_map(plan.handle, a.data_ptr(), b.data_ptr(), in_prec, n, H, W, mb, ws.data_ptr(), ws.numel(),
                                            out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "semdiff_score_map")
        return out


class CLIP_lpips_Unet(_B200Unet):
    """timm `resnet50_clip.openai` trunk; taps stem.conv3 + stages.{s}.2.act (reference :7-171, hook list :27)."""
    _FAMILY = "resnet50_clip.openai"

    def _tap_names(self):
        return ["stem.conv3"] + [f"stages.{s}.{2}.act" for s in range(4)]


class CLIP_lpips_Unet_clsbckbn(_B200Unet):
    """timm `resnet50` (ImageNet) trunk; taps conv1 (the raw conv, before bn1) + layer{s}.2.act3 (reference :175-339, :196)."""
    _FAMILY = "resnet50"

    def _tap_names(self):
        return ["conv1"] + [f"layer{s}.2.act3" for s in range(1, 5)]
