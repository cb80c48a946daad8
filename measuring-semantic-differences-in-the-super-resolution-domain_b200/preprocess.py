"""On-device version of `model.processor` for BATCHES of decoded uint8 images.

The reference preprocesses one PIL image at a time on the CPU (`model.processor(Image.open(...))`,
/root/reference/datasets/global_eval_torch_ds.py:20-21; transform built at /root/reference/models/global_eval_models.py:333-334).
`GpuProcessor` applies the same transform (Pillow-exact 8-bit bicubic resize of the shorter side, center crop, /255,
normalise) to a uint8 [N, H, W, 3] CUDA tensor with the kernels in csrc/preprocess.cu, bit-exactly."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


class GpuProcessor:
    def __init__(self, cfg: dict, device):
        self.size = int(cfg["input_size"][-1])
        self.resize_to = int(math.floor(self.size / cfg.get("crop_pct", 1.0)))
        if cfg.get("interpolation", "bicubic") != "bicubic":
            raise ValueError("GpuProcessor implements the bicubic eval transform only")
        self.mean = (C.c_float * 3)(*cfg["mean"])
        self.std = (C.c_float * 3)(*cfg["std"])
        self.device = torch.device(device)
        self._tables = {}

    def _axis_table(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        if key not in self._tables:
            lib = _lib.load()
            ksize = _lib.check(lib.semdiff_resize_ksize(in_size, out_size), "semdiff_resize_ksize")
            bounds = torch.empty(out_size, 2, dtype=torch.int32)
            coeffs = torch.empty(out_size, ksize, dtype=torch.int32)
            _lib.check(lib.semdiff_resize_coeffs(in_size, out_size, bounds.data_ptr(), coeffs.data_ptr()), "semdiff_resize_coeffs")
            self._tables[key] = (bounds.to(self.device), coeffs.to(self.device), ksize)
        return self._tables[key]

    def geometry(self, H: int, W: int):
        """(Hr, Wr, top, left): torchvision Resize(int) + CenterCrop arithmetic."""
        if H <= W:
            Hr, Wr = self.resize_to, int(self.resize_to * W / H)
        else:
            Hr, Wr = int(self.resize_to * H / W), self.resize_to
        return Hr, Wr, int(round((Hr - self.size) / 2.0)), int(round((Wr - self.size) / 2.0))

    @torch.no_grad()
    def __call__(self, images_u8: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3 or images_u8.device.type != "cuda":
            raise ValueError("expected a CUDA uint8 tensor [N, H, W, 3] (decoded RGB images of one size)")
        n, H, W, _ = images_u8.shape
        Hr, Wr, top, left = self.geometry(H, W)
        bx, cx, kx = self._axis_table(W, Wr)
        by, cy, ky = self._axis_table(H, Hr)
        src = images_u8.contiguous()
        tmp = torch.empty(n, H, self.size, 3, dtype=torch.uint8, device=src.device)
        out = torch.empty(n, 3, self.size, self.size, dtype=dtype, device=src.device)
        prec = {torch.float32: _lib.FP32, torch.bfloat16: _lib.BF16, torch.float16: _lib.FP16}[dtype]
        lib = _lib.load()
        with torch.cuda.device(src.device):
            rc = lib.semdiff_preprocess_u8(src.data_ptr(), n, H, W, Hr, Wr, top, left, self.size, self.size, bx.data_ptr(),
                                           cx.data_ptr(), kx, by.data_ptr(), cy.data_ptr(), ky, self.mean, self.std,
                                           tmp.data_ptr(), out.data_ptr(), prec, _lib.stream_ptr())
        _lib.check(rc, "semdiff_preprocess_u8")
        return out
