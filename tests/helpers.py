"""Test helpers: call the C-ABI (include/semdiff_b200.h) on torch tensors."""
from __future__ import annotations

import ctypes as C

import torch

from semdiff_b200 import _lib

DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32, "fp16x3": torch.float16, "bf16x3": torch.bfloat16}


def split_store(x: torch.Tensor, dtype) -> torch.Tensor:
    """[..., C] (C % 64 == 0) real values -> the split-precision storage [..., 2C] in `dtype` (include/semdiff_b200.h):
    per block of 64 channels, 64 hi values (x rounded) then 64 lo values (x - hi, rounded)."""
    x = x.double()
    hi = x.to(dtype)
    lo = (x - hi.double()).to(dtype)
    c = x.shape[-1]
    assert c % 64 == 0
    out = torch.stack([hi.reshape(*x.shape[:-1], c // 64, 64), lo.reshape(*x.shape[:-1], c // 64, 64)], dim=-2)
    return out.reshape(*x.shape[:-1], 2 * c).contiguous()


def split_value(s: torch.Tensor) -> torch.Tensor:
    """Inverse of split_store: [..., 2C] stored -> [..., C] fp64 values hi + lo."""
    c2 = s.shape[-1]
    v = s.double().reshape(*s.shape[:-1], c2 // 128, 2, 64)
    return (v[..., 0, :] + v[..., 1, :]).reshape(*s.shape[:-1], c2 // 2)


def lib():
    return _lib.load()


def sp():
    return _lib.stream_ptr()


def nhwc(x: torch.Tensor, dtype) -> torch.Tensor:
    """NCHW fp32 -> contiguous NHWC in `dtype`."""
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(x: torch.Tensor) -> torch.Tensor:
    return x.float().permute(0, 3, 1, 2).contiguous()


def conv2d(x_nhwc, w_ohwi, bias, residual, stride, pad, relu, precision: str, impl: int, x2_nhwc=None, w2=None,
           stride2=1, pad_hi=-1):
    """w_ohwi [Cout,KH,KW,Cin]; optional fused second source x2 [n,H2,W2,Cin2] with 1x1 weights w2 [Cout,Cin2]."""
    split = precision in _lib.SPLIT   # x, residual, x2 and the result are in split storage (2C stored channels); w_ohwi / w2
    n, H, W, cin = x_nhwc.shape       # are [Cout,KH,KW,2Cin] / [Cout,2Cin2] with the same interleave along Cin
    cout, kh, kw, _ = w_ohwi.shape
    if split:
        cin //= 2
    pa = pad if pad_hi < 0 else pad_hi
    oh, ow = (H + pad + pa - kh) // stride + 1, (W + pad + pa - kw) // stride + 1
    # compute-sanitizer is closed on this GPU pool, so every conv test carries its own canary: the output sits between
    # two guard regions that must come back untouched (catches out-of-range TMA stores / epilogue writes)
    guard = 8192
    numel = n * oh * ow * cout * (2 if split else 1)
    arena = torch.full((numel + 2 * guard,), 123.0, dtype=x_nhwc.dtype, device=x_nhwc.device)
    out = arena[guard:guard + numel].view(n, oh, ow, cout * (2 if split else 1))
    wmat = w_ohwi.reshape(cout, -1)
    H2 = W2 = cin2 = 0
    if x2_nhwc is not None:
        _, H2, W2, cin2 = x2_nhwc.shape
        if split:
            cin2 //= 2
        wmat = torch.cat([wmat, w2], dim=1)
    wmat = wmat.contiguous()
    rc = lib().semdiff_conv2d(x_nhwc.data_ptr(), wmat.data_ptr(), bias.data_ptr(),
                              residual.data_ptr() if residual is not None else None, out.data_ptr(), n, H, W, cin, cout,
                              kh, kw, stride, pad, int(relu), x2_nhwc.data_ptr() if x2_nhwc is not None else None,
                              H2, W2, cin2, stride2, pad_hi, _lib.PRECISIONS[precision], impl, sp())
    _lib.check(rc, "semdiff_conv2d")
    torch.cuda.synchronize()
    assert bool((arena[:guard] == 123.0).all()) and bool((arena[guard + numel:] == 123.0).all()), "kernel wrote outside its output"
    return out


def conv_reference(x_nhwc, w_ohwi, bias, residual, stride, pad, relu):
    """fp32 torch reference on the SAME (already rounded) operands."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = torch.nn.functional.conv2d(nchw(x_nhwc).double(), w_ohwi.double().permute(0, 3, 1, 2), bias.double(),
                                   stride=stride, padding=pad)
    if residual is not None:
        y = y + nchw(residual).double()
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()
