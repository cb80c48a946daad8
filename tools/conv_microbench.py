#!/usr/bin/env python
"""Time single conv shapes through the C-ABI with the TMA and the cp.async-gather A producers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import lib, sp
from semdiff_b200 import _lib

def run(n, H, W, cin, cout, k, stride, pad, impl, reps=10):
    x = torch.randn(n, H, W, cin, device="cuda").bfloat16()
    w = torch.randn(cout, k * k * cin, device="cuda").bfloat16() * 0.05
    b = torch.zeros(cout, device="cuda")
    oh, ow = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty(n, oh, ow, cout, device="cuda", dtype=torch.bfloat16)
    args = (x.data_ptr(), w.data_ptr(), b.data_ptr(), None, out.data_ptr(), n, H, W, cin, cout, k, k, stride, pad, 1, None, 0, 0, 0, 1, -1, _lib.BF16, impl)
    for _ in range(3): _lib.check(lib().semdiff_conv2d(*args, sp()), "conv")
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(reps): lib().semdiff_conv2d(*args, sp())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * n * oh * ow * cout * k * k * cin / ms / 1e9

for shape in [(512, 56, 56, 64, 64, 3, 1, 1), (512, 28, 28, 128, 128, 3, 1, 1), (512, 14, 14, 256, 256, 3, 1, 1), (512, 56, 56, 64, 64, 1, 1, 0), (512, 112, 112, 64, 64, 3, 1, 1)]:
    for name, impl in (("tma", _lib.CONV_TC_TMA), ("gather", _lib.CONV_TC_GATHER)):
        ms, tf = run(*shape, impl)
        print(f"{shape} {name:7s} {ms:7.3f} ms {tf:7.1f} TF/s", flush=True)
