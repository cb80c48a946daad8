#!/usr/bin/env python
"""Accuracy / speed probe of the split-precision mode: error of one long-K conv against fp64 (sign-aware bias =
truncation towards zero), scorer error against the fp32 and fp64 oracles on sweep-distribution and SR ~ GT pairs,
pairs/s.  Run under different SEMDIFF_X3_CHUNK_KB values to size the promoted accumulation.
usage: python tools/x3_accuracy.py [--precision fp16x3] [--pairs 48]"""
import argparse
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("SEMDIFF_RANDOM_INIT", "1")
import torch  # noqa: E402

import semdiff_b200  # noqa: E402
from helpers import DT, conv2d, split_store, split_value  # noqa: E402
from oracle.restated import RestatedScorer  # noqa: E402
from oracle.synth import make_pairs, set_head  # noqa: E402
from semdiff_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp16x3")
ap.add_argument("--pairs", type=int, default=48)
ap.add_argument("--trunk", default="resnet50")
args = ap.parse_args()
dt = DT[args.precision]
g = torch.Generator(device="cuda").manual_seed(5)
for cin, k in ((512, 3), (2048, 1), (256, 1)):
    x = split_store(torch.randn(4, 7, 7, cin, device="cuda", generator=g), dt)
    w = split_store(torch.randn(512, k, k, cin, device="cuda", generator=g) * 0.02, dt)
    b = torch.zeros(512, device="cuda")
    out = split_value(conv2d(x, w, b, None, 1, k // 2, False, args.precision, _lib.CONV_TC_TMA))
    ref = torch.nn.functional.conv2d(split_value(x).permute(0, 3, 1, 2), split_value(w).permute(0, 3, 1, 2), padding=k // 2).permute(0, 2, 3, 1)
    ref32 = torch.nn.functional.conv2d(split_value(x).float().permute(0, 3, 1, 2), split_value(w).float().permute(0, 3, 1, 2), padding=k // 2).permute(0, 2, 3, 1).double()
    sc = ref.abs().max()
    print(f"conv K={cin * k * k}: max err/max|ref| {((out - ref).abs().max() / sc).item():.3g} (torch fp32 {((ref32 - ref).abs().max() / sc).item():.3g}); "
          f"rms err/rms ref {((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item():.3g}; "
          f"mean((out-ref)*sign(ref))/mean|ref| {(((out - ref) * ref.sign()).mean() / ref.abs().mean()).item():.3g}")

cls = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb if args.trunk == "resnet50" else semdiff_b200.CLIP_lpips_stages_cnn
oracle = set_head(RestatedScorer(args.trunk, 3, seed=0), "abs")
o64 = set_head(RestatedScorer(args.trunk, 3, seed=0), "abs").double()
with contextlib.redirect_stdout(io.StringIO()):
    model = cls(args.trunk, 3, "cuda", precision=args.precision).eval()
model.load_state_dict(oracle.state_dict())
gt, sr = make_pairs(args.pairs, seed=41)
gt2, sr2 = make_pairs(16, seed=43, sigma_lo=0.02, sigma_hi=0.03)
gt, sr = torch.cat([gt, gt2]), torch.cat([sr, sr2])
with torch.no_grad():
    ref = torch.cat([oracle(gt[i:i + 16], sr[i:i + 16]) for i in range(0, gt.shape[0], 16)])
    r64 = torch.cat([o64(gt[i:i + 16].double(), sr[i:i + 16].double()) for i in range(0, gt.shape[0], 16)])
    got = model(gt.cuda(), sr.cuda()).cpu()
rel = lambda a, b: ((a.double() - b.double()).abs() / b.double().abs().clamp_min(1e-3))  # noqa: E731
print(f"scorer {args.trunk} {args.precision} chunk_kb={os.environ.get('SEMDIFF_X3_CHUNK_KB', 'default')}: "
      f"vs oracle fp32 max {rel(got, ref).max().item():.3g} median {rel(got, ref).median().item():.3g}; low-sigma corner max {rel(got[-16:], ref[-16:]).max().item():.3g}; "
      f"vs fp64: ours max {rel(got, r64).max().item():.3g} median {rel(got, r64).median().item():.3g}, oracle fp32 max {rel(ref, r64).max().item():.3g}; "
      f"signed mean rel err vs fp64 {((got.double() - r64) / r64).mean().item():.3g}")
n = 256
a = torch.randn(n, 3, 224, 224, device="cuda")
b = a + 0.1 * torch.randn_like(a)
with torch.no_grad():
    for _ in range(2):
        model(a, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model(a, b)
    e1.record()
    torch.cuda.synchronize()
print(f"throughput: {n * 5 / (e0.elapsed_time(e1) / 1e3):.0f} pairs/s")
