#!/usr/bin/env python
"""ncu target: one warm-up forward, then ONE profiled forward between cudaProfilerStart/Stop
(run under `ncu --profile-from-start off ...`).  usage: python tools/ncu_target.py [--pairs 256] [--trunk resnet50]"""
import argparse
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SEMDIFF_RANDOM_INIT", "1")
import torch  # noqa: E402

import semdiff_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=256)
ap.add_argument("--microbatch", type=int, default=0)
ap.add_argument("--trunk", default="resnet50")
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
cls = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb if args.trunk == "resnet50" else semdiff_b200.CLIP_lpips_stages_cnn
with contextlib.redirect_stdout(io.StringIO()):
    model = cls(clip_name=args.trunk, depth=3, device="cuda", precision=args.precision, microbatch=args.microbatch or None).eval()
gt = torch.randn(args.pairs, 3, 224, 224, device="cuda")
sr = gt + 0.1 * torch.randn_like(gt)
with torch.no_grad():
    model(gt, sr)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    s = model(gt, sr)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", float(s.sum()))
