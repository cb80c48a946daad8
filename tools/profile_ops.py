#!/usr/bin/env python
"""Per-op timing table of the trunk program (CUDA events around every op, semdiff_plan_set_profiling):
shape, ms, TFLOP/s and activation GB/s per conv.  usage: python tools/profile_ops.py [--pairs 256] [--microbatch 64] ..."""
import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SEMDIFF_RANDOM_INIT", "1")
import torch  # noqa: E402

import semdiff_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--microbatch", type=int, nargs="+", default=[64])
    ap.add_argument("--trunk", default="resnet50")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--impl", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    cls = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb if args.trunk == "resnet50" else semdiff_b200.CLIP_lpips_stages_cnn
    with contextlib.redirect_stdout(io.StringIO()):
        model = cls(clip_name=args.trunk, depth=3, device="cuda", precision=args.precision).eval()
    n, S = args.pairs, args.size
    gt = torch.randn(n, 3, S, S, device="cuda")
    sr = gt + 0.1 * torch.randn_like(gt)
    results = {}
    for mb in args.microbatch:
        model.microbatch = mb
        plan = model.plan()
        if args.impl:
            plan.set_conv_impl(args.impl)
        with torch.no_grad():
            for _ in range(2):
                model(gt, sr)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                model(gt, sr)
            e1.record()
            torch.cuda.synchronize()
            step_ms = e0.elapsed_time(e1) / args.steps
            plan.set_profiling(True)
            model(gt, sr)
            plan.profile(reset=True)
            for _ in range(args.steps):
                model(gt, sr)
            ms, cnt = plan.profile(reset=True)
            plan.set_profiling(False)
        ops = plan.program.ops
        shapes = {0: {1: (S // 2 + 3, S // 2, 64), 2: (S // 2 + 1, S // 2, 64), 3: (S // 2, S // 2, 16)}.get(plan.program.input_layout, (S, S, 8))}
        rows, tot_conv, tot_flop = [], 0.0, 0.0
        eb = 4 if args.precision in ("fp32", "fp16x3", "bf16x3") else 2
        for i, op in enumerate(ops):
            h, w, c = shapes[op["src"]]
            t = ms[i] / args.steps
            if op["kind"] == 0:
                ph = op["pad"] + (op["pad"] if op["pad_hi"] < 0 else op["pad_hi"])
                oh = (h + ph - op["kh"]) // op["stride"] + 1
                ow = (w + ph - op["kw"]) // op["stride"] + 1
                shapes[op["dst"]] = (oh, ow, op["cout"])
                fl = 2.0 * oh * ow * op["alg_cout"] * op["alg_k"] * 2 * n
                by = (h * w * c + oh * ow * op["cout"] * (2 if op["res"] >= 0 else 1)) * eb * 2 * n
                rows.append((i, f"conv {op['kh']}x{op['kw']}s{op['stride']} {c}->{op['cout']} @{h}", t, fl / t / 1e9 if t else 0, by / t / 1e6 if t else 0))
                tot_conv += t
                tot_flop += fl
            elif op["kind"] == 1:
                shapes[op["dst"]] = ((h - 1) // 2 + 1, (w - 1) // 2 + 1, c)
                rows.append((i, f"maxpool @{h}", t, 0, (h * w * c * 1.25) * eb * 2 * n / t / 1e6 if t else 0))
            elif op["kind"] == 2:
                shapes[op["dst"]] = (h // op["stride"], w // op["stride"], c)
                rows.append((i, f"avgpool{op['stride']} @{h}", t, 0, (h * w * c * 1.25) * eb * 2 * n / t / 1e6 if t else 0))
        k = len(ops)
        print(f"\n=== microbatch {mb} pairs: step {step_ms:.2f} ms = {n / step_ms * 1e3:.0f} pairs/s; conv total {tot_conv:.2f} ms "
              f"({tot_flop / tot_conv / 1e9:.0f} TFLOP/s), pack {ms[k] / args.steps:.3f} distance {ms[k + 1] / args.steps:.3f} head {ms[k + 2] / args.steps:.3f}")
        for r in rows:
            print(f"{r[0]:3d} {r[1]:34s} {r[2]:8.3f} ms {r[3]:8.1f} TF/s {r[4]:8.0f} GB/s")
        results[mb] = {"step_ms": step_ms, "conv_ms": tot_conv, "rows": rows}
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f)


if __name__ == "__main__":
    main()
