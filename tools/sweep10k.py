#!/usr/bin/env python
"""BASELINE.json configs[3]: synthetic scoring sweep (SR-outputs-dataset shape: 512x512 sources resized to 224),
sharded over the ranks of the job with one all-gather of scores; rank agreement between precisions.

    python tools/sweep10k.py [--pairs 10000] [--modes bf16 fp16] [--oracle 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep10k.py ...

Pairs are generated on the device from per-pair seeds (any shard can regenerate its own pairs).  Reference ranking =
the fp32 mode of the same module (validated against the oracle to ~2e-6); --oracle K also scores the first K pairs with
the CPU oracle (the reference's own arithmetic)."""
import argparse
import contextlib
import io
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SEMDIFF_RANDOM_INIT", "1")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import semdiff_b200  # noqa: E402
from semdiff_b200 import sharding  # noqa: E402


def make_pairs_device(lo, hi, dev, src=512, size=224):
    gts, srs = [], []
    for i in range(lo, hi):
        g = torch.Generator(device=dev).manual_seed(77_000_000 + i)
        gt = torch.randn(1, 3, src, src, device=dev, generator=g)
        u = torch.rand((), device=dev, generator=g).item()
        sigma = math.exp(math.log(0.02) + u * (math.log(2.0) - math.log(0.02)))
        sr = (gt + sigma * torch.randn(1, 3, src, src, device=dev, generator=g)) / math.sqrt(1 + sigma * sigma)
        both = torch.nn.functional.interpolate(torch.cat([gt, sr]), size=(size, size), mode="bicubic", antialias=True, align_corners=False)
        gts.append(both[0]); srs.append(both[1])
    return torch.stack(gts), torch.stack(srs)


def spearman(a, b):
    ra, rb = a.argsort().argsort().double(), b.argsort().argsort().double()
    ra, rb = ra - ra.mean(), rb - rb.mean()
    return float((ra * rb).sum() / (ra.norm() * rb.norm()))


def inversions(ref, got):
    order = ref.argsort()
    g = got[order]
    return int((g[1:] < g[:-1]).sum())   # adjacent inversions along the reference order


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--modes", nargs="+", default=["bf16", "fp16"])
    ap.add_argument("--oracle", type=int, default=0)
    ap.add_argument("--batch", type=int, default=250)
    ap.add_argument("--skip-fp32", action="store_true")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    from oracle.restated import RestatedScorer
    from oracle.synth import set_head
    oracle = set_head(RestatedScorer("resnet50", 3, seed=0), "abs")
    result = {"pairs": args.pairs, "world": world}
    scores = {}
    for mode in (["fp32"] if not args.skip_fp32 else []) + args.modes:
        with contextlib.redirect_stdout(io.StringIO()):
            model = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb("resnet50", 3, str(dev), precision=mode).eval()
        model.load_state_dict(oracle.state_dict())
        lo, hi = sharding.shard_range(args.pairs, world, rank)
        out, t_score = [], 0.0
        with torch.no_grad():
            for b0 in range(lo, hi, args.batch):
                gt, sr = make_pairs_device(b0, min(hi, b0 + args.batch), dev)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                out.append(model(gt, sr))
                torch.cuda.synchronize(); t_score += time.perf_counter() - t0
        local_scores = torch.cat(out) if out else torch.empty(0, device=dev)
        if world > 1:
            full = sharding.gather_scores(local_scores, args.pairs)
        else:
            full = local_scores
        scores[mode] = full.cpu()
        result[mode] = {"scoring_s_rank0": t_score, "pairs_per_s_per_gpu": (hi - lo) / t_score}
    if rank == 0:
        ref = scores.get("fp32")
        for mode in args.modes:
            if ref is None:
                break
            s = scores[mode]
            result[mode].update(spearman_vs_fp32=spearman(ref, s), adjacent_inversions=inversions(ref, s),
                                max_rel_err=float(((s - ref).abs() / ref.abs().clamp_min(1e-3)).max()),
                                median_rel_err=float(((s - ref).abs() / ref.abs().clamp_min(1e-3)).median()))
        if args.oracle and ref is not None:
            k = min(args.oracle, args.pairs)
            gt, sr = make_pairs_device(0, k, dev)
            with torch.no_grad():
                o = torch.cat([oracle(gt[i:i + 16].cpu(), sr[i:i + 16].cpu()) for i in range(0, k, 16)])
            o64m = set_head(RestatedScorer("resnet50", 3, seed=0), "abs").double()
            with torch.no_grad():
                o64 = torch.cat([o64m(gt[i:i + 16].cpu().double(), sr[i:i + 16].cpu().double()) for i in range(0, k, 16)])
            rel64 = lambda x: float(((x.double() - o64).abs() / o64.abs().clamp_min(1e-3)).max())
            result["oracle_subset"] = {"pairs": k, "fp32_max_rel_err": float(((ref[:k] - o).abs() / o.abs().clamp_min(1e-3)).max()),
                                       "vs_fp64_oracle": {"reference_fp32_cpu": rel64(o), "ours_fp32": rel64(ref[:k]),
                                                          **{f"ours_{m}": rel64(scores[m][:k]) for m in args.modes}},
                                       **{f"{m}_spearman": spearman(o, scores[m][:k]) for m in args.modes},
                                       "fp32_spearman": spearman(o, ref[:k])}
        print(json.dumps(result))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
