#!/usr/bin/env python
"""Benchmark of the hot path: GT/SR pairs scored per second (224x224) by the B200 scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward(gt, sr) of the drop-in module over the workload BASELINE.json quotes the metric on
(configs[1]: ImageNet RN50 trunk, depth 3, 256 pairs of 224x224, bf16) per GPU (weak scaling; with N > 1 every rank
scores its own 256 pairs and the per-pair scores are exchanged with ONE NCCL all-gather inside the timed step).
Rank 0 prints ONE JSON line.  `value` = device-resident throughput; `e2e` = the same through the public API with
host buffers (H2D of the images and D2H of the scores inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("SEMDIFF_RANDOM_INIT", "1")   # synthetic benchmark: random-init weights of the named architecture
sys.path.insert(0, ROOT)

METRIC = "GT/SR pairs scored/sec (224^2, whole box)"
UNIT = "pairs/s"
H = W = 224
DIST_BYTES_PER_PAIR_BF16 = 6_021_120      # every tapped GT and SR activation read once, bf16, depth 3


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs.  NVML is
    initialised in the constructor (it takes ~100 ms) so that sampling covers the timed region from its start."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self._halt = [], set(), None, threading.Event()
        self._nv = self._h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def run(self):
        if self._nv is None:
            return
        nv = self._nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.reasons |= {n for bit, n in self.NAMES.items() if r & bit}
            except Exception as e:  # noqa: BLE001
                self.reasons.add(f"nvml_error:{type(e).__name__}")
                return
            time.sleep(0.005)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_throughput(budget_s: float, batch: int = 8):
    """The reference's own CPU path (oracle port; the unmodified reference file when /root/reference exists) on the
    host cores: ImageNet RN50, depth 3, fp32, batches of `batch` pairs (BASELINE.json configs[0])."""
    import torch

    from oracle import reference_loader as rl
    from oracle.restated import RestatedScorer
    from oracle.synth import make_pairs, set_head

    torch.set_num_threads(os.cpu_count())
    if rl.available():
        model, kind = set_head(rl.build_reference_scorer("resnet50", 3, seed=0), "abs"), "reference"
    else:
        model, kind = set_head(RestatedScorer("resnet50", 3, seed=0), "abs"), "port"
    gt, sr = make_pairs(batch, seed=0)
    times = []
    with torch.no_grad():
        for _ in range(2):
            model(gt, sr)
        t_end = time.time() + budget_s
        while time.time() < t_end or len(times) < 3:
            t0 = time.perf_counter()
            model(gt, sr)
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": batch / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{len(times)} forwards of {batch} pairs (224x224 fp32, RN50 depth 3), median; best {batch / times[0]:.1f}"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import torch

    from oracle import reference_loader as rl
    from oracle.restated import RestatedScorer
    from oracle.synth import make_pairs, set_head

    torch.set_num_threads(os.cpu_count())
    batch = 8
    if rl.available():
        model, kind = set_head(rl.build_reference_scorer("resnet50", 3, seed=0), "abs"), "reference"
    else:
        model, kind = set_head(RestatedScorer("resnet50", 3, seed=0), "abs"), "port"
    gt, sr = make_pairs(batch, seed=0)
    with torch.no_grad():
        for _ in range(args.warmup):
            model(gt, sr)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model(gt, sr)
        dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    sample = f"each step = {batch} pairs of the 256-pair workload (bounded CPU sample), fp32, all host threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "CLIP-LPIPS regressor, ImageNet RN50 trunk (random init), depth 3, 256 pairs 224x224 per GPU",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per GPU per step")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--trunk", default="resnet50")
    ap.add_argument("--microbatch", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import semdiff_b200
    from semdiff_b200 import sharding, trunks

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()

    cls = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb if args.trunk == "resnet50" else semdiff_b200.CLIP_lpips_stages_cnn
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = cls(clip_name=args.trunk, depth=3, device=str(dev), precision=args.precision,
                    microbatch=args.microbatch or None).eval()
    with torch.no_grad():
        for m in model.w_layers:
            m.weight.abs_()
            m.bias.abs_()
    n = args.pairs
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    gt = torch.randn(n, 3, H, W, device=dev, generator=g)
    sr = gt + 0.1 * torch.randn(n, 3, H, W, device=dev, generator=g)
    total_pairs = n * world

    def step():
        with torch.no_grad():
            s = model(gt, sr)
            if world > 1:
                s = sharding.gather_scores(s, total_pairs)
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        scores = step()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        scores = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = model.plan().last_launches() * args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = total_pairs * args.steps / (ms / 1e3)

    # ---- e2e: host buffers in, host scores out, through the public API -------------------------
    gt_h = torch.empty(n, 3, H, W, pin_memory=True).copy_(gt)
    sr_h = torch.empty(n, 3, H, W, pin_memory=True).copy_(sr)
    outs_h = [torch.empty(n, pin_memory=True) for _ in range(2)]
    e2e_steps = max(4, args.steps // 2)

    def e2e_loop(k, g_h=None, s_h=None):
        """k steps, two in flight: step i+1's images cross PCIe while step i is scored; every step's scores are read
        back to the host and waited for."""
        g_h, s_h = (gt_h, sr_h) if g_h is None else (g_h, s_h)
        pending = None
        for i in range(k):
            _, ev = model.score_host(g_h, s_h, outs_h[i % 2], wait=False)
            if pending is not None:
                pending.synchronize()
            pending = ev
        pending.synchronize()

    e2e_loop(3)
    barrier()
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e = {"value": total_pairs * e2e_steps / (ms_e2e / 1e3), "unit": UNIT,
           "h2d_bytes_per_step": 2 * gt_h.numel() * 4, "d2h_bytes_per_step": outs_h[0].numel() * 4,
           "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
           "how": "model.score_host(pinned fp32 gt, pinned fp32 sr) -> pinned scores, two steps in flight: the H2D copy "
                  "of step i+1 (copy stream, 2 staging slots) overlaps the scoring of step i; PCIe-bound above ~43k pairs/s "
                  "(308 MB per 256 pairs at the measured 51.7 GB/s)"}
    ref_scores = scores[rank * n:(rank + 1) * n].cpu() if world > 1 else scores.cpu()
    assert torch.equal(outs_h[0], ref_scores) and torch.equal(outs_h[1], ref_scores), "e2e result differs"
    # extra (not the headline): the same loop when the data loader already hands over 16-bit images
    if args.precision in ("bf16", "fp16"):
        dt16 = torch.bfloat16 if args.precision == "bf16" else torch.float16
        g16 = torch.empty(n, 3, H, W, dtype=dt16, pin_memory=True).copy_(gt)
        s16 = torch.empty(n, 3, H, W, dtype=dt16, pin_memory=True).copy_(sr)
        e2e_loop(2, g16, s16)
        barrier()
        e0.record()
        e2e_loop(e2e_steps, g16, s16)
        e1.record()
        barrier()
        ms16 = e0.elapsed_time(e1)
        e2e["with_16bit_host_images"] = {"value": total_pairs * e2e_steps / (ms16 / 1e3), "unit": UNIT,
                                         "h2d_bytes_per_step": 2 * g16.numel() * 2,
                                         "note": "same values already rounded to the trunk's 16-bit type on the host; scores identical"}
        assert torch.equal(outs_h[0], ref_scores), "16-bit-input e2e result differs"

    # ---- roofline of the dominant kernel family (tcgen05 implicit-GEMM convs), per-op CUDA events ----
    plan = model.plan()
    plan.set_profiling(True)
    prof_steps = 3
    step()
    plan.profile(reset=True)
    for _ in range(prof_steps):
        step()
    op_ms, op_cnt = plan.profile(reset=True)
    plan.set_profiling(False)
    ops = plan.program.ops
    conv_ms = sum(op_ms[i] for i, op in enumerate(ops) if op["kind"] == 0) / prof_steps
    conv_launches = sum(op_cnt[i] for i, op in enumerate(ops) if op["kind"] == 0) / prof_steps
    other_ms = sum(op_ms) / prof_steps - conv_ms
    dist_ms = op_ms[len(ops) + 1] / prof_steps
    flops_step = trunks.conv_flops(plan.program, H, W) * 2 * n
    achieved = flops_step / (conv_ms / 1e3) / 1e12
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("conv_tc_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"kernel": "conv_tc_kernel<*> + conv3x3_strip_kernel<*> + conv_chain_kernel<*> (tcgen05 implicit GEMM: every conv launch of the trunk; the stem launch includes the fused max pool)",
                "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "frac_of_burst_peak": achieved / peaks["bf16_burst"],
                "traffic": traffic, "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": conv_launches, "avg_launch_ms": conv_ms / max(conv_launches, 1),
                "algorithmic_flops_per_step": flops_step, "conv_ms_per_step": conv_ms,
                "share_of_step": conv_ms / (conv_ms + other_ms),
                "timing": f"per-op CUDA events on the launch stream over {prof_steps} extra steps identical to the timed ones"}
    dist_bytes = DIST_BYTES_PER_PAIR_BF16 * (2 if args.precision == "fp32" else 1) * n
    roofline_distance = {"kernel": "distance_kernel (fused per-layer distance)", "bound": "hbm",
                         "in_step": {"achieved": dist_bytes / (dist_ms / 1e3) / 1e9 if dist_ms > 0 else None, "unit": "GB/s",
                                     "note": "inside the step the taps are L2-resident (micro-batching), so this is not an HBM figure"}}
    roofline_distance.update(distance_isolated(model, peaks))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"CLIP-LPIPS regressor, ImageNet RN50 trunk (random init), depth 3, {n} pairs 224x224 per GPU"
                       if args.trunk == "resnet50" else f"CLIP-LPIPS regressor, CLIP-RN50 trunk (random init), depth 3, {n} pairs 224x224 per GPU",
                       "pairs_per_gpu": n, "microbatch_pairs": min(model.default_microbatch(H, W), n), "precision": args.precision,
                       "l2": "inputs (2 x %d MB fp32 per step) are larger than the 126 MB L2; no flush needed" % (gt.numel() * 4 >> 20),
                       "collective": "one all_gather_into_tensor of fp32 scores per step" if world > 1 else "none (1 GPU)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "roofline_distance": roofline_distance}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_throughput(args.cpu_budget)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def distance_isolated(model, peaks):
    """The distance kernel alone on tap-sized activations larger than L2 (256 pairs x layer1 taps = 822 MB in bf16):
    achieved HBM GB/s against the measured copy bandwidth."""
    import torch

    from semdiff_b200 import _lib

    lib = _lib.load()
    prec = model.plan().precision
    dt = {0: torch.bfloat16, 1: torch.float16, 2: torch.float32}[prec]
    n_pairs, hw, c = 256, 56 * 56, 256
    act = torch.randn(2 * n_pairs, hw, c, device="cuda", dtype=dt)
    w = torch.rand(c, device="cuda")
    partial = torch.empty(n_pairs, _lib.MAX_PARTS, device="cuda")
    args = (act.data_ptr(), n_pairs, hw, c, w.data_ptr(), 0, partial.data_ptr(), None, 0, prec)
    for _ in range(3):
        _lib.check(lib.semdiff_layer_distance(*args, _lib.stream_ptr()), "distance")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        lib.semdiff_layer_distance(*args, _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = act.numel() * act.element_size()
    gbs = nbytes / (ms / 1e3) / 1e9
    return {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "algorithmic_bytes_per_launch": nbytes, "avg_launch_ms": ms,
            "isolated": "layer1-shaped taps of 256 pairs (822 MB bf16 > L2), 10 launches, CUDA events",
            "peak_source": peaks["source"]}


if __name__ == "__main__":
    main()
