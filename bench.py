#!/usr/bin/env python
"""Benchmark of the hot path: GT/SR pairs scored per second by the B200 scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload pairs224|sweep10k|hires1024]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

--workload pairs224 (default, the headline): one "step" = one forward(gt, sr) of the drop-in module over the workload
BASELINE.json quotes the metric on (configs[1]: ImageNet RN50 trunk, depth 3, 256 pairs of 224x224, bf16) per GPU (weak
scaling; with N > 1 every rank scores its own 256 pairs and the per-pair scores are exchanged with ONE NCCL all-gather
inside the timed step).  Rank 0 prints ONE JSON line.  `value` = device-resident throughput; `e2e` = the same through
the public API with host buffers (H2D of the images and D2H of the scores inside the timed region) next to the measured
pure-copy ceiling of the same bytes; `fp16x3` = the same two numbers for the split-precision mode that holds the
reference's fp32 tolerance on tensor cores.
--workload sweep10k  = BASELINE.json configs[3] (10k synthetic pairs, 512x512 sources resized to 224 on the device, sharded
over the ranks, one all-gather; the gathered scores are compared bit for bit with rank 0 scoring alone).
--workload hires1024 = BASELINE.json configs[4] (32 pairs of 1024x1024 over the ranks).
"""
from __future__ import annotations

import argparse
import contextlib
import hashlib
import io
import json
import math
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
CSRC = os.path.join(ROOT, "measuring-semantic-differences-in-the-super-resolution-domain_b200", "csrc")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def kernel_sources_digest() -> str:
    """sha256 over the CUDA sources: profiles/roofline_traffic.json is stamped with it when ncu captured it, and its
    DRAM-traffic figure is only quoted while the kernels are still the ones that were profiled."""
    h = hashlib.sha256()
    for fn in sorted(os.listdir(CSRC)):
        if fn.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, fn), "rb") as f:
                h.update(fn.encode() + b"\0" + f.read())
    return h.hexdigest()


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


# ---------------------------------------------------------------------------------------------
# CPU reference (oracle) legs
# ---------------------------------------------------------------------------------------------
def _cpu_model():
    import torch

    from oracle import reference_loader as rl
    from oracle.restated import RestatedScorer
    from oracle.synth import set_head

    torch.set_num_threads(os.cpu_count())
    if rl.available():
        return set_head(rl.build_reference_scorer("resnet50", 3, seed=0), "abs"), "reference"
    return set_head(RestatedScorer("resnet50", 3, seed=0), "abs"), "port"


def cpu_reference_throughput(budget_s: float, batch: int = 8):
    """The reference's own CPU path (oracle port; the unmodified reference file when /root/reference exists) on the
    host cores: ImageNet RN50, depth 3, fp32, batches of `batch` pairs (BASELINE.json configs[0])."""
    import torch

    from oracle.synth import make_pairs

    model, kind = _cpu_model()
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

    from oracle.synth import make_pairs

    model, kind = _cpu_model()
    batch = 8
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


# ---------------------------------------------------------------------------------------------
# helpers shared by the GPU workloads
# ---------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device / process-group plumbing of one benchmark process."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        from semdiff_b200 import sharding

        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
        self.numa = sharding.bind_to_gpu_numa_node(self.local_rank)   # before any pinned allocation
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.torch, self.dist = torch, dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def timed(self, fn, steps: int) -> float:
        """ms for `steps` calls of fn: device-timed on the launch stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def build_model(ctx, trunk: str, precision: str, microbatch: int | None = None, state=None):
    import semdiff_b200

    cls = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb if trunk == "resnet50" else semdiff_b200.CLIP_lpips_stages_cnn
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = cls(clip_name=trunk, depth=3, device=str(ctx.dev), precision=precision, microbatch=microbatch or None).eval()
    if state is not None:
        model.load_state_dict(state)
    else:
        with ctx.torch.no_grad():
            for m in model.w_layers:
                m.weight.abs_()
                m.bias.abs_()
    return model


def copy_ceiling(ctx, host_tensors, steps: int):
    """Pure pinned-copy probe: the step's host buffers -> device, one cudaMemcpyAsync per buffer, nothing else; all ranks
    at once, so at N > 1 it measures what the host's memory system and the PCIe switches deliver to N GPUs together."""
    torch = ctx.torch
    dst = [torch.empty_like(t, device=ctx.dev) for t in host_tensors]

    def once():
        for d, s in zip(dst, host_tensors):
            d.copy_(s, non_blocking=True)

    for _ in range(2):
        once()
    ms = ctx.timed(once, steps)
    nbytes = sum(t.numel() * t.element_size() for t in host_tensors)
    return ms / steps, nbytes


def e2e_run(ctx, model, g_h, s_h, outs_h, steps: int, pairs_total: int):
    """`steps` calls of model.score_host with three in flight (its staging ring has three slots): the images of steps i+1
    and i+2 cross PCIe / are queued while step i is scored; every step's scores are read back to the host and waited for."""
    def loop(k):
        pending = []
        for i in range(k):
            _, ev = model.score_host(g_h, s_h, outs_h[i % len(outs_h)], wait=False)
            pending.append(ev)
            if len(pending) > 2:
                pending.pop(0).synchronize()
        for ev in pending:
            ev.synchronize()

    loop(5)
    ms = ctx.timed(lambda: loop(steps), 1)
    nbytes = 2 * g_h.numel() * g_h.element_size()
    res = {"value": pairs_total * steps / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": nbytes,
           "d2h_bytes_per_step": outs_h[0].numel() * 4, "steps": steps, "ms_per_step": ms / steps}
    c_ms, c_bytes = copy_ceiling(ctx, [g_h, s_h], max(4, steps))
    res["ceiling"] = {"value": pairs_total / (c_ms / 1e3), "unit": UNIT, "h2d_gbs_per_gpu": c_bytes / (c_ms / 1e3) / 1e9,
                      "frac": res["value"] / (pairs_total / (c_ms / 1e3)), "pinned": bool(g_h.is_pinned() and s_h.is_pinned()),
                      "how": "the same pinned buffers copied host->device with nothing else running, all ranks concurrently (pure PCIe / host-DRAM ceiling)"}
    return res


# ---------------------------------------------------------------------------------------------
# workload: pairs224 (BASELINE.json configs[1] / configs[2]) - the headline
# ---------------------------------------------------------------------------------------------
def run_pairs224(args, ctx):
    torch = ctx.torch
    from semdiff_b200 import sharding, trunks

    peaks = load_peaks()
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    model = build_model(ctx, args.trunk, args.precision, args.microbatch)
    n = args.pairs
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    gt = torch.randn(n, 3, H, W, device=dev, generator=g)
    sr = gt + 0.1 * torch.randn(n, 3, H, W, device=dev, generator=g)
    total_pairs = n * world

    def make_step(m):
        def step():
            with torch.no_grad():
                s = m(gt, sr)
                if world > 1:
                    s = sharding.gather_scores(s, total_pairs)
            return s
        return step

    step = make_step(model)
    sampler = ClockSampler(ctx.local_rank)
    for _ in range(args.warmup):
        scores = step()
    ctx.barrier()
    sampler.start()
    ms = ctx.timed(step, args.steps)
    clocks = sampler.stop()
    scores = step()
    launches = model.plan().last_launches() * args.steps
    value = total_pairs * args.steps / (ms / 1e3)

    # ---- e2e: host buffers in, host scores out, through the public API -------------------------
    gt_h = torch.empty(n, 3, H, W, pin_memory=True).copy_(gt)
    sr_h = torch.empty(n, 3, H, W, pin_memory=True).copy_(sr)
    outs_h = [torch.empty(n, pin_memory=True) for _ in range(3)]
    e2e_steps = max(20, args.steps)   # the copy / compute pipeline needs a few steps to settle: never fewer than 20
    e2e = e2e_run(ctx, model, gt_h, sr_h, outs_h, e2e_steps, total_pairs)
    e2e["how"] = ("model.score_host(pinned fp32 gt, pinned fp32 sr) -> pinned scores, three steps in flight: the H2D copies of steps i+1 / i+2 "
                  "(copy stream, 3 staging slots) overlap the scoring of step i; `ceiling` = the same bytes copied with nothing else running")
    e2e["numa_binding"] = ctx.numa
    e2e["frac_of_device_value"] = e2e["value"] / value
    ref_scores = scores[rank * n:(rank + 1) * n].cpu() if world > 1 else scores.cpu()
    assert all(torch.equal(o, ref_scores) for o in outs_h), "e2e result differs"
    # compact host inputs: what a data loader that keeps 16-bit tensors / decoded uint8 images hands over
    variants = {}
    if args.precision in ("bf16", "fp16"):
        dt16 = torch.bfloat16 if args.precision == "bf16" else torch.float16
        g16 = torch.empty(n, 3, H, W, dtype=dt16, pin_memory=True).copy_(gt)
        s16 = torch.empty(n, 3, H, W, dtype=dt16, pin_memory=True).copy_(sr)
        v = e2e_run(ctx, model, g16, s16, outs_h, e2e_steps, total_pairs)
        v["note"] = "same values already rounded to the trunk's 16-bit type on the host; scores identical"
        v["frac_of_device_value"] = v["value"] / value
        assert torch.equal(outs_h[0], ref_scores), "16-bit-input e2e result differs"
        variants["host_images_16bit"] = v
        del g16, s16
    gu = torch.empty(n, H, W, 3, dtype=torch.uint8, pin_memory=True).random_(0, 256, generator=torch.Generator().manual_seed(7 + rank))
    su = torch.empty_like(gu, pin_memory=True).copy_(gu).add_(torch.randint(0, 6, gu.shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(9 + rank))).clamp_(max=250)
    v = e2e_run(ctx, model, gu, su, outs_h, e2e_steps, total_pairs)
    v["note"] = ("decoded uint8 [N,224,224,3] host images; the reference's model.processor (resize 235 bicubic, center crop 224, normalise) "
                 "runs on the device, bit-exact (csrc/preprocess.cu), inside the timed region")
    v["frac_of_device_value"] = v["value"] / value
    variants["host_images_uint8"] = v
    e2e["variants"] = variants
    del gu, su

    # ---- the split-precision mode (reference fp32 tolerance on tensor cores): same step, same e2e ----
    x3 = None
    if not args.no_x3 and args.precision != "fp16x3":
        mx = build_model(ctx, args.trunk, "fp16x3", args.microbatch, state=model.state_dict())
        stepx = make_step(mx)
        for _ in range(3):
            sx = stepx()
        x3_steps = max(4, args.steps // 2)
        msx = ctx.timed(stepx, x3_steps)
        ex = e2e_run(ctx, mx, gt_h, sr_h, outs_h, max(10, x3_steps), total_pairs)
        px = mx.plan()
        px.set_profiling(True)
        stepx()
        px.profile(reset=True)
        for _ in range(2):
            stepx()
        x_ms, _ = px.profile(reset=True)
        px.set_profiling(False)
        x_conv_ms = sum(x_ms[i] for i, op in enumerate(px.program.ops) if op["kind"] == 0) / 2
        x_alg = trunks.conv_flops(px.program, H, W) * 2 * n / (x_conv_ms / 1e3) / 1e12
        sx_local = sx[rank * n:(rank + 1) * n] if world > 1 else sx
        dev_rel = ((scores[rank * n:(rank + 1) * n] if world > 1 else scores) - sx_local).abs() / sx_local.abs().clamp_min(1e-3)
        x3 = {"precision": "fp16x3", "value": total_pairs * x3_steps / (msx / 1e3), "unit": UNIT, "ms_per_step": msx / x3_steps,
              "steps": x3_steps, "e2e": ex, "gpu_launches_per_step": mx.plan().last_launches(),
              "what": "every activation / weight = hi + lo fp16 pair, three tcgen05 products per K block, chunk sums promoted to "
                      "registers; <= 1e-5 of the fp32 oracle (tests/test_scorer_gpu.py)",
              "roofline": {"bound": "tensor", "achieved": x_alg, "unit": "TFLOP/s of algorithmic work (conv launches, per-op CUDA events)",
                           "tensor_core_tflops": 3 * x_alg, "frac_of_sustained_peak": 3 * x_alg / peaks["bf16_sustained"],
                           "frac_of_burst_peak": 3 * x_alg / peaks["bf16_burst"],
                           "note": "three fp16 tensor-core products per algorithmic multiply-add; 4-byte activations make the 56x56 / 28x28 1x1 layers HBM-bound"},
              "headline_mode_vs_this_mode_max_rel_diff": float(dev_rel.max())}
        del mx

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
    traffic, traffic_note = None, "no ncu capture on file"
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("kernel_sources_sha256") == kernel_sources_digest():
            traffic, traffic_note = tj.get("conv_tc_dram_bytes_per_launch"), f"ncu capture {tj.get('source')} of exactly these kernel sources"
        else:
            traffic_note = f"{tj.get('source')} was captured from different kernel sources (stamp mismatch): not quoted"
    except Exception:  # noqa: BLE001
        pass
    # which peak: the burst figure if the clock stayed at its maximum through the timed region, else the sustained one
    at_max = clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"] \
        and "sw_power_cap" not in clocks.get("reasons", [])
    peak = peaks["bf16_burst"] if at_max else peaks["bf16_sustained"]
    roofline = {"kernel": "conv_tc_kernel<*> + conv3x3_strip_kernel<*> + conv_chain_kernel<*> (tcgen05 implicit GEMM: every conv launch of the trunk; the stem launch includes the fused max pool)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "frac_of_burst_peak": achieved / peaks["bf16_burst"], "frac_of_sustained_peak": achieved / peaks["bf16_sustained"],
                "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peaks["source"] + (", burst bf16 (SM clock sampled at its maximum, no power cap, during the timed region)" if at_max
                                                  else ", sustained bf16 (SM clock below maximum / power cap during the timed region)"),
                "launches_per_step": conv_launches, "avg_launch_ms": conv_ms / max(conv_launches, 1),
                "algorithmic_flops_per_step": flops_step, "conv_ms_per_step": conv_ms,
                "share_of_step": conv_ms / (conv_ms + other_ms),
                "timing": f"per-op CUDA events on the launch stream over {prof_steps} extra steps identical to the timed ones"}
    dist_bytes = DIST_BYTES_PER_PAIR_BF16 * (2 if args.precision in ("fp32", "fp16x3", "bf16x3") else 1) * n
    roofline_distance = {"kernel": "distance_kernel (fused per-layer distance)", "bound": "hbm",
                         "in_step": {"achieved": dist_bytes / (dist_ms / 1e3) / 1e9 if dist_ms > 0 else None, "unit": "GB/s",
                                     "note": "inside the step part of the taps is still L2-resident, so this is not an HBM figure"}}
    roofline_distance.update(distance_isolated(model, peaks))

    trunk_name = "ImageNet RN50 trunk" if args.trunk == "resnet50" else "CLIP-RN50 trunk"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"CLIP-LPIPS regressor, {trunk_name} (random init), depth 3, {n} pairs 224x224 per GPU",
                       "pairs_per_gpu": n, "microbatch_pairs": min(model.default_microbatch(H, W), n), "precision": args.precision,
                       "l2": "inputs (2 x %d MB fp32 per step) are larger than the 126 MB L2; no flush needed" % (gt.numel() * 4 >> 20),
                       "collective": "one all_gather_into_tensor of fp32 scores per step" if world > 1 else "none (1 GPU)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "roofline_distance": roofline_distance}
    if x3 is not None:
        line["fp16x3"] = x3
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_throughput(args.cpu_budget)
    if rank == 0:
        print(json.dumps(line if args.detail else compact_line(line)), flush=True)


def compact_line(line: dict) -> dict:
    """The contract's keys with their numbers and short labels (< 3 KB on one line, so that a log tail always holds it whole);
    `--detail` prints every note and sub-measurement instead (that is what the records under profiles/ are)."""
    def pick(d, keys):
        return {k: d[k] for k in keys if k in d}
    out = pick(line, ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                      "dtype", "data", "gpu_launches", "clocks"))
    out["config"] = pick(line["config"], ("workload", "pairs_per_gpu", "microbatch_pairs", "precision", "collective"))
    out["config"]["l2"] = "inputs larger than L2, no flush needed"
    e = line["e2e"]
    out["e2e"] = pick(e, ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "steps", "ms_per_step", "frac_of_device_value"))
    out["e2e"]["ceiling"] = pick(e["ceiling"], ("value", "h2d_gbs_per_gpu", "frac"))
    out["e2e"]["variants"] = {k: pick(v, ("value", "h2d_bytes_per_step")) for k, v in e.get("variants", {}).items()}
    r = line["roofline"]
    out["roofline"] = pick(r, ("bound", "achieved", "peak", "unit", "frac", "frac_of_burst_peak", "frac_of_sustained_peak", "traffic",
                               "launches_per_step", "avg_launch_ms", "share_of_step"))
    out["roofline"]["kernel"] = "conv_tc + conv3x3_strip + conv_chain (tcgen05 implicit-GEMM convs of the trunk)"
    out["roofline"]["peak_source"] = "MEASURED_PEAKS.json" if "measured" in r["peak_source"] else "fallback"
    out["roofline_distance"] = pick(line["roofline_distance"], ("bound", "achieved", "peak", "unit", "frac"))
    if "fp16x3" in line:
        x = line["fp16x3"]
        out["fp16x3"] = pick(x, ("value", "unit", "ms_per_step", "steps"))
        out["fp16x3"]["e2e"] = pick(x["e2e"], ("value", "h2d_bytes_per_step", "d2h_bytes_per_step"))
        out["fp16x3"].update(pick(x.get("roofline", {}), ("tensor_core_tflops", "frac_of_sustained_peak")))
    if "cpu_baseline" in line:
        out["cpu_baseline"] = pick(line["cpu_baseline"], ("value", "unit", "cores", "kind"))
        out["cpu_baseline"]["sample"] = line["cpu_baseline"]["sample"][:90]
    return out


def distance_isolated(model, peaks):
    """The distance kernel alone on tap-sized activations larger than L2 (256 pairs x layer1 taps = 822 MB in bf16):
    achieved HBM GB/s against the measured copy bandwidth."""
    import torch

    from semdiff_b200 import _lib

    lib = _lib.load()
    prec = model.plan().precision
    dt = {0: torch.bfloat16, 1: torch.float16, 2: torch.float32, 3: torch.float16, 4: torch.bfloat16}[prec]
    n_pairs, hw, c = 256, 56 * 56, 256
    act = torch.randn(2 * n_pairs, hw, c * (2 if prec >= 3 else 1), device="cuda", dtype=dt)
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
            "isolated": "layer1-shaped taps of 256 pairs (> L2), 10 launches, CUDA events",
            "peak_source": peaks["source"]}


# ---------------------------------------------------------------------------------------------
# workload: sweep10k (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------
SWEEP_BLOCK = 50   # pairs per generator seed: shards regenerate whole blocks and slice, so any rank layout sees the same pairs


def sweep_pairs_device(lo, hi, dev, src=512, size=224):
    """Pairs [lo, hi) of the synthetic sweep, generated on the device from per-block seeds (any shard regenerates exactly its
    own pairs): GT ~ N(0,1) at 512x512, SR = (GT + sigma noise) / sqrt(1 + sigma^2), sigma log-uniform in [0.02, 2]; both
    resized to 224 (bicubic, antialiased) on the device - the SR-outputs-dataset shape of configs[3]."""
    import torch

    gts, srs = [], []
    for blk in range(lo // SWEEP_BLOCK, (hi - 1) // SWEEP_BLOCK + 1):
        g = torch.Generator(device=dev).manual_seed(77_000_000 + blk)
        gt = torch.randn(SWEEP_BLOCK, 3, src, src, device=dev, generator=g)
        u = torch.rand(SWEEP_BLOCK, device=dev, generator=g)
        sigma = torch.exp(math.log(0.02) + u * (math.log(2.0) - math.log(0.02))).view(-1, 1, 1, 1)
        sr = (gt + sigma * torch.randn(SWEEP_BLOCK, 3, src, src, device=dev, generator=g)) / torch.sqrt(1 + sigma * sigma)
        a, b = max(lo, blk * SWEEP_BLOCK) - blk * SWEEP_BLOCK, min(hi, (blk + 1) * SWEEP_BLOCK) - blk * SWEEP_BLOCK
        kw = dict(size=(size, size), mode="bicubic", antialias=True, align_corners=False)
        gts.append(torch.nn.functional.interpolate(gt[a:b], **kw))
        srs.append(torch.nn.functional.interpolate(sr[a:b], **kw))
    return torch.cat(gts), torch.cat(srs)


def spearman(a, b):
    ra, rb = a.argsort().argsort().double(), b.argsort().argsort().double()
    ra, rb = ra - ra.mean(), rb - rb.mean()
    return float((ra * rb).sum() / (ra.norm() * rb.norm()))


def inversions(ref, got):
    """adjacent inversions of `got` along the order of `ref`"""
    g = got[ref.argsort()]
    return int((g[1:] < g[:-1]).sum())


def run_sweep10k(args, ctx):
    """All pairs are generated once per rank (device-resident shard), then each precision mode scores its shard `reps` times:
    the first pass includes plan building / descriptor encoding (reported as `first_pass_s`), the others are steady state.
    Sharding contract (sharding.py): contiguous blocks, tail padded, ONE all-gather of fp32 scores per sweep."""
    torch = ctx.torch
    from oracle.restated import RestatedScorer
    from oracle.synth import set_head
    from semdiff_b200 import sharding

    P, world, rank, dev = args.pairs if args.pairs != 256 else 10000, ctx.world, ctx.rank, ctx.dev
    oracle = set_head(RestatedScorer("resnet50", 3, seed=0), "abs")
    lo, hi = sharding.shard_range(P, world, rank)
    t0 = time.perf_counter()
    gt, sr = sweep_pairs_device(lo, hi, dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    alone_in = None
    if world > 1 and rank == 0:   # every pair once more on rank 0, for the 1-vs-N bit-equality check of each mode
        alone_in = sweep_pairs_device(0, P, dev)
    result = {"workload": "sweep10k", "pairs": P, "n_gpus": world, "pairs_per_rank": hi - lo, "generation_s_rank0": gen_s,
              "ragged": f"{hi - lo} pairs per rank = {(hi - lo) // 512} full micro-batches of 512 + {(hi - lo) % 512} (16-bit modes)", "modes": {}}
    scores = {}
    modes = args.modes.split(",")
    for mode in modes:
        model = build_model(ctx, "resnet50", mode, state=oracle.state_dict())

        def sweep():
            with torch.no_grad():
                local = model(gt, sr)
                return sharding.gather_scores(local, P) if world > 1 else local

        ctx.barrier()
        t0 = time.perf_counter()
        full = sweep()
        ctx.barrier()
        first_s = time.perf_counter() - t0
        reps = 3
        ms = ctx.timed(sweep, reps) / reps
        full2 = sweep()
        assert torch.equal(full, full2), "a sweep must be reproducible bit for bit"
        scores[mode] = full
        result["modes"][mode] = {"first_pass_s": first_s, "steady_ms_per_sweep": ms, "pairs_per_s": P / (ms / 1e3),
                                 "pairs_per_s_per_gpu": P / (ms / 1e3) / world}
        if world > 1:
            # rank 0 scores EVERY pair alone (regenerated from the seeds) and compares with the gathered result
            if rank == 0:
                with torch.no_grad():
                    alone = model(*alone_in)
                result["modes"][mode]["gathered_equals_rank0_alone"] = bool(torch.equal(alone, full))
                result["modes"][mode]["max_abs_diff_vs_rank0_alone"] = float((alone - full).abs().max())
            ctx.barrier()
        del model
    if rank == 0:
        ref_mode = "fp32" if "fp32" in scores else ("fp16x3" if "fp16x3" in scores else None)
        for mode in modes:
            if ref_mode is None or mode == ref_mode:
                continue
            ref, s = scores[ref_mode].cpu(), scores[mode].cpu()
            rel = (s - ref).abs() / ref.abs().clamp_min(1e-3)
            result["modes"][mode].update({f"spearman_vs_{ref_mode}": spearman(ref, s), f"adjacent_inversions_vs_{ref_mode}": inversions(ref, s),
                                          f"identical_order_vs_{ref_mode}": bool(torch.equal(ref.argsort(stable=True), s.argsort(stable=True))),
                                          "max_rel_err": float(rel.max()), "median_rel_err": float(rel.median())})
        k = min(args.oracle, P)
        if k > 0:
            g1, s1 = sweep_pairs_device(0, k, dev)
            g1, s1 = g1.cpu(), s1.cpu()
            o64m = set_head(RestatedScorer("resnet50", 3, seed=0), "abs").double()
            with torch.no_grad():
                o = torch.cat([oracle(g1[i:i + 16], s1[i:i + 16]) for i in range(0, k, 16)])
                o64 = torch.cat([o64m(g1[i:i + 16].double(), s1[i:i + 16].double()) for i in range(0, k, 16)])
            rel64 = lambda x: float(((x.double() - o64).abs() / o64.abs().clamp_min(1e-3)).max())  # noqa: E731
            result["oracle_subset"] = {"pairs": k, "reference_fp32_cpu_vs_fp64": rel64(o), "reference_fp32_inversions_vs_fp64": inversions(o64, o.double()),
                                       **{f"{m}_vs_fp64": rel64(scores[m][:k].cpu()) for m in modes},
                                       **{f"{m}_vs_reference_fp32": float(((scores[m][:k].cpu() - o).abs() / o.abs().clamp_min(1e-3)).max()) for m in modes},
                                       **{f"{m}_inversions_vs_fp64": inversions(o64, scores[m][:k].cpu().double()) for m in modes}}
        print(json.dumps(result), flush=True)


# ---------------------------------------------------------------------------------------------
# workload: hires1024 (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------
def run_hires1024(args, ctx):
    """32 pairs of 1024x1024 (taps 256^2 .. 32^2) sharded over the ranks: 4 pairs per GPU on 8 GPUs.  Timed like the
    headline: K steps of the whole 32-pair batch, one all-gather per step, max over ranks."""
    torch = ctx.torch
    from oracle.restated import RestatedScorer
    from oracle.synth import set_head
    from semdiff_b200 import sharding

    P, world, rank, dev, S = (args.pairs if args.pairs != 256 else 32), ctx.world, ctx.rank, ctx.dev, 1024
    oracle = set_head(RestatedScorer("resnet50", 3, seed=0), "abs")
    lo, hi = sharding.shard_range(P, world, rank)
    gts, srs = [], []
    for i in range(lo, hi):
        g = torch.Generator(device=dev).manual_seed(55_000 + i)
        gt = torch.randn(1, 3, S, S, device=dev, generator=g)
        sigma = 0.02 * (100.0 ** (i / max(P - 1, 1)))
        srs.append((gt + sigma * torch.randn(1, 3, S, S, device=dev, generator=g)) / math.sqrt(1 + sigma * sigma))
        gts.append(gt)
    gt, sr = torch.cat(gts), torch.cat(srs)
    result = {"workload": "hires1024", "pairs": P, "n_gpus": world, "pairs_per_rank": hi - lo, "modes": {}}
    scores = {}
    for mode in args.modes.split(","):
        model = build_model(ctx, "resnet50", mode, state=oracle.state_dict())

        def step():
            with torch.no_grad():
                local = model(gt, sr)
                return sharding.gather_scores(local, P) if world > 1 else local

        for _ in range(max(args.warmup, 3)):
            full = step()
        ms = ctx.timed(step, args.steps) / args.steps
        scores[mode] = full.cpu()
        result["modes"][mode] = {"ms_per_step": ms, "pairs_per_s": P / (ms / 1e3), "launches_per_step": model.plan().last_launches(),
                                 "trunk_tflops": 341.651e9 * P / (ms / 1e3) / 1e12}
        del model
    if rank == 0:
        if args.oracle > 0:   # pair 0 through the CPU oracle (fp32 and fp64): ~10 s
            with torch.no_grad():
                o = oracle(gt[:1].cpu(), sr[:1].cpu())
                o64 = set_head(RestatedScorer("resnet50", 3, seed=0), "abs").double()(gt[:1].cpu().double(), sr[:1].cpu().double())
            result["oracle_pair0"] = {"oracle_fp32": float(o), "oracle_fp64": float(o64),
                                      **{m: float(s[0]) for m, s in scores.items()},
                                      **{f"{m}_rel_err_vs_fp32_oracle": float((s[0] - o[0]).abs() / o[0].abs()) for m, s in scores.items()}}
        ref = scores.get("fp16x3")
        if ref is not None:
            for m, s in scores.items():
                if m != "fp16x3":
                    result["modes"][m]["max_rel_diff_vs_fp16x3"] = float(((s - ref).abs() / ref.abs().clamp_min(1e-3)).max())
        print(json.dumps(result), flush=True)


# ---------------------------------------------------------------------------------------------
# workload: unet (the local-map U-Net, /root/reference/models/local_eval_models.py:7-339)
# ---------------------------------------------------------------------------------------------
def run_unet(args, ctx):
    """Local semantic-difference maps per second: 64 pairs of 224x224 per GPU per step (trunk 2 x 8.2 GFLOP + decoder
    56 GFLOP per pair), device-resident inputs, no collective (maps stay on their rank)."""
    torch = ctx.torch
    import warnings

    import semdiff_b200

    n = args.pairs if args.pairs != 256 else 64
    g = torch.Generator(device=ctx.dev).manual_seed(99 + ctx.rank)
    gt = torch.randn(n, 3, H, W, device=ctx.dev, generator=g)
    sr = gt + 0.1 * torch.randn(n, 3, H, W, device=ctx.dev, generator=g)
    result = {"workload": "unet", "pairs_per_gpu": n, "n_gpus": ctx.world, "unit": "maps/s", "modes": {}}
    for trunk, cls in (("resnet50", semdiff_b200.CLIP_lpips_Unet_clsbckbn), ("resnet50_clip.openai", semdiff_b200.CLIP_lpips_Unet)):
        for mode in args.modes.split(","):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                model = cls(trunk, str(ctx.dev), precision=mode).eval()

            def step():
                with torch.no_grad():
                    return model(gt, sr)

            for _ in range(3):
                m = step()
            steps = max(3, args.steps // 4)
            ms = ctx.timed(step, steps) / steps
            result["modes"][f"{trunk}/{mode}"] = {"ms_per_step": ms, "maps_per_s": n * ctx.world / (ms / 1e3), "launches_per_step": model.plan().last_launches(),
                                                   "tflops": (56.0e9 + (16.35e9 if trunk == "resnet50" else 21.47e9)) * n / (ms / 1e3) / 1e12,
                                                   "map_mean": float(m.mean())}
            del model
            torch.cuda.empty_cache()
    if ctx.rank == 0:
        print(json.dumps(result), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pairs224", help="pairs224 | sweep10k | hires1024 | unet, or a comma-separated list (one JSON line each)")
    ap.add_argument("--pairs", type=int, default=256, help="pairs per GPU per step (pairs224); total pairs (sweep10k: 10000, hires1024: 32)")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--modes", default="fp16x3,bf16,fp16", help="sweep10k / hires1024: comma-separated precision modes (the first is the rank reference)")
    ap.add_argument("--oracle", type=int, default=0, help="sweep10k: also score the first K pairs with the CPU oracle; hires1024: pair 0 if > 0")
    ap.add_argument("--trunk", default="resnet50")
    ap.add_argument("--microbatch", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-x3", action="store_true", help="skip the fp16x3 extra key")
    ap.add_argument("--detail", action="store_true", help="pairs224: print every note and sub-measurement instead of the compact line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))
        return
    ctx = Ctx()
    try:
        for wl in args.workload.split(","):
            {"pairs224": run_pairs224, "sweep10k": run_sweep10k, "hires1024": run_hires1024, "unet": run_unet}[wl](args, ctx)
            ctx.torch.cuda.empty_cache()
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
