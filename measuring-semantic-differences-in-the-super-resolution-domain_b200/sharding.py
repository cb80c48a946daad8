"""Multi-GPU sweep: GT/SR pairs are independent, so they shard across ranks in contiguous blocks with no data-path
collective; the only exchange is ONE all-gather of the fp32 scores at the end.  The reference has no
multi-GPU scorer path (its only multi-GPU code launches independent processes,
/root/reference/CLU_training_sweep_example.py:184-197).  Works with any torch.distributed backend: NCCL over
NVLink/NVSwitch on the B200 box, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_size(n_pairs: int, world: int) -> int:
    return (n_pairs + world - 1) // world


def shard_range(n_pairs: int, world: int, rank: int) -> tuple[int, int]:
    """Rank r scores pairs [r*ceil(P/W), min(P, (r+1)*ceil(P/W)))."""
    per = shard_size(n_pairs, world)
    lo = min(n_pairs, rank * per)
    return lo, min(n_pairs, lo + per)


def gather_scores(local: torch.Tensor, n_pairs: int, group=None) -> torch.Tensor:
    """One all_gather_into_tensor of equal-sized (tail-padded) fp32 blocks -> scores of all `n_pairs` pairs, in pair
    order, on every rank."""
    world = dist.get_world_size(group)
    per = shard_size(n_pairs, world)
    block = torch.zeros(per, dtype=torch.float32, device=local.device)
    block[: local.numel()] = local
    out = torch.empty(world * per, dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(out, block, group=group)
    return out[:n_pairs]


def score_sharded(score_fn, n_pairs: int, load_pairs, group=None) -> torch.Tensor:
    """Score `n_pairs` pairs across the ranks of `group`.

    load_pairs(lo, hi) -> (gt, sr) tensors for pairs [lo, hi) on this rank's device; score_fn(gt, sr) -> [hi-lo]."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_pairs, world, rank)
    if hi > lo:
        gt, sr = load_pairs(lo, hi)
        local = score_fn(gt, sr).float()
    else:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        local = torch.empty(0, dtype=torch.float32, device=dev)
    return gather_scores(local, n_pairs, group)


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (one process per GPU), so that the pinned host
    buffers it allocates afterwards are first-touched on that node and every rank's H2D stream reads node-local DRAM
    instead of all eight PCIe streams pulling from wherever the allocator happened to run.  Best effort: returns what it
    did ({"node": n, "cpus": k}) or why not ({"skipped": reason}); never raises."""
    import os

    try:
        import pynvml as nv

        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:        # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return {"skipped": "single NUMA node (numa_node = -1)"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read()) & os.sched_getaffinity(0)
        if not cpus:
            return {"skipped": f"no allowed CPU on node {node}"}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001 - placement is an optimisation, not a requirement
        return {"skipped": f"{type(e).__name__}: {e}"}
