"""CPU tests of bench.py's host logic: the reference arm prints the contract's JSON line (it is the oracle timed on the host
cores - the one place bench.py may execute oracle/), the sweep generator is shard-invariant, rank statistics, and the
kernel-source stamp that gates `roofline.traffic`."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_other_ranks_do_no_work(monkeypatch):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_sweep_pairs_are_shard_invariant():
    """Any shard regenerates exactly its own pairs (per-block seeds): the basis of the 1-vs-N bit-equality of configs[3]."""
    full = bench.sweep_pairs_device(0, 130, "cpu", src=32, size=16)
    for lo, hi in ((0, 50), (37, 101), (100, 130), (49, 51)):
        part = bench.sweep_pairs_device(lo, hi, "cpu", src=32, size=16)
        assert torch.equal(part[0], full[0][lo:hi]) and torch.equal(part[1], full[1][lo:hi])
    assert full[0].shape == (130, 3, 16, 16) and not torch.equal(full[0], full[1])


def test_rank_statistics():
    a = torch.tensor([0.1, 0.5, 0.3, 0.9, 0.7])
    assert abs(bench.spearman(a, a * 2 + 1) - 1.0) < 1e-12 and bench.inversions(a, a) == 0
    b = a.clone()
    b[1], b[2] = a[2], a[1]           # swap two neighbours in the order
    assert bench.inversions(a, b) == 1 and bench.spearman(a, b) < 1.0
    assert abs(bench.spearman(a, -a) + 1.0) < 1e-12


def test_kernel_source_stamp():
    d = bench.kernel_sources_digest()
    assert len(d) == 64 and d == bench.kernel_sources_digest()
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
        tj = json.load(f)
    assert set(tj) >= {"kernel_sources_sha256", "conv_tc_dram_bytes_per_launch", "source"}
    assert os.path.isfile(os.path.join(ROOT, tj["source"]))


def test_compact_line_keeps_the_contract_keys_and_stays_short():
    """bench.py prints the compact form by default: every key of the contract, numbers only, short enough that a log tail holds
    the whole line (the detailed records under profiles/ come from --detail)."""
    with open(os.path.join(ROOT, "profiles", "r2_bench_1gpu.json")) as f:
        detail = json.load(f)
    c = bench.compact_line(detail)
    line = json.dumps(c)
    assert len(line) < 3000 and "\n" not in line
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in c, k
    assert set(c["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(c["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and c["e2e"]["h2d_bytes_per_step"] > 0
    assert set(c["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert "workload" in c["config"] and c["value"] == detail["value"] and c["fp16x3"]["value"] == detail["fp16x3"]["value"]
