#!/usr/bin/env python
"""PCIe / end-to-end probe: pinned H2D bandwidth and score_host() time for several chunk sizes."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SEMDIFF_RANDOM_INIT", "1")
import torch
import semdiff_b200

def timed(fn, reps=5):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps

for mb in (16, 64, 154, 308):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True); d = torch.empty_like(h, device="cuda")
    ms = timed(lambda: d.copy_(h, non_blocking=True)); print(f"H2D pinned {mb} MB: {mb / 1024 / (ms / 1e3):.1f} GB/s")
    ms = timed(lambda: h.copy_(d, non_blocking=True)); print(f"D2H pinned {mb} MB: {mb / 1024 / (ms / 1e3):.1f} GB/s")
with contextlib.redirect_stdout(io.StringIO()):
    model = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb("resnet50", 3, "cuda").eval()
n = 256
gt = torch.randn(n, 3, 224, 224).pin_memory(); sr = (gt + 0.1 * torch.randn(n, 3, 224, 224)).pin_memory()
out = torch.empty(n).pin_memory()
gd, sd = gt.cuda(), sr.cuda()
with torch.no_grad():
    print(f"device-resident forward: {timed(lambda: model(gd, sd)):.2f} ms")
for chunk in (32, 64, 128, 256):
    ms = timed(lambda: model.score_host(gt, sr, out, chunk_pairs=chunk))
    def two():
        _, e = model.score_host(gt, sr, out, chunk_pairs=chunk, wait=False); _, e2 = model.score_host(gt, sr, out, chunk_pairs=chunk, wait=False); e.synchronize(); e2.synchronize()
    print(f"  two calls in flight: {timed(two) / 2:.2f} ms per call")
    print(f"score_host chunk={chunk}: {ms:.2f} ms -> {n / ms * 1e3:.0f} pairs/s")
h16g, h16s = gt.to(torch.bfloat16).pin_memory(), sr.to(torch.bfloat16).pin_memory()
print("bf16 host tensors:", h16g.dtype, h16g.is_pinned())
