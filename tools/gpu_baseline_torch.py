#!/usr/bin/env python
"""The 'existing Blackwell kernels' bar: the oracle module (== the reference's forward) run by PyTorch
eager + cuDNN on the same B200, 256 pairs of 224x224, in fp32 (TF32 off / on), bf16 and fp16 channels_last."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.restated import RestatedScorer
from oracle.synth import set_head

def bench(model, gt, sr, reps=5):
    with torch.no_grad():
        for _ in range(2): model(gt, sr)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
        for _ in range(reps): model(gt, sr)
        e1.record(); torch.cuda.synchronize()
    return gt.shape[0] * reps / (e0.elapsed_time(e1) / 1e3)

trunk = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
n = 256
gt = torch.randn(n, 3, 224, 224, device="cuda"); sr = gt + 0.1 * torch.randn_like(gt)
out = {"trunk": trunk, "pairs": n, "gpu": torch.cuda.get_device_name(0), "cudnn": torch.backends.cudnn.version()}
torch.backends.cudnn.benchmark = True
for name, dtype, tf32 in (("fp32_tf32_off", torch.float32, False), ("fp32_tf32_on(default)", torch.float32, True),
                          ("bf16_channels_last", torch.bfloat16, True), ("fp16_channels_last", torch.float16, True)):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    m = set_head(RestatedScorer(trunk, 3, seed=0), "abs").cuda().to(dtype)
    a, b = gt.to(dtype), sr.to(dtype)
    if dtype != torch.float32:
        m = m.to(memory_format=torch.channels_last); a = a.contiguous(memory_format=torch.channels_last); b = b.contiguous(memory_format=torch.channels_last)
    out[name + "_pairs_per_s"] = bench(m, a, b)
print(json.dumps(out))
