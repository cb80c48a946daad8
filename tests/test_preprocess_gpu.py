"""GPU parity of the on-device `processor` (csrc/preprocess.cu) against the oracle (== Pillow + torchvision, see
test_preprocess_cpu.py): bit-exact fp32 tensors, then identical scores through score_uint8."""
import numpy as np
import pytest
import torch

import semdiff_b200
from oracle.pil_resize import eval_transform
from oracle.restated import RestatedScorer
from oracle.synth import set_head

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("trunk,cls", [("resnet50", "CLIP_lpips_stages_cnn_clsbckb"), ("resnet50_clip.openai", "CLIP_lpips_stages_cnn")])
@pytest.mark.parametrize("hw", [(512, 512), (480, 640), (300, 200), (100, 130)])
def test_gpu_processor_bit_exact(trunk, cls, hw):
    model = getattr(semdiff_b200, cls)(trunk, 0, "cuda")
    gp = model.gpu_processor
    imgs = (np.random.default_rng(hw[0]).random((3, *hw, 3)) * 255).astype(np.uint8)
    got = gp(torch.from_numpy(imgs).cuda()).cpu().numpy()
    mean, std = [float(x) for x in gp.mean], [float(x) for x in gp.std]
    for i in range(3):
        ref = eval_transform(imgs[i], gp.resize_to, gp.size, mean, std)
        assert np.array_equal(got[i], ref), np.abs(got[i] - ref).max()


def test_score_uint8_equals_cpu_processor_then_forward():
    oracle = set_head(RestatedScorer("resnet50", 3, seed=0), "abs")
    model = semdiff_b200.CLIP_lpips_stages_cnn_clsbckb("resnet50", 3, "cuda", precision="fp32")
    model.load_state_dict(oracle.state_dict())
    rng = np.random.default_rng(5)
    hq = (rng.random((2, 512, 512, 3)) * 255).astype(np.uint8)
    sr = np.clip(hq.astype(np.int32) + rng.integers(-25, 25, hq.shape), 0, 255).astype(np.uint8)
    gp = model.gpu_processor
    mean, std = [float(x) for x in gp.mean], [float(x) for x in gp.std]
    a = torch.from_numpy(np.stack([eval_transform(x, gp.resize_to, gp.size, mean, std) for x in sr]))
    b = torch.from_numpy(np.stack([eval_transform(x, gp.resize_to, gp.size, mean, std) for x in hq]))
    ref = oracle(a, b)
    with torch.no_grad():
        got = model.score_uint8(torch.from_numpy(sr).cuda(), torch.from_numpy(hq).cuda()).cpu()
    assert ((got - ref).abs() / ref.abs().clamp_min(1e-3)).max().item() < 1e-5
