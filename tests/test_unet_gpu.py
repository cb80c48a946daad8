"""GPU parity of the local-map path: decoder helper kernels against the torch ops the reference calls
(/root/reference/models/local_eval_models.py:84, :115, :121), and the drop-in U-Net modules against the oracle
(oracle/restated.py RestatedUnet == the reference file executed verbatim, tests/test_oracle.py) and the committed goldens."""
import json
import os

import pytest
import torch

import semdiff_b200
from helpers import DT, lib, sp, split_store, split_value
from oracle.restated import RestatedUnet, calibrate_unet_decoder
from oracle.synth import make_pairs
from semdiff_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "unet_goldens.json")
CLS = {"resnet50": semdiff_b200.CLIP_lpips_Unet_clsbckbn, "resnet50_clip.openai": semdiff_b200.CLIP_lpips_Unet}
_cache = {}


def store(x, precision):
    return split_store(x, DT[precision]) if precision in _lib.SPLIT else x.to(DT[precision]).contiguous()


def value(x, precision):
    return split_value(x) if precision in _lib.SPLIT else x.double()


def op(what, a, b, out, n, H, W, c, c2, precision):
    _lib.check(lib().semdiff_decoder_op(what, a.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), n, H, W, c, c2,
                                        _lib.PRECISIONS[precision], sp()), "decoder_op")
    torch.cuda.synchronize()


@pytest.mark.parametrize("precision", ["bf16", "fp32", "fp16x3"])
def test_decoder_ops(precision):
    g = torch.Generator(device=DEV).manual_seed(4)
    mult = 2 if precision in _lib.SPLIT else 1
    tol = {"bf16": 8e-3, "fp32": 1e-6, "fp16x3": 1e-6}[precision]
    n, H, W, C, C2 = 3, 7, 5, 128, 64
    # squared difference of the GT / SR halves (:115), incl. differences far below the 16-bit resolution of the operands
    a = torch.randn(n, H, W, C, device=DEV, generator=g)
    b = a + torch.tensor([1e-3, 0.1, 1.0], device=DEV)[:, None, None, None] * torch.randn(n, H, W, C, device=DEV, generator=g)
    x = store(torch.cat([a, b]), precision)
    out = torch.empty(n, H, W, C * mult, dtype=DT[precision], device=DEV)
    op(0, x, None, out, n, H, W, C, 0, precision)
    xv = value(x, precision)
    ref = (xv[:n] - xv[n:]) ** 2
    # fp16 halves: full 22-bit resolution for d^2 in [6.1e-5, 6.5e4]; below, the fp16 subnormal spacing (6e-8) is the floor
    assert ((value(out, precision) - ref).abs() <= tol * ref.abs() + (6e-8 if precision == "fp16x3" else 1e-30)).all()
    # channel concat (:121)
    y = store(torch.randn(n, H, W, C2, device=DEV, generator=g), precision)
    cat = torch.empty(n, H, W, (C + C2) * mult, dtype=DT[precision], device=DEV)
    op(1, out, y, cat, n, H, W, C, C2, precision)
    assert torch.equal(value(cat, precision), torch.cat([value(out, precision), value(y, precision)], dim=-1))
    # nn.UpsamplingBilinear2d(scale_factor=2) (:84)
    up = torch.empty(n, 2 * H, 2 * W, C2 * mult, dtype=DT[precision], device=DEV)
    op(2, y, None, up, n, H, W, C2, 0, precision)
    ref = torch.nn.UpsamplingBilinear2d(scale_factor=2)(value(y, precision).permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert ((value(up, precision) - ref).abs().max() / ref.abs().max()).item() < tol
    # channel 0 -> upsample -> sigmoid (:123-125)
    m = torch.empty(n, 1, 2 * H, 2 * W, device=DEV)
    op(3, y, None, m, n, H, W, C2, 0, precision)
    ref = torch.sigmoid(torch.nn.UpsamplingBilinear2d(scale_factor=2)(value(y, precision)[..., :1].permute(0, 3, 1, 2)))
    assert (m.double() - ref).abs().max().item() < 1e-6


def oracle_and_module(trunk, precision):
    if trunk not in _cache:
        _cache[trunk] = calibrate_unet_decoder(RestatedUnet(trunk, seed=0))
    oracle = _cache[trunk]
    model = CLS[trunk](clip_name=trunk, device="cuda", precision=precision)
    res = model.load_state_dict(oracle.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return oracle, model.eval()


def logit(p):
    p = p.double().clamp(1e-12, 1 - 1e-12)
    return torch.log(p / (1 - p))


@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
@pytest.mark.parametrize("precision,tol", [("fp16x3", 2e-5), ("fp32", 2e-5), ("bf16", 0.5)])
def test_unet_matches_oracle(trunk, precision, tol):
    """Map of the B200 path vs the oracle on the same seeded weights and pairs (one SR ~ GT pair, one far pair).  Compared
    in logit space (the decoder's output before the sigmoid, relative to its range) so that saturated pixels cannot hide
    an error; the 16-bit mode is only bounded."""
    oracle, model = oracle_and_module(trunk, precision)
    gt, sr = make_pairs(2, seed=5)
    ref = oracle(gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    assert got.shape == ref.shape == (2, 1, 224, 224) and got.dtype == torch.float32
    scale = logit(ref).abs().max()
    err = ((logit(got) - logit(ref)).abs().max() / scale).item()
    print(f"[unet] {trunk} {precision}: max logit err / max |logit| = {err:.3g}; max map err {(got - ref).abs().max().item():.3g}; "
          f"launches {model.plan().last_launches()}")
    assert err < tol
    if precision != "bf16":
        assert (got - ref).abs().max().item() < 1e-5


def test_unet_goldens():
    """The reference's own outputs (tests/golden/unet_goldens.json, generator `python -m oracle.make_goldens unet`)."""
    with open(GOLDEN) as f:
        records = json.load(f)["records"]
    for rec in records:
        _, model = oracle_and_module(rec["trunk"], "fp16x3")
        gt, sr = make_pairs(rec["n_pairs"], seed=rec["input_seed"])
        with torch.no_grad():
            m = model(gt.cuda(), sr.cuda()).cpu()
        got = m[:, 0, 8::16, 8::16].reshape(rec["n_pairs"], -1)
        err = (got - torch.tensor(rec["map"])).abs().max().item()
        print(f"[unet golden] {rec['trunk']}: {err:.3g}")
        assert err < 2e-5


def test_unet_module_contract(tmp_path):
    oracle, model = oracle_and_module("resnet50", "fp16x3")
    assert model.wanted_layers == ["conv1", "layer1.2.act3", "layer2.2.act3", "layer3.2.act3", "layer4.2.act3"]
    assert list(model.state_dict().keys()) == list(oracle.state_dict().keys())
    p = str(tmp_path / "dec.pt")
    model.save_model(p)
    assert set(torch.load(p, weights_only=True).keys()) == set(oracle.decoder.state_dict().keys())
    gt, sr = make_pairs(3, seed=2)
    gt, sr = gt.cuda(), sr.cuda()
    with torch.no_grad():
        m0 = model(gt, sr)
        model.microbatch = 2                      # ragged micro-batches, same maps
        assert torch.equal(model(gt, sr), m0)
        assert torch.equal(model(sr, gt), m0)     # (a - b)^2: symmetric
        list(model.decoder[0].children())[3].bias.add_(0.5)
        model.refresh()
        assert not torch.equal(model(gt, sr), m0)
        model.load_model(p)
        assert torch.equal(model(gt, sr), m0)
        assert model(gt[:0], sr[:0]).shape == (0, 1, 224, 224)
    with pytest.raises(NotImplementedError, match="inference-only"):
        model(gt, sr)
    with pytest.raises(NotImplementedError, match="lora_rank"):
        CLS["resnet50"]("resnet50", "cuda", lora_rank=4)
    with pytest.raises(ValueError, match="multiples of 32"):
        with torch.no_grad():
            model(gt[:, :, :200], sr[:, :, :200])
