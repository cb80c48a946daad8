"""CPU tests that pin the oracle: (1) oracle/restated.py == the UNMODIFIED reference file run through the timm shim
(only where /root/reference exists, i.e. the build container); (2) oracle/restated.py reproduces the committed golden
vectors that the reference itself produced (tests/golden/scorer_goldens.json, generator: oracle/make_goldens.py)."""
import json
import os

import pytest
import torch

from oracle import reference_loader as rl
from oracle.restated import RestatedScorer, tap_names
from oracle.synth import make_pairs, set_head

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "scorer_goldens.json")


@pytest.mark.skipif(not rl.available(), reason="/root/reference not present on this host")
@pytest.mark.parametrize("trunk,depth", [("resnet50", 3), ("resnet50", 1), ("resnet50_clip.openai", 3), ("resnet50_clip.openai", 0)])
def test_restated_equals_reference_file(trunk, depth):
    ref = set_head(rl.build_reference_scorer(trunk, depth, seed=0), "abs")
    mine = set_head(RestatedScorer(trunk, depth, seed=0), "abs")
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.equal(v, mine.state_dict()[k]), k
    assert ref.wanted_layers == mine.wanted_layers == tap_names(trunk, depth)
    gt, sr = make_pairs(3, seed=21)
    with torch.no_grad():
        a, b = ref(gt, sr), mine(gt, sr)
    assert torch.equal(a, b)   # same torch ops in the same order on the same host: bit-identical


def test_restated_reproduces_goldens():
    with open(GOLDEN) as f:
        records = json.load(f)["records"]
    assert len(records) >= 8
    for rec in records:
        if rec["n_pairs"] > 5 and rec["head"] == "signed":
            continue  # keep the CPU suite short; the abs twin covers the same trunk pass
        model = set_head(RestatedScorer(rec["trunk"], rec["depth"], seed=rec["weight_seed"]), rec["head"])
        gt, sr = make_pairs(rec["n_pairs"], seed=rec["input_seed"])
        got = model(gt, sr)
        pre = model(gt, sr, pre_relu=True) if rec["n_pairs"] <= 5 else None
        assert len(model.state_dict()) == rec["state_dict_keys"]
        ref = torch.tensor(rec["scores"])
        scale = torch.tensor(rec["pre_relu"]).abs().clamp_min(1e-3)
        assert ((got - ref).abs() / scale).max().item() < 5e-5, rec
        if pre is not None:
            assert torch.allclose(pre, torch.tensor(rec["pre_relu"]), rtol=5e-5, atol=1e-5)


def test_tap_shapes_and_depths():
    model = RestatedScorer("resnet50", 3, seed=0)
    feats = model.features(torch.randn(1, 3, 224, 224))
    assert [tuple(f.shape) for f in feats] == [(1, 256, 56, 56), (1, 512, 28, 28), (1, 1024, 14, 14), (1, 2048, 7, 7)]
    for depth in range(4):
        assert len(tap_names("resnet50", depth)) == depth + 1 == len(tap_names("resnet50_clip.openai", depth))


@pytest.mark.skipif(not rl.available(), reason="/root/reference not present on this host")
def test_restated_wperlay_equals_reference_file():
    ref = set_head(rl.build_reference_scorer("resnet50_clip.openai", 5, seed=0, variant="wperlay"), "abs")
    mine = set_head(RestatedScorer("resnet50_clip.openai", 5, seed=0, variant="wperlay"), "abs")
    assert ref.wanted_layers == mine.wanted_layers and len(mine.w_layers) == 6
    assert [m.in_channels for m in ref.w_layers] == [m.in_channels for m in mine.w_layers] == [1024] * 3 + [2048] * 3
    gt, sr = make_pairs(2, seed=5)
    with torch.no_grad():
        assert torch.equal(ref(gt, sr), mine(gt, sr))


def _block2_outputs(model, x):
    """Outputs of layer{1..4}[2] (the modules the reference hooks, :701) of a torchvision-style ResNet."""
    outs, hooks = [], []
    for li in range(1, 5):
        hooks.append(getattr(model, f"layer{li}")[2].register_forward_hook(lambda m, i, o: outs.append(o)))
    with torch.no_grad():
        model(x)
    for h in hooks:
        h.remove()
    return outs


def test_imagenet_oracle_trunk_is_torchvision_resnet50():
    """Pins oracle/trunks.py::resnet50 (a restatement with timm's module names): its weights loaded into
    torchvision.models.resnet50 - and, where /root/reference exists, into the torchvision copy the reference vendors
    (additional_approaches/src/transalnet/utils/resnet.py:326-334) - give bit-identical tap activations."""
    import importlib.util

    import torchvision

    from oracle.trunks import build_trunk
    mine = build_trunk("resnet50", seed=0, calibrate_bn=True).eval()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    feats = RestatedScorer("resnet50", 3, seed=0)
    feats.clip.load_state_dict(mine.state_dict())
    want = feats.features(x)
    candidates = [("torchvision", torchvision.models.resnet50(weights=None))]
    vendored = "/root/reference/additional_approaches/src/transalnet/utils/resnet.py"
    if os.path.isfile(vendored):
        spec = importlib.util.spec_from_file_location("_ref_vendored_resnet", vendored)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        candidates.append(("reference's vendored resnet.py", mod.resnet50(pretrained=False)))
    for name, tv in candidates:
        missing = tv.load_state_dict(mine.state_dict(), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys, name
        got = _block2_outputs(tv.eval(), x)
        assert len(got) == 4
        for a, b in zip(got, want):
            assert torch.equal(a, b), name


# ---- local maps: /root/reference/models/local_eval_models.py:7-339 ----
UNET_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "unet_goldens.json")


@pytest.mark.skipif(not os.path.isfile(rl.REFERENCE_LOCAL_FILE), reason="/root/reference not present on this host")
@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
def test_restated_unet_equals_reference_file(trunk):
    from oracle.restated import RestatedUnet, calibrate_unet_decoder, unet_tap_names
    ref = calibrate_unet_decoder(rl.build_reference_unet(trunk, seed=0))
    mine = calibrate_unet_decoder(RestatedUnet(trunk, seed=0))
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.equal(v, mine.state_dict()[k]), k
    assert ref.wanted_layers == mine.wanted_layers == unet_tap_names(trunk)
    gt, sr = make_pairs(1, seed=9)
    with torch.no_grad():
        a, b = ref(gt, sr), mine(gt, sr)
    assert a.shape == (1, 1, 224, 224) and torch.equal(a, b)


def test_restated_unet_reproduces_goldens():
    from oracle.restated import RestatedUnet, calibrate_unet_decoder
    with open(UNET_GOLDEN) as f:
        records = json.load(f)["records"]
    assert {r["trunk"] for r in records} == {"resnet50", "resnet50_clip.openai"}
    rec = records[0]   # one trunk keeps the CPU suite short; the GPU suite checks both
    model = calibrate_unet_decoder(RestatedUnet(rec["trunk"], seed=rec["weight_seed"]))
    gt, sr = make_pairs(rec["n_pairs"], seed=rec["input_seed"])
    m = model(gt, sr)
    assert list(m.shape) == rec["shape"] and len(model.state_dict()) == rec["state_dict_keys"]
    got = m[:, 0, 8::16, 8::16].reshape(rec["n_pairs"], -1)
    assert torch.allclose(got, torch.tensor(rec["map"]), rtol=0, atol=2e-5)
    assert float(got.std()) > 0.01   # a real map, not a saturated constant
