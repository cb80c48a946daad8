"""End-to-end GPU parity: the drop-in module (C-ABI semdiff_score underneath) against the oracle
(oracle/restated.py == the unmodified reference file, see test_oracle.py) on the same seeded inputs and weights,
and against the committed golden vectors produced by the reference itself (tests/golden/).

Tolerances (BASELINE.json north_star): 1e-5 relative for the fp32 mode AND for the split-precision tensor-core mode
(fp16x3, the module's default: well inside the 1e-3 the north_star asks of the tensor-core trunk); the plain 16-bit modes
are reported and bounded by what their storage type allows on random-init weights (1e-3 is not reachable
with bf16 activations - rounding every inter-layer tensor to 8 bits destroys (A-B)^2 on SR ~ GT pairs)."""
import json
import os

import pytest
import torch

import semdiff_b200
from oracle.restated import RestatedScorer
from oracle.synth import make_pairs, set_head

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "scorer_goldens.json")
CLS = {"resnet50": semdiff_b200.CLIP_lpips_stages_cnn_clsbckb, "resnet50_clip.openai": semdiff_b200.CLIP_lpips_stages_cnn}
_cache = {}


def oracle_and_module(trunk, depth, precision, head="abs"):
    key = (trunk, depth, head)
    if key not in _cache:
        _cache[key] = set_head(RestatedScorer(trunk, depth, seed=0), head)
    oracle = _cache[key]
    model = CLS[trunk](clip_name=trunk, depth=depth, device="cuda", precision=precision)
    missing = model.load_state_dict(oracle.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return oracle, model.eval()


def rel_err(got, ref):
    return ((got - ref).abs() / ref.abs().clamp_min(1e-3)).max().item()


@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
def test_fp32_mode_matches_oracle(trunk):
    oracle, model = oracle_and_module(trunk, 3, "fp32")
    gt, sr = make_pairs(4, seed=11)
    ref = oracle(gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    assert got.shape == ref.shape and got.dtype == torch.float32
    e = rel_err(got, ref)
    print(f"[parity] {trunk} fp32 max rel err {e:.3g}")
    assert e < 1e-5, (got, ref)


def oracle_fp64(oracle, gt, sr, chunk=16):
    """The oracle's arithmetic in fp64 (the yardstick both fp32-level implementations are measured against)."""
    m = RestatedScorer(oracle.trunk_name, oracle.depth, seed=0)
    m.load_state_dict(oracle.state_dict())
    m = m.double()
    with torch.no_grad():
        return torch.cat([m(gt[i:i + chunk].double(), sr[i:i + chunk].double()) for i in range(0, gt.shape[0], chunk)])


@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
def test_fp16x3_mode_matches_oracle(trunk):
    """The split-precision tensor-core mode (three tcgen05 products per K block on hi + lo fp16 pairs) against the
    oracle: north_star tolerance 1e-3 for the tensor-core trunk, 1e-5 for fp32-level parity."""
    oracle, model = oracle_and_module(trunk, 3, "fp16x3")
    gt, sr = make_pairs(8, seed=0)
    ref = oracle(gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    e = rel_err(got, ref)
    r64 = oracle_fp64(oracle, gt, sr)
    e64, o64 = rel_err(got.double(), r64), rel_err(ref.double(), r64)
    print(f"[parity] {trunk} fp16x3 max rel err vs oracle fp32 {e:.3g}; vs fp64: ours {e64:.3g}, oracle fp32 {o64:.3g}; "
          f"launches {model.plan().last_launches()}")
    assert e < 1e-5, (got, ref)


def test_bf16x3_mode_within_north_star_tolerance():
    """The same split kernels with bf16 halves (16 significant bits, fp32 exponent range: no saturation risk on trunks whose
    activations exceed fp16's 65504): inside the 1e-3 the north_star asks of the tensor-core trunk, incl. SR ~ GT pairs."""
    oracle, model = oracle_and_module("resnet50", 3, "bf16x3")
    gt, sr = make_pairs(8, seed=0)
    gt2, sr2 = make_pairs(8, seed=43, sigma_lo=0.02, sigma_hi=0.03)
    gt, sr = torch.cat([gt, gt2]), torch.cat([sr, sr2])
    ref = oracle(gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    e = rel_err(got, ref)
    print(f"[parity] resnet50 bf16x3 max rel err vs oracle fp32 {e:.3g} (low-sigma half {rel_err(got[8:], ref[8:]):.3g})")
    assert e < 1e-3, (got, ref)


def test_fp16x3_low_sigma_and_sweep_distribution():
    """>= 128 pairs of the sweep distribution (sigma log-uniform in [0.02, 2]) plus 32 pairs forced into the SR ~ GT corner
    (sigma in [0.02, 0.03]) where 16-bit trunks lose the difference in their rounding: fp16x3 must hold the north_star
    tolerance on every pair, and the rank order of the oracle."""
    from scipy.stats import spearmanr
    oracle, model = oracle_and_module("resnet50", 3, "fp16x3")
    gt, sr = make_pairs(128, seed=41)
    gt2, sr2 = make_pairs(32, seed=43, sigma_lo=0.02, sigma_hi=0.03)
    gt, sr = torch.cat([gt, gt2]), torch.cat([sr, sr2])
    ref = torch.cat([oracle(gt[i:i + 16], sr[i:i + 16]) for i in range(0, 160, 16)])
    r64 = oracle_fp64(oracle, gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    e, e_low = rel_err(got, ref), rel_err(got[128:], ref[128:])
    e64, o64 = rel_err(got.double(), r64), rel_err(ref.double(), r64)
    inv = lambda a, b: int((b[a.argsort()][1:] < b[a.argsort()][:-1]).sum())
    print(f"[parity] fp16x3 over 160 pairs: max rel err vs oracle fp32 {e:.3g} (low-sigma corner {e_low:.3g}); vs fp64: ours {e64:.3g}, "
          f"oracle fp32 {o64:.3g}; adjacent inversions vs fp64 order: ours {inv(r64, got.double())}, oracle fp32 {inv(r64, ref.double())}; "
          f"spearman vs oracle {spearmanr(ref.numpy(), got.numpy())[0]:.9f}")
    assert e < 1e-3 and e_low < 1e-3
    assert e64 < 3 * o64 + 1e-5          # as close to the exact result as the reference's own fp32 path
    assert inv(r64, got.double()) <= inv(r64, ref.double()) + 1


def torch_16bit_error(oracle, gt, sr, ref, dtype):
    """The same oracle module run by PyTorch/cuDNN itself in `dtype` (channels_last) on this GPU: the yardstick for
    what 16-bit storage costs on these weights."""
    m = RestatedScorer(oracle.trunk_name, oracle.depth, seed=0)   # fresh module: forward hooks do not deep-copy
    m.load_state_dict(oracle.state_dict())
    m = m.cuda().to(dtype).to(memory_format=torch.channels_last)
    with torch.no_grad():
        got = m(gt.cuda().to(dtype).contiguous(memory_format=torch.channels_last),
                sr.cuda().to(dtype).contiguous(memory_format=torch.channels_last)).float().cpu()
    return rel_err(got, ref)


# Bounds: 16-bit storage of weights AND of every inter-layer activation (SURVEY.md 7.3 measured 4e-3..2e-2 for
# bf16 and ~1e-3..3e-3 for fp16 on calibrated random-init trunks; the north_star's 1e-3 is not reachable with bf16
# weights).  The second assertion is the meaningful one: never worse than PyTorch's own 16-bit path.
@pytest.mark.parametrize("precision,bound", [("bf16", 6e-2), ("fp16", 8e-3)])
@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
def test_16bit_modes_vs_oracle(trunk, precision, bound):
    oracle, model = oracle_and_module(trunk, 3, precision)
    gt, sr = make_pairs(8, seed=0)
    ref = oracle(gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    e = rel_err(got, ref)
    e_torch = torch_16bit_error(oracle, gt, sr, ref, torch.bfloat16 if precision == "bf16" else torch.float16)
    print(f"[parity] {trunk} {precision} max rel err {e:.3g} (torch/cuDNN {precision} on the same GPU: {e_torch:.3g})")
    assert e < bound
    assert e < 1.5 * e_torch + 1e-3


def test_tc_path_equals_simt_path_bf16():
    """Same bf16 operands through the tcgen05 kernels and through the CUDA-core kernels: only fp32 summation order
    differs, so after 50 layers the scores agree far tighter than bf16 resolution."""
    oracle, model = oracle_and_module("resnet50", 3, "bf16")
    gt, sr = make_pairs(4, seed=5)
    with torch.no_grad():
        tc = model(gt.cuda(), sr.cuda()).cpu()
        model.plan().set_conv_impl(semdiff_b200._lib.CONV_SIMT)
        simt = model(gt.cuda(), sr.cuda()).cpu()
        model.plan().set_conv_impl(semdiff_b200._lib.CONV_TC_GATHER)
        gather = model(gt.cuda(), sr.cuda()).cpu()
    print(f"[parity] tc {tc.tolist()} simt {simt.tolist()} gather {gather.tolist()}")
    assert rel_err(tc, simt) < 5e-3
    assert rel_err(gather, simt) < 5e-3


@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_chained_block_boundaries_bit_exact(trunk, precision):
    """The plan fuses conv3 -> next conv1 across the 256-channel block boundaries (csrc/conv_chain.cu); an explicit conv
    impl keeps one launch per conv.  Same rounded tile either way, so scores and launch counts differ by the fusion only."""
    oracle, model = oracle_and_module(trunk, 3, precision)
    gt, sr = make_pairs(6, seed=11)
    with torch.no_grad():
        fused = model(gt.cuda(), sr.cuda()).cpu()
        n_fused = model.plan().last_launches()
        model.plan().set_conv_impl(semdiff_b200._lib.CONV_TC_TMA)
        plain = model(gt.cuda(), sr.cuda()).cpu()
        n_plain = model.plan().last_launches()
    print(f"[chain] {trunk} {precision}: {n_plain} -> {n_fused} launches")
    # 256-channel stage: three block boundaries; 512-channel stage: the two identity-block boundaries; the stem's
    # conv / pool pair (max pool | CLIP: 2x2 average pool)
    assert n_plain - n_fused == 6
    assert torch.equal(fused, plain)


def test_goldens_fp32():
    with open(GOLDEN) as f:
        records = json.load(f)["records"]
    for rec in records:
        oracle, model = oracle_and_module(rec["trunk"], rec["depth"], "fp32", rec["head"])
        gt, sr = make_pairs(rec["n_pairs"], seed=rec["input_seed"])
        with torch.no_grad():
            got = model(gt.cuda(), sr.cuda()).cpu()
        ref = torch.tensor(rec["scores"])
        scale = torch.tensor(rec["pre_relu"]).abs().clamp_min(1e-3)
        err = ((got - ref).abs() / scale).max().item()
        print(f"[golden] {rec['trunk']} depth={rec['depth']} head={rec['head']} err {err:.3g}")
        tol = 1e-5 if rec["head"] == "abs" else 2e-3   # signed default-init heads cancel catastrophically
        assert err < tol, (rec, got)


def test_module_contract(tmp_path):
    oracle, model = oracle_and_module("resnet50", 2, "bf16")
    assert model.wanted_layers == ["layer2.2.act3", "layer3.2.act3", "layer4.2.act3"]
    assert [tuple(m.weight.shape) for m in model.w_layers] == [(1, 512, 1, 1), (1, 1024, 1, 1), (1, 2048, 1, 1)]
    assert set(model.state_dict().keys()) == set(oracle.state_dict().keys())
    p = str(tmp_path / "w.pt")
    model.save_model(p)
    assert set(torch.load(p, weights_only=True).keys()) == {"0.weight", "0.bias", "1.weight", "1.bias", "2.weight", "2.bias"}
    gt, sr = make_pairs(3, seed=2)
    with torch.no_grad():
        s0 = model(gt.cuda(), sr.cuda())
        for m in model.w_layers:
            m.weight.mul_(2.0)
        s1 = model(gt.cuda(), sr.cuda())
        model.load_model(p)
        s2 = model(gt.cuda(), sr.cuda())
    assert not torch.equal(s0, s1) and torch.equal(s0, s2)
    model.train()
    assert not model.clip.training
    with torch.no_grad():
        assert model(gt[:0].cuda(), sr[:0].cuda()).shape == (0,)
    with pytest.raises(NotImplementedError):
        CLS["resnet50"]("resnet50", 1, "cuda", enc_ft=True)


def test_microbatch_and_batch_position_invariance():
    oracle, model = oracle_and_module("resnet50", 3, "bf16")
    gt, sr = make_pairs(6, seed=9)
    gt, sr = gt.cuda(), sr.cuda()
    with torch.no_grad():
        model.microbatch = 6
        full = model(gt, sr)
        model.microbatch = 4          # ragged: 4 + 2
        ragged = model(gt, sr)
        single = torch.cat([model(gt[i:i + 1], sr[i:i + 1]) for i in range(6)])
    assert torch.equal(full, ragged) and torch.equal(full, single)


def test_default_microbatch_512_pairs_same_scores():
    """16-bit modes pass up to 512 pairs (1024 images) per kernel-program pass by default: tensors beyond 2^31 bytes,
    twice the tiles per launch - and a pair's score must not depend on it (one deterministic reduction order per pair)."""
    oracle, model = oracle_and_module("resnet50", 3, "bf16")
    assert model.default_microbatch(224, 224) == 512 and model.default_microbatch(1024, 1024) == 24
    g = torch.Generator(device="cuda").manual_seed(3)
    gt = torch.randn(520, 3, 224, 224, device="cuda", generator=g)
    sr = gt + 0.2 * torch.randn(520, 3, 224, 224, device="cuda", generator=g)
    with torch.no_grad():
        big = model(gt, sr)               # 512 + 8
        model.microbatch = 130
        small = model(gt, sr)
    assert torch.equal(big, small) and bool(torch.isfinite(big).all()) and float(big.std()) > 0


def test_forward_is_symmetric_and_zero_on_identical_pairs():
    """Domain properties of the reference's forward: (a - b)^2 makes it symmetric in its arguments - the
    callers pass (SR, HQ) - and a pair of identical images scores relu(mean of the biases)."""
    oracle, model = oracle_and_module("resnet50", 3, "bf16")
    gt, sr = make_pairs(5, seed=13)
    gt, sr = gt.cuda(), sr.cuda()
    with torch.no_grad():
        ab, ba = model(gt, sr), model(sr, gt)
        same = model(gt, gt)
    assert torch.equal(ab, ba)
    bias = torch.stack([m.bias.detach().reshape(()) for m in model.w_layers]).mean()
    assert torch.allclose(same, torch.relu(bias).expand(5), rtol=0, atol=1e-7)


def test_head_gradients():
    oracle, model = oracle_and_module("resnet50", 1, "fp32")
    gt, sr = make_pairs(3, seed=4)
    target = torch.tensor([1.0, 0.0, 0.5])
    out = model(gt.cuda(), sr.cuda())
    torch.nn.functional.mse_loss(out, target.cuda()).backward()
    ref_model = RestatedScorer("resnet50", 1, seed=0)
    set_head(ref_model, "abs")
    fa, fb = ref_model.features(gt), ref_model.features(sr)
    with torch.enable_grad():
        per = []
        for j, (xa, xb) in enumerate(zip(fa, fb)):
            per.append(ref_model.w_layers[j](((xa - xb) ** 2).detach()).squeeze(1).mean((-1, -2)))
        ref_out = torch.relu(torch.stack(per).mean(0))
        torch.nn.functional.mse_loss(ref_out, target).backward()
    for m, r in zip(model.w_layers, ref_model.w_layers):
        assert torch.allclose(m.weight.grad.cpu(), r.weight.grad, rtol=1e-3, atol=1e-6)
        assert torch.allclose(m.bias.grad.cpu(), r.bias.grad, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("size", [(225, 223), (256, 320)])
def test_other_image_sizes_fp32(size):
    """Fully convolutional like the reference: non-224 sizes; odd sizes take the generic (non space-to-depth) stem."""
    oracle, model = oracle_and_module("resnet50", 3, "fp32")
    g = torch.Generator().manual_seed(3)
    gt = torch.randn(2, 3, *size, generator=g)
    sr = gt + 0.2 * torch.randn(2, 3, *size, generator=g)
    ref = oracle(gt, sr)
    with torch.no_grad():
        got = model(gt.cuda(), sr.cuda()).cpu()
    assert rel_err(got, ref) < 1e-5, (got, ref)
    _, m16 = oracle_and_module("resnet50", 3, "bf16")
    with torch.no_grad():
        got16 = m16(gt.cuda(), sr.cuda()).cpu()
    assert rel_err(got16, ref) < 6e-2


@pytest.mark.parametrize("trunk", ["resnet50", "resnet50_clip.openai"])
def test_head_chunking_is_bit_identical(trunk, monkeypatch):
    """Running pack -> stem -> pool in L2-sized image chunks (ragged last chunk included) must not change a bit."""
    gt, sr = make_pairs(5, seed=13)
    monkeypatch.setenv("SEMDIFF_HEAD_L2_MB", "0")
    _, plain = oracle_and_module(trunk, 3, "bf16")
    with torch.no_grad():
        ref = plain(gt.cuda(), sr.cuda())
    monkeypatch.setenv("SEMDIFF_HEAD_L2_MB", "12")      # ~3 images per chunk -> 10 images = 3 + 3 + 3 + 1
    _, chunked = oracle_and_module(trunk, 3, "bf16")
    with torch.no_grad():
        got = chunked(gt.cuda(), sr.cuda())
    assert chunked.plan().last_launches() > plain.plan().last_launches()
    assert torch.equal(ref, got)


def test_rank_order_16bit_vs_fp32_mode():
    """BASELINE.json asks for identical Spearman order; with continuous scores whose neighbours differ by less than the
    16-bit error that is not attainable, so the test reports rho and bounds it."""
    from scipy.stats import spearmanr
    oracle, m32 = oracle_and_module("resnet50", 3, "fp32")
    _, m16 = oracle_and_module("resnet50", 3, "bf16")
    _, mh = oracle_and_module("resnet50", 3, "fp16")
    gt, sr = make_pairs(96, seed=31)
    with torch.no_grad():
        s32 = m32(gt.cuda(), sr.cuda()).cpu()
        s16 = m16(gt.cuda(), sr.cuda()).cpu()
        sh = mh(gt.cuda(), sr.cuda()).cpu()
    ref = torch.cat([oracle(gt[i:i + 16], sr[i:i + 16]) for i in range(0, 96, 16)])
    rho32, rho16, rhoh = (spearmanr(ref.numpy(), s.numpy())[0] for s in (s32, s16, sh))
    print(f"[rank] spearman vs oracle over 96 pairs: fp32 {rho32:.6f} fp16 {rhoh:.6f} bf16 {rho16:.6f}")
    assert rho32 > 0.99999 and rhoh > 0.9995 and rho16 > 0.995


@pytest.mark.parametrize("depth", [11, 4])
def test_wperlay_variant(depth):
    """CLIP_lpips_wperlay_cnn (reference :815-914): up to 12 taps, one per block output."""
    oracle = set_head(RestatedScorer("resnet50_clip.openai", depth, seed=0, variant="wperlay"), "abs")
    gt, sr = make_pairs(3, seed=17)
    ref = oracle(gt, sr)
    for precision, tol in (("fp32", 1e-5), ("fp16x3", 1e-5), ("bf16", 6e-2)):
        model = semdiff_b200.CLIP_lpips_wperlay_cnn("resnet50_clip.openai", depth, "cuda", precision=precision)
        assert model.wanted_layers == oracle.wanted_layers
        model.load_state_dict(oracle.state_dict(), strict=True)
        with torch.no_grad():
            got = model(gt.cuda(), sr.cuda()).cpu()
        assert rel_err(got, ref) < tol, (precision, got, ref)


def test_high_resolution_pair_1024():
    """BASELINE.json configs[4] geometry: 1024x1024 pairs (taps 256^2 .. 32^2); bf16 tensor-core path vs the fp32 path."""
    _, m32 = oracle_and_module("resnet50", 3, "fp32")
    _, m16 = oracle_and_module("resnet50", 3, "bf16")
    g = torch.Generator().manual_seed(8)
    gt = torch.randn(1, 3, 1024, 1024, generator=g)
    sr = (gt + 0.3 * torch.randn(1, 3, 1024, 1024, generator=g)) / (1 + 0.09) ** 0.5
    with torch.no_grad():
        s32 = m32(gt.cuda(), sr.cuda()).cpu()
        s16 = m16(gt.cuda(), sr.cuda()).cpu()
    assert m16.default_microbatch(1024, 1024) == 24 and m32.default_microbatch(1024, 1024) == 12
    assert torch.isfinite(s32).all() and rel_err(s16, s32) < 6e-2, (s16, s32)


def test_high_resolution_pair_1024_vs_oracle():
    """BASELINE.json configs[4] geometry against the CPU oracle itself (the reference is fully convolutional, :341-397): one
    1024x1024 pair; fp32 mode and the split-precision tensor-core mode within 1e-5, the 16-bit modes bounded."""
    oracle, m32 = oracle_and_module("resnet50", 3, "fp32")
    g = torch.Generator().manual_seed(8)
    gt = torch.randn(1, 3, 1024, 1024, generator=g)
    sr = (gt + 0.05 * torch.randn(1, 3, 1024, 1024, generator=g)) / (1 + 0.0025) ** 0.5
    ref = oracle(gt, sr)
    got = {}
    for precision in ("fp32", "fp16x3", "fp16", "bf16"):
        _, m = oracle_and_module("resnet50", 3, precision)
        with torch.no_grad():
            got[precision] = m(gt.cuda(), sr.cuda()).cpu()
    errs = {k: rel_err(v, ref) for k, v in got.items()}
    print(f"[parity] 1024x1024 pair vs oracle fp32 ({ref.item():.6g}): {errs}")
    assert errs["fp32"] < 1e-5 and errs["fp16x3"] < 1e-5 and errs["fp16"] < 2e-2 and errs["bf16"] < 0.25


def test_goldens_fp16x3():
    """The committed outputs of the reference itself (tests/golden/, generator oracle/make_goldens.py), split-precision mode."""
    with open(GOLDEN) as f:
        records = json.load(f)["records"]
    for rec in records:
        oracle, model = oracle_and_module(rec["trunk"], rec["depth"], "fp16x3", rec["head"])
        gt, sr = make_pairs(rec["n_pairs"], seed=rec["input_seed"])
        with torch.no_grad():
            got = model(gt.cuda(), sr.cuda()).cpu()
        ref = torch.tensor(rec["scores"])
        scale = torch.tensor(rec["pre_relu"]).abs().clamp_min(1e-3)
        err = ((got - ref).abs() / scale).max().item()
        print(f"[golden] fp16x3 {rec['trunk']} depth={rec['depth']} head={rec['head']} err {err:.3g}")
        tol = 1e-5 if rec["head"] == "abs" else 2e-3   # signed default-init heads cancel catastrophically
        assert err < tol, (rec, got)


def test_fp16x3_invariances():
    """Split-precision mode: a pair's score does not depend on the micro-batch, its batch position or the argument order,
    identical images score relu(mean bias), and the head gradients match the fp32 mode's."""
    oracle, model = oracle_and_module("resnet50", 2, "fp16x3")
    gt, sr = make_pairs(6, seed=9)
    gt, sr = gt.cuda(), sr.cuda()
    with torch.no_grad():
        model.microbatch = 6
        full = model(gt, sr)
        model.microbatch = 4          # ragged: 4 + 2
        ragged = model(gt, sr)
        single = torch.cat([model(gt[i:i + 1], sr[i:i + 1]) for i in range(6)])
        swapped = model(sr, gt)
        same = model(gt, gt)
    assert torch.equal(full, ragged) and torch.equal(full, single) and torch.equal(full, swapped)
    bias = torch.stack([m.bias.detach().reshape(()) for m in model.w_layers]).mean()
    assert torch.allclose(same, torch.relu(bias).expand(6), rtol=0, atol=1e-7)
    _, m32 = oracle_and_module("resnet50", 2, "fp32")
    target = torch.linspace(0.0, 1.0, 6).cuda()
    for m in (model, m32):
        m.zero_grad()
        torch.nn.functional.mse_loss(m(gt, sr), target).backward()
    for a, b in zip(model.w_layers, m32.w_layers):
        assert torch.allclose(a.weight.grad, b.weight.grad, rtol=1e-4, atol=1e-9)
        assert torch.allclose(a.bias.grad, b.bias.grad, rtol=1e-4, atol=1e-9)
    with pytest.raises(ValueError, match="even image sizes"):
        model(gt[:, :, :223], sr[:, :, :223])


def test_split_kernels_are_repeatable_under_load():
    """compute-sanitizer is closed on this pool, so races in the mbarrier / TMEM protocols of the split kernel (four TMEM
    stages, chunk sums drained while the next chunks accumulate, runtime ring split) have to show up as non-determinism:
    300 pairs (ragged micro-batches, many tiles per CTA) scored ten times, interleaved with another plan's kernels on the
    same stream and with a second stream hammering HBM, must give bit-identical scores every time."""
    oracle, mx = oracle_and_module("resnet50", 3, "fp16x3")
    _, mb = oracle_and_module("resnet50", 3, "bf16")
    g = torch.Generator(device="cuda").manual_seed(17)
    gt = torch.randn(300, 3, 224, 224, device="cuda", generator=g)
    sr = gt + 0.05 * torch.randn(300, 3, 224, 224, device="cuda", generator=g)
    noise, side = torch.empty(64 << 20, device="cuda"), torch.cuda.Stream()
    with torch.no_grad():
        mx.microbatch = 128          # 128 + 128 + 44
        first = mx(gt, sr)
        for rep in range(10):
            with torch.cuda.stream(side):
                noise.normal_()
            mb(gt[:32], sr[:32])
            mx.microbatch = (128, 77, 300)[rep % 3]
            assert torch.equal(mx(gt, sr), first), rep
    torch.cuda.synchronize()
    assert bool(torch.isfinite(first).all()) and float(first.std()) > 0


def test_normalize_has_no_gradient_path():
    """the LPIPS-style normalised variant must not hand back gradients of the un-normalised function."""
    model = CLS["resnet50"]("resnet50", 0, "cuda", precision="bf16", normalize=True)
    gt, sr = make_pairs(2, seed=3)
    with pytest.raises(NotImplementedError, match="normalize"):
        model(gt.cuda(), sr.cuda())
    with torch.no_grad():
        assert torch.isfinite(model(gt.cuda(), sr.cuda())).all()


def test_score_host_uint8_matches_device_path():
    """score_host on decoded uint8 host images == gpu_processor + forward on the device (same kernels, chunked + staged)."""
    oracle, model = oracle_and_module("resnet50", 3, "bf16")
    g = torch.Generator().manual_seed(5)
    a = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory()
    b = torch.randint(0, 256, (5, 224, 224, 3), dtype=torch.uint8, generator=g).pin_memory()
    with torch.no_grad():
        want = model.score_uint8(a.cuda(), b.cuda()).cpu()
        got = model.score_host(a, b, chunk_pairs=2)
    assert torch.equal(got, want)


def test_c_abi_error_paths():
    """Every failure is a negative return code + message, never a crash or a silent fallback."""
    import ctypes as C
    from semdiff_b200 import _lib
    oracle, model = oracle_and_module("resnet50", 0, "bf16")
    plan = model.plan()
    lib = plan.lib
    gt, sr = make_pairs(1, seed=1)
    gt, sr = gt.cuda(), sr.cuda()
    hw = torch.ones(2048, device="cuda"); hb = torch.zeros(1, device="cuda"); out = torch.empty(1, device="cuda")
    small = torch.empty(1024, dtype=torch.uint8, device="cuda")
    rc = lib.semdiff_score(plan.handle, gt.data_ptr(), sr.data_ptr(), _lib.FP32, 1, 224, 224, 1, hw.data_ptr(), hb.data_ptr(), 0,
                           small.data_ptr(), small.numel(), out.data_ptr(), None, None, _lib.stream_ptr())
    assert rc == -1 and b"workspace too small" in lib.semdiff_last_error()
    rc = lib.semdiff_score(plan.handle, None, sr.data_ptr(), _lib.FP32, 1, 224, 224, 1, hw.data_ptr(), hb.data_ptr(), 0,
                           small.data_ptr(), small.numel(), out.data_ptr(), None, None, _lib.stream_ptr())
    assert rc == -1 and b"null argument" in lib.semdiff_last_error()
    assert lib.semdiff_workspace_bytes(plan.handle, 1, 225, 224) < 0          # s2d stem layout needs even sizes
    x = torch.zeros(1, 8, 8, 24, dtype=torch.bfloat16, device="cuda")           # cin 24 is not a tensor-core shape
    w = torch.zeros(64, 24, dtype=torch.bfloat16, device="cuda"); b = torch.zeros(64, device="cuda")
    o = torch.empty(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    rc = lib.semdiff_conv2d(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, o.data_ptr(), 1, 8, 8, 24, 64, 1, 1, 1, 0, 1,
                            None, 0, 0, 0, 1, -1, _lib.BF16, _lib.CONV_TC_TMA, _lib.stream_ptr())
    assert rc == -3 and b"unsupported shape" in lib.semdiff_last_error()
    with pytest.raises(ValueError):
        model(gt, sr[:, :, :100])
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        model(gt.cpu(), sr.cpu())


def test_reference_training_loop_runs_and_tracks_the_oracle():
    """The caller in /root/reference/CLIPLPIPS_REG_training_sweep_example.py:48-100 (Adam over model.parameters(), MSE
    loss, model.train()/eval()) must work unchanged: a few steps here follow the same loss trajectory as the oracle
    trained with torch autograd (frozen eval-mode trunk)."""
    oracle, model = oracle_and_module("resnet50", 2, "fp32")
    import copy
    ref = RestatedScorer("resnet50", 2, seed=0)
    ref.load_state_dict(oracle.state_dict())
    gt, sr = make_pairs(4, seed=23)
    target = torch.tensor([0.2, 0.9, 0.4, 0.6]) * 40
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    opt_ref = torch.optim.Adam(ref.w_layers.parameters(), lr=1e-2)
    fa, fb = ref.features(gt), ref.features(sr)
    d2 = [((a - b) ** 2).detach() for a, b in zip(fa, fb)]
    losses, losses_ref = [], []
    for _ in range(4):
        model.train()
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(sr.cuda(), gt.cuda()), target.cuda())
        loss.backward()
        opt.step()
        losses.append(loss.item())
        opt_ref.zero_grad()
        out = torch.relu(torch.stack([ref.w_layers[j](d).squeeze(1).mean((-1, -2)) for j, d in enumerate(d2)]).mean(0))
        lr_ = torch.nn.functional.mse_loss(out, target)
        lr_.backward()
        opt_ref.step()
        losses_ref.append(lr_.item())
    print(f"[train] losses {losses} oracle {losses_ref}")
    assert losses[-1] < losses[0]
    assert all(abs(a - b) <= 1e-3 * abs(b) + 1e-6 for a, b in zip(losses, losses_ref))
    assert not model.clip.training and all(p.grad is None for p in model.clip.parameters())
