"""GPU parity of the split-precision kernels (SEMDIFF_FP16X3 / SEMDIFF_BF16X3, include/semdiff_b200.h): every value is
hi + lo of two 16-bit numbers and every conv is three tensor-core products per K block (Ah*Wh + Al*Wh + Ah*Wl).
References are fp64 torch ops on the SAME (already split) operand values, so the tolerances measure the kernels, not
the storage type: fp16x3 carries 22 significant bits, bf16x3 16."""
import pytest
import torch

from helpers import DT, conv2d, lib, sp, split_store, split_value
from semdiff_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"
# relative to the largest magnitude of the reference tensor: dropped Al*Wl term + output rounding of hi + lo
TOL = {"fp16x3": 3e-6, "bf16x3": 2e-4}

SPLIT_CONV_CASES = [
    # n, H, W, cin, cout, k, stride, pad, residual
    (2, 56, 56, 64, 64, 1, 1, 0, False),       # BLOCK_N = 64
    (2, 56, 56, 64, 256, 1, 1, 0, True),       # residual (hi + lo tiles prefetched into the C ring)
    (3, 14, 14, 256, 256, 3, 1, 1, False),     # im2col, M = 588: ragged last tile, BLOCK_N = 256
    (2, 28, 28, 128, 128, 3, 2, 1, False),     # strided im2col
    (5, 7, 7, 512, 2048, 1, 1, 0, True),       # many n-tiles, tile crosses images
    (2, 14, 14, 1024, 256, 1, 1, 0, False),    # long K (48 ring fills per tile)
    (1, 9, 9, 64, 64, 3, 1, 1, False),         # tiny tensor (im2col descriptor workaround path)
    (2, 31, 16, 64, 64, 4, 1, 0, False),       # 4x4 pad-0 window (the ROW4 stem is 4x1)
    (16, 56, 56, 64, 64, 3, 1, 1, False),      # 392 tiles on 148 CTAs: ring, TMEM-stage and C-ring counters wrap across tiles
    (8, 56, 56, 128, 256, 1, 1, 0, True),      # 392 tiles x 2 n-tiles with residual prefetch across tile boundaries
    (6, 28, 28, 1024, 256, 1, 1, 0, False),    # 256-wide tile (16 K blocks): 4 output groups through the single C slot
]


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
@pytest.mark.parametrize("case", SPLIT_CONV_CASES)
def test_conv_split(case, precision):
    n, H, W, cin, cout, k, stride, pad, with_res = case
    dt = DT[precision]
    g = torch.Generator(device=DEV).manual_seed(7)
    x = split_store(torch.randn(n, H, W, cin, device=DEV, generator=g), dt)
    w = split_store(torch.randn(cout, k, k, cin, device=DEV, generator=g) * (2.0 / (k * k * cin)) ** 0.5, dt)
    b = torch.randn(cout, device=DEV, generator=g) * 0.1
    oh, ow = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = split_store(torch.randn(n, oh, ow, cout, device=DEV, generator=g), dt) if with_res else None
    out = split_value(conv2d(x, w, b, res, stride, pad, True, precision, _lib.CONV_TC_TMA))
    ref = torch.nn.functional.conv2d(split_value(x).permute(0, 3, 1, 2), split_value(w).permute(0, 3, 1, 2), b.double(),
                                     stride=stride, padding=pad).permute(0, 2, 3, 1)
    if with_res:
        ref = ref + split_value(res)
    ref = torch.relu(ref)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"[split conv] {case} {precision}: max err / max |ref| = {err:.3g}")
    assert out.shape == ref.shape and err < TOL[precision]


def test_conv_split_random_shapes():
    """Seeded random geometries (odd sizes, strides, 1x1 / 3x3 / 5x5 windows, channel counts that are not powers of two,
    residual on / off): every combination of tile width, A mode and ring configuration the host can pick."""
    import random
    rng = random.Random(1234)
    for trial in range(16):
        k = rng.choice([1, 1, 3, 3, 5])
        stride = rng.choice([1, 1, 2])
        cin, cout = rng.choice([64, 128, 192, 320]), rng.choice([64, 128, 192, 256, 512])
        n, H, W = rng.randint(1, 5), rng.randint(k + 2, 33), rng.randint(k + 2, 33)
        with_res = rng.random() < 0.4
        pad = k // 2
        g = torch.Generator(device=DEV).manual_seed(100 + trial)
        x = split_store(torch.randn(n, H, W, cin, device=DEV, generator=g), torch.float16)
        w = split_store(torch.randn(cout, k, k, cin, device=DEV, generator=g) * (2.0 / (k * k * cin)) ** 0.5, torch.float16)
        b = torch.randn(cout, device=DEV, generator=g) * 0.1
        oh, ow = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        res = split_store(torch.randn(n, oh, ow, cout, device=DEV, generator=g), torch.float16) if with_res else None
        out = split_value(conv2d(x, w, b, res, stride, pad, True, "fp16x3", _lib.CONV_TC_TMA))
        ref = torch.nn.functional.conv2d(split_value(x).permute(0, 3, 1, 2), split_value(w).permute(0, 3, 1, 2), b.double(),
                                         stride=stride, padding=pad).permute(0, 2, 3, 1)
        ref = torch.relu(ref + split_value(res)) if with_res else torch.relu(ref)
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        assert err < TOL["fp16x3"], (trial, n, H, W, cin, cout, k, stride, with_res, err)


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
def test_conv_split_second_source(precision):
    """Projection shortcut folded into conv3's launch: K blocks of a second, strided activation tensor."""
    dt = DT[precision]
    g = torch.Generator(device=DEV).manual_seed(3)
    n, H, W, cin, cout, cin2 = 2, 14, 14, 128, 512, 256
    x = split_store(torch.randn(n, H, W, cin, device=DEV, generator=g), dt)
    x2 = split_store(torch.randn(n, 2 * H, 2 * W, cin2, device=DEV, generator=g), dt)
    w = split_store(torch.randn(cout, 1, 1, cin, device=DEV, generator=g) * 0.1, dt)
    w2 = split_store(torch.randn(cout, cin2, device=DEV, generator=g) * 0.1, dt)
    b = torch.randn(cout, device=DEV, generator=g) * 0.1
    out = split_value(conv2d(x, w, b, None, 1, 0, True, precision, _lib.CONV_TC_TMA, x2_nhwc=x2, w2=w2, stride2=2))
    ref = torch.nn.functional.conv2d(split_value(x).permute(0, 3, 1, 2), split_value(w).permute(0, 3, 1, 2), b.double())
    ref = ref + torch.nn.functional.conv2d(split_value(x2).permute(0, 3, 1, 2), split_value(w2)[:, :, None, None], stride=2)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f"[split conv, second source] {precision}: {err:.3g}")
    assert err < TOL[precision]


def test_conv_split_accumulation_long_k():
    """Isolates the tensor-core accumulation: operands exactly representable in fp16 (lo halves zero), K = 4608 - the
    error left is the fp32 accumulator's, which bounds what the fp16x3 trunk can reach."""
    g = torch.Generator(device=DEV).manual_seed(5)
    n, H, W, cin, cout = 4, 7, 7, 512, 512
    x = split_store(torch.randn(n, H, W, cin, device=DEV, generator=g).half(), torch.float16)
    w = split_store((torch.randn(cout, 3, 3, cin, device=DEV, generator=g) * 0.02).half(), torch.float16)
    b = torch.zeros(cout, device=DEV)
    out = split_value(conv2d(x, w, b, None, 1, 1, False, "fp16x3", _lib.CONV_TC_TMA))
    ref = torch.nn.functional.conv2d(split_value(x).permute(0, 3, 1, 2), split_value(w).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    ref32 = torch.nn.functional.conv2d(split_value(x).float().permute(0, 3, 1, 2), split_value(w).float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    scale = ref.abs().max()
    err, err32 = ((out - ref).abs().max() / scale).item(), ((ref32.double() - ref).abs().max() / scale).item()
    bias = (((out - ref) * ref.sign()).mean() / ref.abs().mean()).item()   # < 0: truncation towards zero
    print(f"[split conv] K = 4608 accumulation: tcgen05 max err {err:.3g} (torch fp32 conv {err32:.3g}), mean (out - ref) sign(ref) / mean |ref| {bias:.3g}")
    assert err < 2e-6


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
def test_pack_split_row_window(precision):
    """pack -> SEMDIFF_INPUT_S2D_ROW4 in split storage: hi + lo reproduces the fp32 image to the split resolution and the
    hi halves are exactly the 16-bit pack."""
    H, W = 12, 16
    dt = DT[precision]
    g = torch.Generator(device=DEV).manual_seed(0)
    gt = torch.randn(2, 3, H, W, device=DEV, generator=g)
    sr = torch.randn(2, 3, H, W, device=DEV, generator=g)
    out = torch.full((4, H // 2 + 3, W // 2, 128), 7.0, dtype=dt, device=DEV)
    _lib.check(lib().semdiff_pack_input(gt.data_ptr(), sr.data_ptr(), _lib.FP32, 2, H, W, out.data_ptr(), _lib.PRECISIONS[precision],
                                        _lib.INPUT_S2D_ROW4, sp()), "pack")
    plain = torch.full((4, H // 2 + 3, W // 2, 64), 7.0, dtype=torch.float32, device=DEV)
    _lib.check(lib().semdiff_pack_input(gt.data_ptr(), sr.data_ptr(), _lib.FP32, 2, H, W, plain.data_ptr(), _lib.FP32,
                                        _lib.INPUT_S2D_ROW4, sp()), "pack")
    assert torch.equal(out, split_store(plain, dt))


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
def test_pools_split(precision):
    dt = DT[precision]
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(2, 13, 9, 128, device=DEV, generator=g)
    x[0, 3:6, 2:5, :] = x[0, 4, 3, :]          # ties in hi AND lo
    base = (x[1, :, :, :64] * 4).round() / 4
    x[1, :, :, :64] = base + 1e-4 * torch.randn(base.shape, device=DEV, generator=g)   # many equal hi halves, different lo
    xs = split_store(x, dt)
    xv = split_value(xs)
    out = torch.empty(2, 7, 5, 256, dtype=dt, device=DEV)
    _lib.check(lib().semdiff_maxpool3x3s2(xs.data_ptr(), out.data_ptr(), 2, 13, 9, 128, _lib.PRECISIONS[precision], sp()), "maxpool")
    ref = torch.nn.functional.max_pool2d(xv.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(split_value(out), ref)      # the maximum of the represented values, exactly
    out2 = torch.empty(2, 6, 4, 256, dtype=dt, device=DEV)
    _lib.check(lib().semdiff_avgpool(xs.data_ptr(), out2.data_ptr(), 2, 13, 9, 128, 2, _lib.PRECISIONS[precision], sp()), "avgpool")
    ref2 = torch.nn.functional.avg_pool2d(xv.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert ((split_value(out2) - ref2).abs().max() / ref2.abs().max()).item() < TOL[precision]


@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3"])
@pytest.mark.parametrize("shape", [(3, 56 * 56, 256), (2, 49, 2048), (2, 30, 192)])
@pytest.mark.parametrize("normalize", [0, 1])
def test_layer_distance_split(shape, precision, normalize):
    """(:379-384) on split taps, incl. SR ~ GT pairs whose difference lives entirely in the lo halves."""
    n_pairs, hw, c = shape
    dt = DT[precision]
    g = torch.Generator(device=DEV).manual_seed(2)
    a = torch.randn(n_pairs, hw, c, device=DEV, generator=g).abs()
    bb = a + torch.tensor([1e-4, 3e-2, 1.0], device=DEV)[:n_pairs, None, None] * torch.randn(n_pairs, hw, c, device=DEV, generator=g)
    act = split_store(torch.cat([a, bb]), dt)
    w = torch.rand(c, device=DEV, generator=g)
    parts = lib().semdiff_distance_parts(hw, c)
    partial = torch.zeros(n_pairs, _lib.MAX_PARTS, device=DEV)
    chan = torch.zeros(n_pairs, c, device=DEV)
    _lib.check(lib().semdiff_layer_distance(act.data_ptr(), n_pairs, hw, c, w.data_ptr(), normalize, partial.data_ptr(),
                                            chan.data_ptr(), c, _lib.PRECISIONS[precision], sp()), "distance")
    v = split_value(act)
    va, vb = v[:n_pairs], v[n_pairs:]
    if normalize:
        va = va / (va.norm(dim=-1, keepdim=True) + 1e-10)
        vb = vb / (vb.norm(dim=-1, keepdim=True) + 1e-10)
    d2 = (va - vb) ** 2
    ref = (d2 * w.double()).sum((1, 2))
    got = partial[:, :parts].double().sum(1)
    rel = ((got - ref).abs() / ref).max().item()
    print(f"[split distance] {shape} {precision} normalize={normalize}: {rel:.3g}")
    assert rel < (5e-5 if normalize else 2e-6)   # the unit-normalisation itself runs in fp32
    if not normalize:
        refc = (split_value(act)[:n_pairs] - split_value(act)[n_pairs:]) ** 2
        assert torch.allclose(chan.double(), refc.mean(1), rtol=1e-5, atol=1e-12)
