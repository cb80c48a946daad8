"""GPU parity of the tcgen05/TMEM/TMA implicit-GEMM conv (csrc/conv_tc.cu) against an fp64 torch conv on the
same 16-bit operands.  Tolerance: fp32 accumulation of <= 4608 products of O(1) terms, then ONE rounding to the
16-bit output type: |err| <= 2^-8 * |ref| + 2e-3 for bf16 (2^-11 for fp16)."""
import pytest
import torch

from semdiff_b200 import _lib
from test_kernels_gpu import CONV_CASES, _conv_case

pytestmark = pytest.mark.gpu


def _check(out, ref, precision):
    rel = 2.0 ** -8 if precision == "bf16" else 2.0 ** -11
    bad = (out - ref).abs() > rel * ref.abs() + 2e-3
    assert not bad.any(), f"{int(bad.sum())} / {bad.numel()} mismatches, max abs err {(out - ref).abs().max().item():.4g}"


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tc_gather(case, precision):
    out, ref = _conv_case(case, precision, _lib.CONV_TC_GATHER)
    _check(out, ref, precision)


IM2COL_CASES = [
    # n, H, W, cin, cout, k, stride, pad, residual  (TMA im2col mode: anything but 1x1 stride 1, Cin % 64 == 0)
    (2, 56, 56, 64, 64, 3, 1, 1, False),      # tiles span image rows and image boundaries
    (3, 56, 40, 128, 128, 3, 2, 1, False),    # H != W, stride 2
    (5, 7, 7, 512, 512, 3, 1, 1, False),      # M = 245: tile crosses several images, ragged tail
    (2, 14, 14, 1024, 2048, 1, 2, 0, False),  # strided 1x1
    (1, 9, 9, 64, 64, 3, 1, 1, False),        # tensor < 128 KiB (descriptor workaround path)
    (2, 15, 15, 64, 128, 3, 2, 1, False),     # odd size, stride 2
]


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[3] % 64 == 0] + IM2COL_CASES)
def test_conv_tc_tma(case, precision):
    out, ref = _conv_case(case, precision, _lib.CONV_TC_TMA)
    _check(out, ref, precision)


STRIP_CASES = [
    # 3x3 stride-1 pad-1 64 -> 64 (csrc/conv3x3_strip.cu): every strip pitch, ragged row groups, several images
    (3, 56, 56, 64, 64, 3, 1, 1, False),     # P = 64, 2 output rows per tile
    (2, 28, 28, 64, 64, 3, 1, 1, False),     # P = 32, 4 rows per tile
    (5, 14, 14, 64, 64, 3, 1, 1, False),     # P = 16, 8 rows per tile (14 = 8 + 6: ragged)
    (2, 112, 112, 64, 64, 3, 1, 1, False),   # P = 128, 1 row per tile (CLIP stem geometry)
    (2, 37, 50, 64, 64, 3, 1, 1, False),     # H odd (ragged last row group), W not a power of two
    (1, 6, 126, 64, 64, 3, 1, 1, False),     # W + 2 == 128: the widest row one strip holds
    (2, 9, 256, 64, 64, 3, 1, 1, False),     # wider: 3 column blocks of 126 (last one partial), 1024x1024 layer1 geometry
    (1, 5, 127, 64, 64, 3, 1, 1, False),     # 2 column blocks, the second a single pixel wide
]


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("case", STRIP_CASES)
def test_conv3x3_strip(case, precision):
    out, ref = _conv_case(case, precision, _lib.CONV_TC_TMA)
    _check(out, ref, precision)


@pytest.mark.parametrize("geom", [(2, 115, 112, 4), (3, 59, 56, 4), (2, 113, 112, 2), (1, 20, 14, 3), (2, 12, 126, 4)])
def test_conv_rowwindow_strip(geom):
    """KH x 1 pad-0 'row-window' stems (4x1 over the S2D_ROW4 input, 2x1 over S2D_ROW2): strip kernel, RG = 2."""
    import torch
    from helpers import conv2d, conv_reference
    n, H, W, kh = geom
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n, H, W, 64, device="cuda", generator=g).bfloat16()
    w = (torch.randn(64, kh, 1, 64, device="cuda", generator=g) * (2.0 / (kh * 64)) ** 0.5).bfloat16()
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    # the helper takes one padding for both axes: 0 here
    out = conv2d(x, w, b, None, 1, 0, True, "bf16", _lib.CONV_TC_TMA).double()
    ref = torch.relu(torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(0, 3, 1, 2), b.double())).permute(0, 2, 3, 1)
    _check(out, ref, "bf16")


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("geom", [(2, 112, 112), (3, 56, 56), (1, 11, 125), (2, 7, 20), (1, 6, 512), (2, 5, 126)])
def test_conv_s2d16_stem_strip(geom, precision):
    """4x4 stride-1 conv, padding 2 before / 1 after, over 16-channel pixels (the 7x7/2 stem on SEMDIFF_INPUT_S2D16):
    strip kernel with 32-byte rows (SWIZZLE_32B views shifted by whole pixels)."""
    import torch
    from helpers import DT, conv2d
    n, H, W = geom
    g = torch.Generator(device="cuda").manual_seed(19)
    dt = DT[precision]
    x = torch.randn(n, H, W, 16, device="cuda", generator=g).to(dt)
    w = (torch.randn(64, 4, 4, 16, device="cuda", generator=g) * (2.0 / 256) ** 0.5).to(dt)
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    out = conv2d(x, w, b, None, 1, 2, True, precision, _lib.CONV_TC_TMA, pad_hi=1).double()
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (2, 1, 2, 1))
    ref = torch.relu(torch.nn.functional.conv2d(xp, w.double().permute(0, 3, 1, 2), b.double())).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    _check(out, ref, precision)


@pytest.mark.parametrize("geom", [(2, 112, 112), (3, 20, 37)])
def test_conv_s2d16_clip_stem_strip(geom):
    """2x2 stride-1 conv, padding 1 before / 0 after, over 16-channel pixels (the CLIP 3x3/2 stem on SEMDIFF_INPUT_S2D16)."""
    import torch
    from helpers import conv2d
    n, H, W = geom
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(n, H, W, 16, device="cuda", generator=g).bfloat16()
    w = (torch.randn(64, 2, 2, 16, device="cuda", generator=g) * (2.0 / 64) ** 0.5).bfloat16()
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    out = conv2d(x, w, b, None, 1, 1, True, "bf16", _lib.CONV_TC_TMA, pad_hi=0).double()
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (1, 0, 1, 0))
    ref = torch.relu(torch.nn.functional.conv2d(xp, w.double().permute(0, 3, 1, 2), b.double())).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    _check(out, ref, "bf16")


def test_conv_tc_many_tiles_persistent():
    """More tiles than SMs: exercises the persistent loop, both TMEM accumulator stages and smem ring wrap."""
    case = (8, 56, 56, 64, 256, 1, 1, 0, True)     # M = 25088 -> 196 m-tiles x 1 n-tile
    for impl in (_lib.CONV_TC_TMA, _lib.CONV_TC_GATHER):
        out, ref = _conv_case(case, "bf16", impl, seed=3)
        _check(out, ref, "bf16")
    case = (4, 28, 28, 128, 128, 3, 1, 1, False)   # K = 1152 -> 18 k-blocks per tile
    out, ref = _conv_case(case, "bf16", _lib.CONV_TC_GATHER, seed=4)
    _check(out, ref, "bf16")


POOL_GEOMS = [
    # n, H, W of the space-to-depth input (= conv output size)
    (3, 112, 112),   # the 224 x 224 geometry: P = 128, one conv row per row group, 28 tiles per image
    (300, 16, 24),   # many small images: ranges of contiguous tiles start mid-image on most CTAs (warm-up tiles)
    (5, 56, 56),     # P = 64: two conv rows per row group
    (2, 37, 50),     # odd conv height: ragged last tile, pooled row that reaches below the image
    (2, 9, 125),     # the widest row one strip holds, odd height
    (1, 4, 8),       # a single tile
    (150, 112, 8),   # tall, narrow: 28 tiles per image, CTAs change image inside their range
    (40, 30, 100),   # P = 128 (register pooling path): many images, ranges start mid-image, ragged last tile (30 = 7 * 4 + 2)
    (7, 113, 63),    # P = 128, odd height and odd width: last pooled row / column have two taps
    (2, 5, 64),      # P = 128 by tie-break, W a power of two
]


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("geom", POOL_GEOMS)
def test_stem_conv_maxpool_fused(geom, precision):
    """Stem conv (S2D16 layout) with max_pool2d(3, 2, 1) in the epilogue (csrc/conv3x3_strip.cu, kPool) == the strip conv
    followed by semdiff_maxpool3x3s2, bit for bit (max of the same 16-bit values), and == torch max_pool2d of it."""
    import torch
    from helpers import DT, conv2d, lib, sp
    n, H, W = geom
    g = torch.Generator(device="cuda").manual_seed(29 + H + W)
    dt = DT[precision]
    x = torch.randn(n, H, W, 16, device="cuda", generator=g).to(dt)
    w = (torch.randn(64, 4, 4, 16, device="cuda", generator=g) * (2.0 / 256) ** 0.5).to(dt)
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    for relu in (True, False):
        conv = conv2d(x, w, b, None, 1, 2, relu, precision, _lib.CONV_TC_TMA, pad_hi=1)
        ph, pw = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        guard = 8192
        arena = torch.full((n * ph * pw * 64 + 2 * guard,), 123.0, dtype=dt, device="cuda")
        out = arena[guard:guard + n * ph * pw * 64].view(n, ph, pw, 64)
        rc = lib().semdiff_conv2d_maxpool(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, H, W, 16, 64, 4, 4, 2, 1,
                                          int(relu), _lib.PRECISIONS[precision], sp())
        _lib.check(rc, "semdiff_conv2d_maxpool")
        torch.cuda.synchronize()
        assert bool((arena[:guard] == 123.0).all()) and bool((arena[guard + out.numel():] == 123.0).all()), "wrote outside its output"
        ref = torch.nn.functional.max_pool2d(conv.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).to(dt)
        assert torch.equal(out, ref), f"relu={relu}: {(out.float() - ref.float()).abs().max().item()} max abs diff, " \
                                      f"{int((out != ref).sum())} / {out.numel()} elements"
        sep = torch.empty_like(ref)
        _lib.check(lib().semdiff_maxpool3x3s2(conv.data_ptr(), sep.data_ptr(), n, H, W, 64, _lib.PRECISIONS[precision], sp()), "maxpool")
        torch.cuda.synchronize()
        assert torch.equal(out, sep)


def test_stem_conv_maxpool_unsupported_width():
    import torch
    from helpers import lib, sp
    x = torch.zeros(1, 8, 200, 16, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(64, 4, 4, 16, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(64, device="cuda")
    out = torch.zeros(1, 4, 100, 64, device="cuda", dtype=torch.bfloat16)
    rc = lib().semdiff_conv2d_maxpool(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 1, 8, 200, 16, 64, 4, 4, 2, 1, 1, 0, sp())
    assert rc == -3 and b"conv_strip_pool" in lib().semdiff_last_error()


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("cin", [64, 32])
@pytest.mark.parametrize("geom", [(3, 112, 112), (40, 30, 100), (2, 8, 62), (2, 6, 126), (5, 112, 63), (300, 2, 64)])
def test_conv3x3_avgpool_fused(geom, precision, cin):
    """3x3 64 -> 64 conv with avg_pool2d(2) in the epilogue (csrc/conv3x3_strip.cu, kPool = 2: CLIP stem.conv3 + stem.pool)
    == the strip conv followed by semdiff_avgpool, bit for bit (same summation order, one rounding)."""
    import torch
    from helpers import DT, conv2d, lib, sp
    n, H, W = geom
    g = torch.Generator(device="cuda").manual_seed(31 + H + W)
    dt = DT[precision]
    x = torch.randn(n, H, W, cin, device="cuda", generator=g).to(dt)
    w = (torch.randn(64, 3, 3, cin, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5).to(dt)
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    conv = conv2d(x, w, b, None, 1, 1, True, precision, _lib.CONV_TC_TMA)
    ph, pw = H // 2, W // 2
    guard = 8192
    arena = torch.full((n * ph * pw * 64 + 2 * guard,), 123.0, dtype=dt, device="cuda")
    out = arena[guard:guard + n * ph * pw * 64].view(n, ph, pw, 64)
    rc = lib().semdiff_conv2d_avgpool(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), n, H, W, cin, 1, _lib.PRECISIONS[precision], sp())
    _lib.check(rc, "semdiff_conv2d_avgpool")
    torch.cuda.synchronize()
    assert bool((arena[:guard] == 123.0).all()) and bool((arena[guard + out.numel():] == 123.0).all()), "wrote outside its output"
    sep = torch.empty_like(out)
    _lib.check(lib().semdiff_avgpool(conv.data_ptr(), sep.data_ptr(), n, H, W, 64, 2, _lib.PRECISIONS[precision], sp()), "avgpool")
    torch.cuda.synchronize()
    assert torch.equal(out, sep), f"{int((out != sep).sum())} / {out.numel()} elements differ, max {(out.float() - sep.float()).abs().max().item()}"
    ref = torch.nn.functional.avg_pool2d(conv.float().permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2.0 ** -7 * ref.abs().max().item()


STRIP32_CASES = [
    # n, H, W, cin, cout: 3x3 stride-1 pad-1 over 32-channel pixels (64-byte rows, SWIZZLE_64B): CLIP stem conv2 / conv3
    (3, 112, 112, 32, 32), (3, 112, 112, 32, 64), (2, 56, 56, 32, 32), (5, 14, 14, 32, 64), (2, 37, 50, 32, 32),
    (1, 6, 126, 32, 64), (2, 9, 256, 32, 32), (1, 5, 127, 32, 64),
]


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("case", STRIP32_CASES)
def test_conv3x3_strip_32_channels(case, precision):
    import torch
    from helpers import DT, conv2d, conv_reference
    n, H, W, cin, cout = case
    g = torch.Generator(device="cuda").manual_seed(37 + H + W + cout)
    dt = DT[precision]
    x = torch.randn(n, H, W, cin, device="cuda", generator=g).to(dt)
    w = (torch.randn(cout, 3, 3, cin, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5).to(dt)
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    out = conv2d(x, w, b, None, 1, 1, True, precision, _lib.CONV_TC_TMA).double()
    _check(out, conv_reference(x, w, b, None, 1, 1, True), precision)


@pytest.mark.parametrize("geom", [(2, 112, 112), (3, 20, 37), (1, 7, 300)])
def test_conv_s2d16_clip_stem_strip_32_outputs(geom):
    """The CLIP 3x3/2 stem conv over SEMDIFF_INPUT_S2D16 with its real 32 output channels (64-byte output rows)."""
    import torch
    from helpers import conv2d
    n, H, W = geom
    g = torch.Generator(device="cuda").manual_seed(41)
    x = torch.randn(n, H, W, 16, device="cuda", generator=g).bfloat16()
    w = (torch.randn(32, 2, 2, 16, device="cuda", generator=g) * (2.0 / 64) ** 0.5).bfloat16()
    b = torch.randn(32, device="cuda", generator=g) * 0.1
    out = conv2d(x, w, b, None, 1, 1, True, "bf16", _lib.CONV_TC_TMA, pad_hi=0).double()
    xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (1, 0, 1, 0))
    ref = torch.relu(torch.nn.functional.conv2d(xp, w.double().permute(0, 3, 1, 2), b.double())).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    _check(out, ref, "bf16")
