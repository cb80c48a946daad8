"""GPU parity of the single kernels behind the C-ABI against plain torch ops on identical tensors
(kernel unit tests)."""
import ctypes as C

import pytest
import torch

from helpers import DT, conv2d, conv_reference, lib, nchw, nhwc, sp
from semdiff_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("precision", ["bf16", "fp16", "fp32"])
def test_pack(precision):
    g = torch.Generator(device=DEV).manual_seed(0)
    gt = torch.randn(3, 3, 20, 28, device=DEV, generator=g)
    sr = torch.randn(3, 3, 20, 28, device=DEV, generator=g)
    out = torch.full((6, 20, 28, 8), 7.0, dtype=DT[precision], device=DEV)
    _lib.check(lib().semdiff_pack_input(gt.data_ptr(), sr.data_ptr(), _lib.FP32, 3, 20, 28, out.data_ptr(), _lib.PRECISIONS[precision],
                                        _lib.INPUT_NHWC8, sp()), "pack")
    ref = torch.cat([gt, sr]).permute(0, 2, 3, 1).to(DT[precision])
    assert torch.equal(out[..., :3], ref)
    assert torch.count_nonzero(out[..., 3:]) == 0


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_pack_s2d_row_window(precision):
    H, W = 12, 16
    g = torch.Generator(device=DEV).manual_seed(0)
    gt = torch.randn(2, 3, H, W, device=DEV, generator=g)
    sr = torch.randn(2, 3, H, W, device=DEV, generator=g)
    out = torch.full((4, H // 2 + 3, W // 2, 64), 7.0, dtype=DT[precision], device=DEV)
    _lib.check(lib().semdiff_pack_input(gt.data_ptr(), sr.data_ptr(), _lib.FP32, 2, H, W, out.data_ptr(), _lib.PRECISIONS[precision],
                                        _lib.INPUT_S2D_ROW4, sp()), "pack")
    x = torch.cat([gt, sr])
    ref = torch.zeros(4, H // 2 + 3, W // 2, 64, device=DEV)
    for i in range(H // 2 + 3):
        for q in range(W // 2):
            for j in range(4):
                for dy in range(2):
                    for dx in range(2):
                        y, xx = 2 * (i - 2) + dy, 2 * (q - 2 + j) + dx
                        if 0 <= y < H and 0 <= xx < W:
                            c = j * 16 + (dy * 2 + dx) * 3
                            ref[:, i, q, c:c + 3] = x[:, :, y, xx]
    assert torch.equal(out, ref.to(DT[precision]))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_pack_s2d16(precision):
    H, W = 12, 16
    gt = torch.randn(2, 3, H, W, device=DEV)
    sr = torch.randn(2, 3, H, W, device=DEV)
    out = torch.full((4, H // 2, W // 2, 16), 7.0, dtype=DT[precision], device=DEV)
    _lib.check(lib().semdiff_pack_input(gt.data_ptr(), sr.data_ptr(), _lib.FP32, 2, H, W, out.data_ptr(), _lib.PRECISIONS[precision],
                                        _lib.INPUT_S2D16, sp()), "pack")
    x = torch.cat([gt, sr])
    ref = torch.zeros(4, H // 2, W // 2, 16, device=DEV)
    for dy in range(2):
        for dx in range(2):
            ref[..., (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = x[:, :, dy::2, dx::2].permute(0, 2, 3, 1)
    assert torch.equal(out, ref.to(DT[precision]))


@pytest.mark.parametrize("layout", [_lib.INPUT_NHWC8, _lib.INPUT_S2D_ROW4, _lib.INPUT_S2D_ROW2])
def test_pack_16bit_inputs_equal_fp32_inputs_of_the_same_values(layout):
    """bf16 host images packed directly == the same values passed as fp32 (the pack kernel rounds to bf16 anyway)."""
    H, W = 16, 24
    gt = torch.randn(2, 3, H, W, device=DEV).bfloat16()
    sr = torch.randn(2, 3, H, W, device=DEV).bfloat16()
    shape = {0: (4, H, W, 8), 1: (4, H // 2 + 3, W // 2, 64), 2: (4, H // 2 + 1, W // 2, 64)}[layout]
    a = torch.full(shape, 3.0, dtype=torch.bfloat16, device=DEV)
    b = torch.full(shape, 5.0, dtype=torch.bfloat16, device=DEV)
    _lib.check(lib().semdiff_pack_input(gt.data_ptr(), sr.data_ptr(), _lib.BF16, 2, H, W, a.data_ptr(), _lib.BF16, layout, sp()), "pack")
    g32, s32 = gt.float(), sr.float()
    _lib.check(lib().semdiff_pack_input(g32.data_ptr(), s32.data_ptr(), _lib.FP32, 2, H, W, b.data_ptr(), _lib.BF16, layout, sp()), "pack")
    assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("hw", [(112, 112), (9, 7)])
def test_maxpool(precision, hw):
    H, W = hw
    x = torch.randn(2, 16, H, W, device=DEV)
    xin = nhwc(x, DT[precision])
    oh, ow = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(2, oh, ow, 16, dtype=DT[precision], device=DEV)
    _lib.check(lib().semdiff_maxpool3x3s2(xin.data_ptr(), out.data_ptr(), 2, H, W, 16, _lib.PRECISIONS[precision], sp()), "maxpool")
    ref = torch.nn.functional.max_pool2d(nchw(xin), 3, 2, 1)
    assert torch.equal(nchw(out), ref)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_avgpool(precision):
    x = torch.randn(2, 32, 14, 14, device=DEV)
    xin = nhwc(x, DT[precision])
    out = torch.empty(2, 7, 7, 32, dtype=DT[precision], device=DEV)
    _lib.check(lib().semdiff_avgpool(xin.data_ptr(), out.data_ptr(), 2, 14, 14, 32, 2, _lib.PRECISIONS[precision], sp()), "avgpool")
    ref = torch.nn.functional.avg_pool2d(nchw(xin), 2)
    tol = 1e-6 if precision == "fp32" else 8e-3
    assert torch.allclose(nchw(out), ref, rtol=tol, atol=tol)


CONV_CASES = [
    # n, H, W, cin, cout, k, stride, pad, residual
    (2, 56, 56, 64, 64, 1, 1, 0, False),
    (2, 56, 56, 64, 256, 1, 1, 0, True),
    (3, 14, 14, 256, 256, 3, 1, 1, False),    # M = 588: ragged last tile
    (2, 28, 28, 128, 128, 3, 2, 1, False),
    (2, 28, 28, 256, 512, 1, 2, 0, False),    # strided 1x1 (downsample)
    (2, 32, 32, 8, 64, 7, 2, 3, False),       # stem geometry (Cin padded 3 -> 8)
    (1, 7, 7, 512, 2048, 1, 1, 0, True),
    (2, 16, 16, 32, 32, 3, 1, 1, False),      # CLIP stem conv2
]


def _conv_case(case, precision, impl, seed=0):
    n, H, W, cin, cout, k, stride, pad, use_res = case
    g = torch.Generator(device=DEV).manual_seed(seed)
    dt = DT[precision]
    x = torch.randn(n, H, W, cin, device=DEV, generator=g).to(dt)
    w = (torch.randn(cout, k, k, cin, device=DEV, generator=g) * (2.0 / (k * k * cin)) ** 0.5).to(dt)
    b = torch.randn(cout, device=DEV, generator=g) * 0.1
    oh, ow = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = torch.randn(n, oh, ow, cout, device=DEV, generator=g).to(dt) if use_res else None
    out = conv2d(x, w, b, res, stride, pad, True, precision, impl)
    torch.cuda.synchronize()
    ref = conv_reference(x, w, b, res, stride, pad, True)
    return out.double(), ref


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_fp32(case):
    out, ref = _conv_case(case, "fp32", _lib.CONV_SIMT)
    err = (out - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_bf16(case):
    out, ref = _conv_case(case, "bf16", _lib.CONV_SIMT)
    # only the final rounding to bf16 differs from the fp64 reference on identical operands
    assert torch.allclose(out, ref, rtol=8e-3, atol=8e-3), (out - ref).abs().max().item()


@pytest.mark.parametrize("precision,impl", [("fp32", _lib.CONV_SIMT), ("bf16", _lib.CONV_SIMT), ("bf16", _lib.CONV_TC_TMA)])
@pytest.mark.parametrize("geom", [(2, 56, 56, 64, 64, 256, 1), (3, 14, 14, 256, 512, 1024, 2), (1, 7, 7, 512, 1024, 2048, 2)])
def test_conv_with_fused_projection_shortcut(geom, precision, impl):
    """conv3(t2) + downsample(x) in ONE launch (second activation source appended along K)."""
    n, oh, ow, c1, c2, cout, s2 = geom
    g = torch.Generator(device=DEV).manual_seed(7)
    dt = DT[precision]
    x1 = torch.randn(n, oh, ow, c1, device=DEV, generator=g).to(dt)
    H2, W2 = oh * s2, ow * s2
    x2 = torch.randn(n, H2, W2, c2, device=DEV, generator=g).to(dt)
    w1 = (torch.randn(cout, 1, 1, c1, device=DEV, generator=g) * (1.0 / c1) ** 0.5).to(dt)
    w2 = (torch.randn(cout, c2, device=DEV, generator=g) * (1.0 / c2) ** 0.5).to(dt)
    b = torch.randn(cout, device=DEV, generator=g) * 0.1
    out = conv2d(x1, w1, b, None, 1, 0, True, precision, impl, x2_nhwc=x2, w2=w2, stride2=s2).double()
    ref = conv_reference(x1, w1, b, None, 1, 0, False) + conv_reference(x2, w2.reshape(cout, 1, 1, c2), torch.zeros_like(b), None, s2, 0, False)
    ref = torch.relu(ref)
    tol = 2e-5 if precision == "fp32" else 8e-3
    assert torch.allclose(out, ref, rtol=tol, atol=tol), (out - ref).abs().max().item()


def _distance_ref(act, n_pairs, w):
    a, b = act[:n_pairs].double(), act[n_pairs:].double()
    return (((a - b) ** 2) * w.double()).sum(dim=(1, 2))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("shape", [(56 * 56, 256), (49, 2048), (28 * 28, 512), (10, 24)])
def test_layer_distance_and_head(precision, shape):
    hw, c = shape
    n_pairs = 3
    g = torch.Generator(device=DEV).manual_seed(1)
    act = torch.randn(2 * n_pairs, hw, c, device=DEV, generator=g).to(DT[precision])
    w = torch.randn(c, device=DEV, generator=g)
    parts = lib().semdiff_distance_parts(hw, c)
    assert 1 <= parts <= _lib.MAX_PARTS
    partial = torch.zeros(n_pairs, _lib.MAX_PARTS, device=DEV)
    chan = torch.zeros(n_pairs, c, device=DEV)
    _lib.check(lib().semdiff_layer_distance(act.data_ptr(), n_pairs, hw, c, w.data_ptr(), 0, partial.data_ptr(),
                                            chan.data_ptr(), c, _lib.PRECISIONS[precision], sp()), "distance")
    got = partial[:, :parts].double().sum(1)
    ref = _distance_ref(act, n_pairs, w)
    assert torch.allclose(got, ref, rtol=2e-5, atol=1e-4), (got, ref)
    cm_ref = ((act[:n_pairs].double() - act[n_pairs:].double()) ** 2).mean(1)
    assert torch.allclose(chan.double(), cm_ref, rtol=1e-5, atol=1e-6)
    # head on this single layer
    bias = torch.tensor([0.25], device=DEV)
    out, pre = torch.empty(n_pairs, device=DEV), torch.empty(n_pairs, device=DEV)
    np_arr, hw_arr = (C.c_int32 * 1)(parts), (C.c_int32 * 1)(hw)
    _lib.check(lib().semdiff_head(partial.data_ptr(), 1, n_pairs, np_arr, hw_arr, bias.data_ptr(), out.data_ptr(),
                                  pre.data_ptr(), sp()), "head")
    ref_pre = ref / hw + 0.25
    assert torch.allclose(pre.double(), ref_pre, rtol=2e-5, atol=1e-5)
    assert torch.equal(out, torch.relu(pre))


def test_layer_distance_batch_independent():
    """A pair's partial sums must not depend on batch size / position (bit-identical sharding)."""
    hw, c = 28 * 28, 512
    g = torch.Generator(device=DEV).manual_seed(2)
    act = torch.randn(8, hw, c, device=DEV, generator=g).bfloat16()
    w = torch.randn(c, device=DEV, generator=g)
    p4 = torch.zeros(4, _lib.MAX_PARTS, device=DEV)
    _lib.check(lib().semdiff_layer_distance(act.data_ptr(), 4, hw, c, w.data_ptr(), 0, p4.data_ptr(), None, 0, _lib.BF16, sp()), "d")
    sub = torch.stack([act[2], act[6]]).contiguous()
    p1 = torch.zeros(1, _lib.MAX_PARTS, device=DEV)
    _lib.check(lib().semdiff_layer_distance(sub.data_ptr(), 1, hw, c, w.data_ptr(), 0, p1.data_ptr(), None, 0, _lib.BF16, sp()), "d")
    assert torch.equal(p4[2], p1[0])


def test_layer_distance_normalized():
    hw, c = 49, 256
    g = torch.Generator(device=DEV).manual_seed(3)
    act = torch.randn(4, hw, c, device=DEV, generator=g).bfloat16()
    w = torch.rand(c, device=DEV, generator=g)
    parts = lib().semdiff_distance_parts(hw, c)
    partial = torch.zeros(2, _lib.MAX_PARTS, device=DEV)
    _lib.check(lib().semdiff_layer_distance(act.data_ptr(), 2, hw, c, w.data_ptr(), 1, partial.data_ptr(), None, 0, _lib.BF16, sp()), "d")
    x = act.double()
    x = x / (x.norm(dim=2, keepdim=True) + 1e-10)
    ref = (((x[:2] - x[2:]) ** 2) * w.double()).sum(dim=(1, 2))
    assert torch.allclose(partial[:, :parts].double().sum(1), ref, rtol=1e-4, atol=1e-6)
