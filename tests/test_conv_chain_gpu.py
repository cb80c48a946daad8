"""GPU parity of the chained pointwise convs (csrc/conv_chain.cu): conv3 (+residual | +fused shortcut) of one bottleneck
and conv1 of the next, one launch.  The chain must produce EXACTLY what two semdiff_conv2d launches produce (the second
GEMM consumes the same rounded 16-bit tile from shared memory that the unfused path re-reads from HBM; same k order),
and both outputs must match an fp64 torch reference on the same operands within one 16-bit rounding."""
import pytest
import torch

from helpers import DT, conv2d, lib, sp
from semdiff_b200 import _lib

pytestmark = pytest.mark.gpu

CHAIN_CASES = [
    # m (pixels), cin, cin2, cout2, residual [, cout1 = 256]
    (128 * 300, 64, 0, 64, True),      # layer1 identity block -> next conv1 (many tiles per CTA: every ring wraps)
    (128 * 300 + 77, 64, 0, 64, True), # ragged pixel tail
    (50, 64, 0, 64, True),             # less than one tile
    (128 * 200 + 5, 64, 0, 128, True), # layer1 -> layer2 boundary (cout2 = 128)
    (128 * 200 + 5, 64, 64, 64, False),# first block: projection shortcut fused as second source, no residual
    (128 * 149, 128, 0, 64, True),     # two k-blocks from one source
    (128 * 150, 64, 0, 64, False),     # plain (no residual, no second source)
    # 512-channel stage (weights streamed per chunk): identity block -> next conv1
    (128 * 300 + 41, 128, 0, 128, True, 512),
    (128 * 149, 128, 0, 128, True, 512),
    (90, 128, 0, 128, True, 512),
]


def _run_chain(x, x2, w1, b1, res, w2, b2, relu1, relu2, precision):
    m, cin = x.shape
    cout2, N1 = w2.shape[0], w1.shape[0]
    guard = 8192
    arena1 = torch.full((m * N1 + 2 * guard,), 123.0, dtype=x.dtype, device="cuda")
    arena2 = torch.full((m * cout2 + 2 * guard,), 123.0, dtype=x.dtype, device="cuda")
    out1 = arena1[guard:guard + m * N1].view(m, N1)
    out2 = arena2[guard:guard + m * cout2].view(m, cout2)
    rc = lib().semdiff_conv1x1_chain(x.data_ptr(), x2.data_ptr() if x2 is not None else None, w1.data_ptr(), b1.data_ptr(),
                                     res.data_ptr() if res is not None else None, out1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                     out2.data_ptr(), m, cin, x2.shape[1] if x2 is not None else 0, N1, cout2, int(relu1), int(relu2),
                                     _lib.PRECISIONS[precision], sp())
    _lib.check(rc, "semdiff_conv1x1_chain")
    torch.cuda.synchronize()
    for a, n in ((arena1, m * N1), (arena2, m * cout2)):
        assert bool((a[:guard] == 123.0).all()) and bool((a[guard + n:] == 123.0).all()), "kernel wrote outside its output"
    return out1, out2


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("case", CHAIN_CASES)
def test_chain_matches_two_launches(case, precision):
    m, cin, cin2, cout2, has_res = case[:5]
    N1 = case[5] if len(case) > 5 else 256
    dt = DT[precision]
    g = torch.Generator(device="cuda").manual_seed(1234 + m % 1000 + cin + cin2 + cout2)
    rnd = lambda *s: torch.randn(*s, generator=g, device="cuda")
    x = rnd(m, cin).to(dt)
    x2 = rnd(m, cin2).to(dt) if cin2 else None
    w1 = (rnd(N1, cin + cin2) / (cin + cin2) ** 0.5).to(dt)
    b1 = rnd(N1) * 0.1
    res = rnd(m, N1).to(dt) if has_res else None
    w2 = (rnd(cout2, N1) / N1 ** 0.5).to(dt)
    b2 = rnd(cout2) * 0.1
    out1, out2 = _run_chain(x, x2, w1, b1, res, w2, b2, True, True, precision)

    # the same two convs as separate launches of the generic kernel
    xi = x.view(1, m, 1, cin)
    x2i = x2.view(1, m, 1, cin2) if cin2 else None
    y = conv2d(xi, w1[:, :cin].reshape(N1, 1, 1, cin).contiguous(), b1, res.view(1, m, 1, N1) if has_res else None, 1, 0,
               True, precision, _lib.CONV_TC_TMA, x2_nhwc=x2i, w2=w1[:, cin:].contiguous() if cin2 else None)
    t = conv2d(y, w2.reshape(cout2, 1, 1, N1), b2, None, 1, 0, True, precision, _lib.CONV_TC_TMA)
    assert torch.equal(out1, y.view(m, N1)), f"out1 differs from the unfused conv: max {(out1.float() - y.view(m, N1).float()).abs().max().item()}"
    assert torch.equal(out2, t.view(m, cout2)), f"out2 differs from the unfused conv: max {(out2.float() - t.view(m, cout2).float()).abs().max().item()}"

    # and an fp64 reference on the same operands
    xin = x.double() if not cin2 else torch.cat([x.double(), x2.double()], dim=1)
    ref1 = xin @ w1.double().t() + b1.double()
    if has_res:
        ref1 = ref1 + res.double()
    ref1 = torch.relu(ref1)
    rel = 2.0 ** -8 if precision == "bf16" else 2.0 ** -11
    assert not ((out1.double() - ref1).abs() > rel * ref1.abs() + 2e-3).any()
    ref2 = torch.relu(out1.double() @ w2.double().t() + b2.double())   # from the ROUNDED out1, as the reference trunk does
    assert not ((out2.double() - ref2).abs() > rel * ref2.abs() + 2e-3).any()


def test_chain_no_relu_and_errors():
    m = 128 * 20 + 3
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(m, 64, generator=g, device="cuda").bfloat16()
    w1 = (torch.randn(256, 64, generator=g, device="cuda") / 8).bfloat16()
    w2 = (torch.randn(64, 256, generator=g, device="cuda") / 16).bfloat16()
    b1 = torch.randn(256, generator=g, device="cuda")
    b2 = torch.randn(64, generator=g, device="cuda")
    out1, out2 = _run_chain(x, None, w1, b1, None, w2, b2, False, False, "bf16")
    ref1 = x.double() @ w1.double().t() + b1.double()
    assert (ref1 < 0).any() and not ((out1.double() - ref1).abs() > 2.0 ** -8 * ref1.abs() + 2e-3).any()
    ref2 = out1.double() @ w2.double().t() + b2.double()
    assert not ((out2.double() - ref2).abs() > 2.0 ** -8 * ref2.abs() + 2e-3).any()
    # unsupported shapes fail loudly instead of falling back
    bad = torch.empty(m, 96, device="cuda", dtype=torch.bfloat16)
    rc = lib().semdiff_conv1x1_chain(x.data_ptr(), None, w1.data_ptr(), b1.data_ptr(), None, out1.data_ptr(), w2.data_ptr(),
                                     b2.data_ptr(), bad.data_ptr(), m, 64, 0, 256, 96, 1, 1, 0, sp())
    assert rc == -3 and b"conv_chain" in lib().semdiff_last_error()
    rc = lib().semdiff_conv1x1_chain(x.data_ptr(), None, w1.data_ptr(), b1.data_ptr(), None, out1.data_ptr(), w2.data_ptr(),
                                     b2.data_ptr(), out2.data_ptr(), m, 64, 0, 256, 64, 1, 1, 2, sp())   # fp32: SIMT path only
    assert rc == -3
