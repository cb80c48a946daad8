"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the parameter trees keep
the reference's state_dict layout, BatchNorm folding and lowering are correct, and the product refuses to run
without a GPU (no fallback)."""
import os
import re

import pytest
import torch

import semdiff_b200
from oracle.trunks import build_trunk
from semdiff_b200 import _lib, trunks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    with open(os.path.join(ROOT, "include", "semdiff_b200.h")) as f:
        header = f.read()
    declared = set(re.findall(r"\b(semdiff_[a-z0-9_]+)\s*\(", header))
    declared -= {"semdiff_stream_t"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.semdiff_version()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "measuring-semantic-differences-in-the-super-resolution-domain_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn


@pytest.mark.parametrize("name", ["resnet50", "resnet50_clip.openai"])
def test_state_dict_layout_matches_oracle_trunk(name):
    mine, ref = trunks.create_trunk(name), build_trunk(name, calibrate_bn=False)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert mine.state_dict()[k].shape == v.shape, k
    with pytest.raises(RuntimeError):
        mine(torch.zeros(1, 3, 224, 224))   # parameter container only: no PyTorch fallback


def test_fold_conv_bn_matches_torch():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 16, 7, 2, 3, bias=False).double()
    bn = torch.nn.BatchNorm2d(16).double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    x = torch.randn(2, 3, 20, 20, dtype=torch.double)
    w, b = trunks.fold_conv_bn(conv, bn, cin_pad=8)
    assert w.shape == (16, 7, 7, 8) and torch.count_nonzero(w[..., 3:]) == 0
    y = torch.nn.functional.conv2d(x, w[..., :3].permute(0, 3, 1, 2), b, stride=2, padding=3)
    assert torch.allclose(y, bn(conv(x)), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("name,convs,gflop", [("resnet50", 53, 8.174272512), ("resnet50_clip.openai", 55, 10.734452736)])
def test_lowering(name, convs, gflop):
    tree = trunks.create_trunk(name)
    generic = trunks.LOWER[name](tree, 3, False)
    assert generic.input_layout == _lib.INPUT_NHWC8
    assert abs(trunks.conv_flops(generic, 224, 224) / 1e9 - gflop) < 1e-9
    for depth in range(4):
        prog = trunks.LOWER[name](tree, depth)
        assert sum(op.get("n_convs", 0) for op in prog.ops) == convs            # every reference conv is covered
        assert sum(op["kind"] == _lib.OP_CONV for op in prog.ops) == convs - 4  # 4 projection shortcuts are fused
        taps = [op["tap"] for op in prog.ops if op["kind"] == _lib.OP_TAP]
        assert taps == list(range(depth + 1))
        for op in prog.ops:   # a conv never reads or adds the buffer it writes
            if op["kind"] != _lib.OP_TAP:
                assert op["dst"] not in (op["src"], op["res"], op["src2"]) and 0 < op["dst"] < prog.n_bufs
    assert abs(trunks.conv_flops(prog, 224, 224) / 1e9 - gflop) < 1e-9   # SURVEY.md 8d algorithmic FLOPs


def test_no_cpu_fallback_and_argument_checks():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        semdiff_b200.CLIP_lpips_stages_cnn_clsbckb(clip_name="resnet50", depth=3, device="cpu")
    with pytest.raises(ValueError):
        trunks.trunk_family("vit_base_patch16_clip_224.openai")


def test_plan_api_rejects_bad_arguments_without_gpu():
    lib = _lib.load()
    import ctypes as C
    handle = C.c_void_p()
    assert lib.semdiff_plan_create(None, 0, 0, 0, 0, 0, C.byref(handle)) == -1
    assert b"bad arguments" in lib.semdiff_last_error()
    assert lib.semdiff_distance_parts(56 * 56, 256) == 25 and lib.semdiff_distance_parts(49, 2048) == 4
    assert lib.semdiff_distance_parts(1, 8) == 1 and lib.semdiff_distance_parts(1024 * 1024, 256) == 64


def test_s2d_stem_weights_reproduce_the_7x7_conv():
    """The 4x1 conv over the space-to-depth row-window layout (pack_s2d_kernel) equals the 7x7/2 pad-3 conv."""
    torch.manual_seed(0)
    tree = trunks.create_trunk("resnet50")
    with torch.no_grad():
        tree.bn1.running_mean.normal_(); tree.bn1.running_var.uniform_(0.5, 2.0); tree.bn1.bias.normal_()
    prog = trunks.lower_resnet50(tree, 0)
    assert prog.input_layout == _lib.INPUT_S2D_ROW4
    op = prog.ops[0]
    assert (op["kh"], op["kw"], op["cin"], op["stride"], op["pad"], op["alg_k"]) == (4, 1, 64, 1, 0, 147)
    H, W = 12, 16
    x = torch.randn(2, 3, H, W, dtype=torch.double)
    # the layout the pack kernel writes: X2[n, i, q, j*16 + (dy*2+dx)*3 + ci] = x[n, ci, 2(i-2)+dy, 2(q-2+j)+dx]
    X2 = torch.zeros(2, H // 2 + 3, W // 2, 64, dtype=torch.double)
    for i in range(H // 2 + 3):
        for q in range(W // 2):
            for j in range(4):
                for dy in range(2):
                    for dx in range(2):
                        y, xx = 2 * (i - 2) + dy, 2 * (q - 2 + j) + dx
                        if 0 <= y < H and 0 <= xx < W:
                            X2[:, i, q, j * 16 + (dy * 2 + dx) * 3: j * 16 + (dy * 2 + dx) * 3 + 3] = x[:, :, y, xx]
    got = torch.nn.functional.conv2d(X2.permute(0, 3, 1, 2), op["w"].reshape(64, 4, 1, 64).permute(0, 3, 1, 2), op["b"])
    ref = tree.bn1.double().eval()(tree.conv1.double()(x))
    assert got.shape == ref.shape and torch.allclose(got, ref, rtol=1e-10, atol=1e-10)


def test_clip_s2d_stem_weights_reproduce_the_3x3_stride2_conv():
    torch.manual_seed(1)
    tree = trunks.create_trunk("resnet50_clip.openai")
    bn = tree.stem.conv1.bn
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.bias.normal_()
    prog = trunks.lower_clip_resnet50(tree, 0)
    assert prog.input_layout == _lib.INPUT_S2D_ROW2 and prog.head_ops == 4
    op = prog.ops[0]
    assert (op["kh"], op["kw"], op["cin"], op["cout"], op["alg_k"]) == (2, 1, 64, 64, 27)
    H, W = 12, 16
    x = torch.randn(2, 3, H, W, dtype=torch.double)
    X2 = torch.zeros(2, H // 2 + 1, W // 2, 64, dtype=torch.double)
    for i in range(H // 2 + 1):
        for q in range(W // 2):
            for j in range(2):
                for dy in range(2):
                    for dx in range(2):
                        y, xx = 2 * (i - 1) + dy, 2 * (q - 1 + j) + dx
                        if 0 <= y < H and 0 <= xx < W:
                            c = j * 16 + (dy * 2 + dx) * 3
                            X2[:, i, q, c:c + 3] = x[:, :, y, xx]
    got = torch.nn.functional.conv2d(X2.permute(0, 3, 1, 2), op["w"].reshape(64, 2, 1, 64).permute(0, 3, 1, 2), op["b"])
    ref = bn.double().eval()(tree.stem.conv1.conv.double()(x))
    assert got.shape[1] == 64 and torch.allclose(got[:, :32], ref, rtol=1e-10, atol=1e-10)
    assert torch.count_nonzero(got[:, 32:]) == 0
    # the padded 64-channel stem chain: conv2 / conv3 ignore the zero half
    assert prog.ops[1]["cin"] == 64 and prog.ops[1]["cout"] == 64 and prog.ops[2]["cin"] == 64 and prog.ops[2]["cout"] == 64


def test_s2d16_stem_lowering_reproduces_the_7x7_conv():
    """4x4 stride-1 conv (pad 2 before / 1 after) over the plain space-to-depth layout == 7x7 stride-2 pad-3 conv."""
    torch.manual_seed(2)
    tree = trunks.create_trunk("resnet50")
    with torch.no_grad():
        tree.bn1.running_mean.normal_(); tree.bn1.running_var.uniform_(0.5, 2.0); tree.bn1.bias.normal_()
    prog = trunks.lower_resnet50(tree, 0, "s2d16")
    assert prog.input_layout == _lib.INPUT_S2D16
    op = prog.ops[0]
    assert (op["kh"], op["kw"], op["cin"], op["pad"], op["pad_hi"], op["alg_k"]) == (4, 4, 16, 2, 1, 147)
    H, W = 12, 20
    x = torch.randn(2, 3, H, W, dtype=torch.double)
    s2d = torch.zeros(2, 16, H // 2, W // 2, dtype=torch.double)
    for dy in range(2):
        for dx in range(2):
            s2d[:, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = x[:, :, dy::2, dx::2]
    w = op["w"].reshape(64, 4, 4, 16).permute(0, 3, 1, 2)
    got = torch.nn.functional.conv2d(torch.nn.functional.pad(s2d, (2, 1, 2, 1)), w, op["b"])
    ref = tree.bn1.double().eval()(tree.conv1.double()(x))
    assert got.shape == ref.shape and torch.allclose(got, ref, rtol=1e-10, atol=1e-10)
    assert abs(trunks.conv_flops(prog, 224, 224) / 1e9 - 8.174272512) < 1e-9


def test_clip_s2d16_stem_lowering_reproduces_the_3x3_stride2_conv():
    torch.manual_seed(3)
    tree = trunks.create_trunk("resnet50_clip.openai")
    bn = tree.stem.conv1.bn
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.bias.normal_()
    prog = trunks.lower_clip_resnet50(tree, 0, "s2d16")
    op = prog.ops[0]
    assert prog.input_layout == _lib.INPUT_S2D16
    assert (op["kh"], op["kw"], op["cin"], op["cout"], op["pad"], op["pad_hi"], op["alg_k"]) == (2, 2, 16, 32, 1, 0, 27)
    # the 32-channel stem activations stay 32 wide in this variant (strip kernels, 64-byte pixel rows)
    assert [(o["cin"], o["cout"]) for o in prog.ops[1:3]] == [(32, 32), (32, 64)]
    H, W = 12, 20
    x = torch.randn(2, 3, H, W, dtype=torch.double)
    s2d = torch.zeros(2, 16, H // 2, W // 2, dtype=torch.double)
    for dy in range(2):
        for dx in range(2):
            s2d[:, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = x[:, :, dy::2, dx::2]
    w = op["w"].reshape(32, 2, 2, 16).permute(0, 3, 1, 2)
    got = torch.nn.functional.conv2d(torch.nn.functional.pad(s2d, (1, 0, 1, 0)), w, op["b"])
    ref = bn.double().eval()(tree.stem.conv1.conv.double()(x))
    assert got.shape == ref.shape and torch.allclose(got, ref, rtol=1e-10, atol=1e-10)
    assert abs(trunks.conv_flops(prog, 224, 224) / 1e9 - 10.734452736) < 1e-9


def test_split_weight_layout_and_resolution():
    """Split-precision weights (include/semdiff_b200.h): per 64-column K block [64 hi | 64 lo], scaled by a power of two
    into [1024, 2048); hi + lo reproduces the scaled fp64 weight to 2^-22 (fp16) / 2^-16 (bf16) of the layer maximum."""
    torch.manual_seed(1)
    w = torch.randn(96, 192, dtype=torch.float64) * 0.02
    w[3, 5] = 1e-7    # a vanishing weight must not break anything
    for dt, bits in ((torch.float16, 22), (torch.bfloat16, 16)):
        s, scale = trunks.split_weight(w, dt)
        assert s.shape == (96, 384) and s.dtype == dt
        assert 1024 <= float((w * scale).abs().max()) < 2048 and scale == 2.0 ** round(torch.log2(torch.tensor(scale)).item())
        v = s.double().reshape(96, 3, 2, 64)
        assert torch.equal(v[:, :, 0, :].reshape(96, 192), (w * scale).to(dt).double())      # hi halves = the rounded weight
        err = ((v[:, :, 0, :] + v[:, :, 1, :]).reshape(96, 192) - w * scale).abs().max() / (w * scale).abs().max()
        assert float(err) < 2.0 ** -bits


def test_random_init_trunk_is_an_explicit_opt_in(monkeypatch):
    """Without timm there are no pretrained weights: constructing a trunk must raise unless the caller asked for a seeded
    random tree (no silent random-init scores)."""
    monkeypatch.delenv("SEMDIFF_RANDOM_INIT", raising=False)
    with pytest.raises(RuntimeError, match="pretrained"):
        trunks.create_trunk("resnet50")
    with pytest.warns(UserWarning, match="RANDOM"):
        tree = trunks.create_trunk("resnet50", pretrained=False)
    trunks.check_trunk_keys(tree, "resnet50")
    delattr(getattr(tree.layer1, "0"), "conv1")
    with pytest.raises(RuntimeError, match="1 missing"):
        trunks.check_trunk_keys(tree, "resnet50")


@pytest.mark.parametrize("name", ["resnet50", "resnet50_clip.openai"])
def test_unet_lowering_is_a_valid_map_program(name):
    """The local-map program (trunk + five SQDIFF taps + decoder) passes the C-ABI's host-side shape inference: buffer
    shapes chain, the map comes out at the input size, and score / score_map refuse each other's plans."""
    import ctypes as C

    from semdiff_b200.local_eval_models import _decoder
    tree, dec = trunks.create_trunk(name), _decoder()
    prog = trunks.lower_unet(tree, dec, name)
    kinds = [op["kind"] for op in prog.ops]
    assert kinds.count(_lib.OP_SQDIFF) == 5 and kinds.count(_lib.OP_CONCAT) == 4 and kinds.count(_lib.OP_UPSAMPLE2X) == 4
    assert kinds[-1] == _lib.OP_MAP_OUT and _lib.OP_TAP not in kinds
    dec_convs = [op for op in prog.ops if op["kind"] == _lib.OP_CONV][-10:]
    assert [(op["cin"], op["cout"]) for op in dec_convs] == [(2048, 2048), (2048, 2048), (3072, 1024), (1024, 1024), (1536, 512),
                                                               (512, 512), (768, 256), (256, 256), (320, 64), (64, 64)]
    arr = (_lib.SemdiffOp * len(prog.ops))()
    for i, op in enumerate(prog.ops):
        arr[i] = _lib.SemdiffOp(op["kind"], op["src"], op["dst"], op["res"], op["cin"], op["cout"], op["kh"], op["kw"], op["stride"],
                                op["pad"], op["relu"], op["tap"], op["src2"], op["cin2"], op["stride2"], op["pad_hi"], None, None, 1.0, 0)
    lib = _lib.load()
    handle = C.c_void_p()
    for precision in (_lib.BF16, _lib.FP16X3):
        assert lib.semdiff_plan_create(arr, len(prog.ops), prog.n_bufs, precision, prog.input_layout, 0, C.byref(handle)) == 0, lib.semdiff_last_error()
        need = lib.semdiff_workspace_bytes(handle, 4, 224, 224)
        assert need > 0, lib.semdiff_last_error()
        assert lib.semdiff_workspace_bytes(handle, 4, 224, 224) < lib.semdiff_workspace_bytes(handle, 8, 224, 224)
        one = C.c_void_p(16)
        assert lib.semdiff_score(handle, one, one, _lib.FP32, 1, 224, 224, 1, one, one, 0, one, 16, one, None, None, None) == -1
        assert b"semdiff_score_map" in lib.semdiff_last_error()
        lib.semdiff_plan_destroy(handle)
