"""Trunk parameter containers and trunk -> kernel-program lowering.

The reference gets its trunk from timm (`timm.create_model(clip_name, pretrained=True)`,
/root/reference/models/global_eval_models.py:315 and :689).  Here the trunk never runs in PyTorch:
`self.clip` only has to (a) own the parameters under timm's key names so that state_dict()/load_state_dict()
round-trip with reference checkpoints and (b) be walkable by name so it can be lowered to the
op list libsemdiff_b200.so executes (BatchNorm folded into the conv weights in fp64).

If a real `timm` is installed, its model object is used as the container (pretrained weights);
otherwise a structurally identical tree with a seeded random init is built (there is no network here).
"""
from __future__ import annotations

import math
import os
import warnings

import torch
import torch.nn as nn

from . import _lib

_NO_FWD = ("this trunk runs inside libsemdiff_b200.so (hand-written sm_100a kernels); the PyTorch module only "
           "holds its parameters and there is no PyTorch/CPU fallback")


class ParamTree(nn.Module):
    """A named bag of sub-modules/parameters; calling it is an error by design."""

    def forward(self, *args, **kwargs):
        raise RuntimeError(_NO_FWD)


def _seq(mods):
    t = ParamTree()
    for i, m in enumerate(mods):
        t.add_module(str(i), m)
    return t


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=False)


# ---- ImageNet ResNet-50 (timm `resnet50`): keys conv1, bn1, layer{1..4}.{i}.{conv,bn}{1,2,3}, downsample.{0,1}, fc
RESNET50_STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))  # planes, blocks, stride of block 0


def resnet50_tree() -> nn.Module:
    root = ParamTree()
    root.conv1, root.bn1, root.act1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False), nn.BatchNorm2d(64), nn.ReLU()
    root.maxpool = nn.MaxPool2d(3, 2, 1)
    cin = 64
    for li, (planes, blocks, stride) in enumerate(RESNET50_STAGES, start=1):
        blks = []
        for b in range(blocks):
            blk = ParamTree()
            s = stride if b == 0 else 1
            blk.conv1, blk.bn1, blk.act1 = _conv(cin, planes, 1), nn.BatchNorm2d(planes), nn.ReLU()
            blk.conv2, blk.bn2, blk.act2 = _conv(planes, planes, 3, s), nn.BatchNorm2d(planes), nn.ReLU()
            blk.conv3, blk.bn3, blk.act3 = _conv(planes, planes * 4, 1), nn.BatchNorm2d(planes * 4), nn.ReLU()
            if b == 0:
                blk.downsample = _seq([nn.Conv2d(cin, planes * 4, 1, s, bias=False), nn.BatchNorm2d(planes * 4)])
            blks.append(blk)
            cin = planes * 4
        root.add_module(f"layer{li}", _seq(blks))
    root.global_pool = nn.AdaptiveAvgPool2d(1)
    root.fc = nn.Linear(2048, 1000)
    root.pretrained_cfg = {"input_size": (3, 224, 224), "interpolation": "bicubic", "crop_pct": 0.95,
                           "mean": (0.485, 0.456, 0.406), "std": (0.229, 0.224, 0.225)}
    return root


# ---- CLIP ResNet-50 (timm ByobNet `resnet50_clip.openai`): stem.conv{1,2,3}.{conv,bn}, stages.{s}.{b}.*, head.*
def _cna(cin, cout, k, stride=1):
    t = ParamTree()
    t.conv, t.bn = _conv(cin, cout, k, stride), nn.BatchNorm2d(cout)
    return t


def clip_resnet50_tree() -> nn.Module:
    root = ParamTree()
    stem = ParamTree()
    stem.conv1, stem.conv2, stem.conv3 = _cna(3, 32, 3, 2), _cna(32, 32, 3), _cna(32, 64, 3)
    stem.pool = nn.AvgPool2d(2)
    root.stem = stem
    stages, cin = [], 64
    for si, (planes, blocks, _) in enumerate(RESNET50_STAGES):
        blks = []
        for b in range(blocks):
            blk = ParamTree()
            stride = 2 if (b == 0 and si > 0) else 1
            if b == 0:
                sc = ParamTree()
                sc.pool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
                sc.conv = _cna(cin, planes * 4, 1)
                blk.shortcut = sc
            blk.conv1_1x1 = _cna(cin, planes, 1)
            blk.conv2_kxk = _cna(planes, planes, 3)
            blk.conv2_kxk.aa = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
            blk.conv3_1x1 = _cna(planes, planes * 4, 1)
            blk.act = nn.ReLU()
            blks.append(blk)
            cin = planes * 4
        stages.append(_seq(blks))
    root.stages = _seq(stages)
    head = ParamTree()  # attention pool: dead for the score, kept for state_dict parity
    head.pos_embed = nn.Parameter(torch.randn(50, 2048) / 2048 ** 0.5)
    head.q, head.k, head.v, head.proj = nn.Linear(2048, 2048), nn.Linear(2048, 2048), nn.Linear(2048, 2048), nn.Linear(2048, 1024)
    root.head = head
    root.pretrained_cfg = {"input_size": (3, 224, 224), "interpolation": "bicubic", "crop_pct": 1.0,
                           "mean": (0.48145466, 0.4578275, 0.40821073), "std": (0.26862954, 0.26130258, 0.27577711)}
    return root


TREES = {"resnet50": resnet50_tree, "resnet50_clip.openai": clip_resnet50_tree}


def trunk_family(clip_name: str) -> str:
    base = clip_name.split(".")[0]
    if base == "resnet50":
        return "resnet50"
    if base == "resnet50_clip":
        return "resnet50_clip.openai"
    raise ValueError(f"trunk {clip_name!r} is not supported by the B200 scorer (resnet50, resnet50_clip.openai)")


def expected_keys(family: str) -> set:
    """state_dict keys of the trunk container this package lowers (timm's names for the family)."""
    return set(TREES[family]().state_dict().keys())


def check_trunk_keys(clip: nn.Module, family: str):
    """A real timm model is only usable as the parameter container if it has exactly the parameters the lowering walks
    (the key names for resnet50_clip.openai are restated from memory of timm >= 1.0): fail at
    construction, naming the first difference, instead of scoring with a half-lowered trunk."""
    got = {k for k in clip.state_dict().keys() if not k.endswith("num_batches_tracked")}
    want = {k for k in expected_keys(family) if not k.endswith("num_batches_tracked")}
    missing, unexpected = sorted(want - got), sorted(got - want)
    if missing or unexpected:
        raise RuntimeError(
            f"timm trunk {family!r} does not have the parameter names this package lowers: "
            f"{len(missing)} missing (first: {missing[:1]}), {len(unexpected)} unexpected (first: {unexpected[:1]})")


def create_trunk(clip_name: str, seed: int = 0, pretrained: bool | None = None) -> nn.Module:
    """timm.create_model(clip_name, pretrained=True) - the reference's call (:315 / :689) - when a real timm is
    installed.  Without timm there are no pretrained weights to load: that is an error unless the caller opted into a
    seeded random-init trunk (`pretrained=False`, or SEMDIFF_RANDOM_INIT=1 as the tests and the benchmark do), because
    scores from a random trunk are meaningless and must never appear silently."""
    family = trunk_family(clip_name)
    try:
        import timm  # noqa: PLC0415
    except ImportError:
        timm = None
    if timm is not None and getattr(timm, "__version__", None) and pretrained is not False:  # a real install, not a test shim
        clip = timm.create_model(clip_name, pretrained=True)
        check_trunk_keys(clip, family)
        return clip
    if pretrained is None and os.environ.get("SEMDIFF_RANDOM_INIT") != "1":
        raise RuntimeError(
            f"timm is not installed, so the pretrained weights of {clip_name!r} cannot be loaded (the reference calls "
            "timm.create_model(..., pretrained=True)).  Install timm, or opt into a seeded random-init trunk explicitly with "
            "pretrained=False / SEMDIFF_RANDOM_INIT=1 and load a state_dict yourself.")
    warnings.warn(f"semdiff_b200: {clip_name!r} trunk is a seeded RANDOM initialisation (no timm / pretrained=False); "
                  "load a state_dict before trusting any score", stacklevel=3)
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        tree = TREES[family]()
        for m in tree.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    return tree


# ---------------------------------------------------------------------------------------------
# lowering: module tree -> op list with folded weights
# ---------------------------------------------------------------------------------------------

def fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d | None, cin_pad: int | None = None):
    """conv (with or without bias) followed by eval-mode BatchNorm (or by nothing: bn=None)
    -> (weight [Cout,KH,KW,Cin_pad] fp64, bias [Cout] fp64)."""
    w = conv.weight.detach().double().cpu()
    cb = conv.bias.detach().double().cpu() if conv.bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    if bn is None:
        scale, shift = torch.ones(w.shape[0], dtype=torch.float64), cb
    else:
        gamma, beta = bn.weight.detach().double().cpu(), bn.bias.detach().double().cpu()
        mean, var = bn.running_mean.detach().double().cpu(), bn.running_var.detach().double().cpu()
        scale = gamma / torch.sqrt(var + bn.eps)
        shift = beta + (cb - mean) * scale
    w = (w * scale[:, None, None, None]).permute(0, 2, 3, 1).contiguous()
    if cin_pad is not None and cin_pad > w.shape[-1]:
        w = torch.nn.functional.pad(w, (0, cin_pad - w.shape[-1]))
    return w, shift


def split_weight(w: torch.Tensor, dt16: torch.dtype):
    """Folded fp64 weights [Cout, K] (K % 64 == 0) -> ([Cout, 2K] in `dt16`, wscale) for the split precisions
    (include/semdiff_b200.h): w * wscale = hi + lo, stored per 64-column K block as [64 hi | 64 lo].  wscale is the power
    of two that brings the largest weight of the layer into [1024, 2048): for fp16 that keeps the lo halves of all
    but vanishing weights in the normal range (22 significant bits); the conv epilogue multiplies by 1/wscale."""
    cout, k = w.shape
    assert k % 64 == 0, "split precisions need whole 64-column K blocks"
    wmax = float(w.abs().max())
    scale = 2.0 ** (11 - math.frexp(wmax)[1]) if wmax > 0 else 1.0
    ws = w * scale
    hi = ws.to(dt16)
    lo = (ws - hi.double()).to(dt16)
    out = torch.stack([hi.reshape(cout, k // 64, 64), lo.reshape(cout, k // 64, 64)], dim=2)
    return out.reshape(cout, 2 * k).contiguous(), scale


class Program:
    """Ops in execution order; `weights` holds (w fp64, b fp64) per conv op index."""

    def __init__(self):
        self.ops: list[dict] = []
        self.n_bufs = 0
        self.input_layout = _lib.INPUT_NHWC8
        self.head_ops = 0   # leading ops (stem + first pool) the plan may run in L2-sized image chunks
        self.sqdiff_bufs = None   # local-map programs: tap j becomes a SQDIFF op into buffer sqdiff_bufs[j] (see tap())

    def conv(self, conv, bn, src, dst, res=-1, relu=True, cin_pad=None, second=None, cout_pad=None):
        """One fused conv+BN(+ReLU) op.  `second=(conv1x1, bn, src2)` folds a projection shortcut into the same
        launch: its (folded) weights are appended along K and the kernel reads `src2` for those K columns, so
        out = act(conv(src) + conv1x1(src2) + b + b2) never materialises the shortcut tensor."""
        w, b = fold_conv_bn(conv, bn, cin_pad)
        if cout_pad is not None and cout_pad > w.shape[0]:   # zero filters: the extra output channels are exactly 0
            w = torch.cat([w, torch.zeros(cout_pad - w.shape[0], *w.shape[1:], dtype=w.dtype)])
            b = torch.cat([b, torch.zeros(cout_pad - b.shape[0], dtype=b.dtype)])
        pad = conv.padding[0] if not isinstance(conv.padding, str) else {"same": conv.kernel_size[0] // 2, "valid": 0}[conv.padding]
        op = dict(kind=_lib.OP_CONV, src=src, dst=dst, res=res, cin=w.shape[3], cout=w.shape[0],
                  kh=w.shape[1], kw=w.shape[2], stride=conv.stride[0], pad=pad,
                  relu=int(relu), tap=-1, src2=-1, cin2=0, stride2=1, pad_hi=-1,
                  alg_k=conv.in_channels * conv.kernel_size[0] * conv.kernel_size[1], alg_cout=conv.out_channels,
                  n_convs=1)
        w = w.reshape(w.shape[0], -1)
        if second is not None:
            conv2, bn2, src2 = second
            assert conv2.kernel_size == (1, 1) and conv2.padding == (0, 0) and res == -1
            w2, b2 = fold_conv_bn(conv2, bn2)
            w = torch.cat([w, w2.reshape(w2.shape[0], -1)], dim=1)
            b = b + b2
            op.update(src2=src2, cin2=conv2.in_channels, stride2=conv2.stride[0], alg_k=op["alg_k"] + conv2.in_channels,
                      n_convs=2)
        op.update(w=w.contiguous(), b=b)
        self.ops.append(op)

    def stem7_s2d(self, conv, bn, src, dst):
        """7x7 stride-2 pad-3 stem as a 4x1 stride-1 conv over the SEMDIFF_INPUT_S2D_ROW4 buffer (csrc/elementwise.cu):
        tap ky -> (row r = (ky+1)//2, sub-row dy = (ky+1)%2), kx -> (window pixel j = (kx+1)//2, dx = (kx+1)%2),
        channel slot j*16 + (dy*2+dx)*3 + ci.  K = 4*64 = 256 instead of 49*8 = 392 for a channel-padded 7x7."""
        w, b = fold_conv_bn(conv, bn)                      # [Cout, 7, 7, 3]
        w2 = torch.zeros(w.shape[0], 4, 1, 64, dtype=w.dtype)
        for ky in range(7):
            r, dy = (ky + 1) // 2, (ky + 1) % 2
            for kx in range(7):
                j, dx = (kx + 1) // 2, (kx + 1) % 2
                base = j * 16 + (dy * 2 + dx) * 3
                w2[:, r, 0, base:base + 3] = w[:, ky, kx, :]
        self.input_layout = _lib.INPUT_S2D_ROW4
        self.ops.append(dict(kind=_lib.OP_CONV, src=src, dst=dst, res=-1, cin=64, cout=w.shape[0], kh=4, kw=1, stride=1,
                             pad=0, relu=1, tap=-1, src2=-1, cin2=0, stride2=1, pad_hi=-1,
                             w=w2.reshape(w2.shape[0], -1).contiguous(), b=b, alg_k=147, alg_cout=w.shape[0], n_convs=1))

    def stem7_s2d16(self, conv, bn, src, dst):
        """The same stem over the 4x smaller SEMDIFF_INPUT_S2D16 buffer ([2n, H/2, W/2, 16], channel (dy*2+dx)*3+ci):
        a 4x4 stride-1 conv with padding 2 before / 1 after.  Same folded weights, flattened [Cout][r][j][16]; the
        strip kernel (csrc/conv3x3_strip.cu) forms the 16 taps as shifted shared-memory views."""
        self.stem7_s2d(conv, bn, src, dst)     # weight bytes are identical: (r, j*16 + c) == (r, j, c)
        op = self.ops[-1]
        op.update(cin=16, kh=4, kw=4, pad=2, pad_hi=1)
        self.input_layout = _lib.INPUT_S2D16


    def pool(self, kind, src, dst, window):
        self.ops.append(dict(kind=kind, src=src, dst=dst, res=-1, cin=0, cout=0, kh=window, kw=window, stride=window,
                             pad=0, relu=0, tap=-1, src2=-1, cin2=0, stride2=1, pad_hi=-1, w=None, b=None))

    def simple(self, kind, src, dst, src2=-1):
        """SQDIFF / CONCAT / UPSAMPLE2X / MAP_OUT (include/semdiff_b200.h)"""
        self.ops.append(dict(kind=kind, src=src, dst=dst, res=-1, cin=0, cout=0, kh=0, kw=0, stride=0, pad=0, relu=0, tap=-1,
                             src2=src2, cin2=0, stride2=1, pad_hi=-1, w=None, b=None))

    def tap(self, src, j):
        if self.sqdiff_bufs is not None:     # local maps: the decoder consumes (a - b)^2 of the tapped activation
            return self.simple(_lib.OP_SQDIFF, src, self.sqdiff_bufs[j])
        self.ops.append(dict(kind=_lib.OP_TAP, src=src, dst=-1, res=-1, cin=0, cout=0, kh=0, kw=0, stride=0, pad=0,
                             relu=0, tap=j, src2=-1, cin2=0, stride2=1, pad_hi=-1, w=None, b=None))


    def stem3_s2d(self, conv, bn, src, dst, cout_pad=64):
        """3x3 stride-2 pad-1 stem (CLIP) as a 2x1 stride-1 conv over the SEMDIFF_INPUT_S2D_ROW2 buffer:
        ky -> (row r = (ky+1)//2, dy = (ky+1)%2), kx -> (window pixel j = (kx+1)//2, dx = (kx+1)%2)."""
        w, b = fold_conv_bn(conv, bn)                      # [Cout, 3, 3, 3]
        cout = max(w.shape[0], cout_pad)
        w2 = torch.zeros(cout, 2, 1, 64, dtype=w.dtype)
        for ky in range(3):
            r, dy = (ky + 1) // 2, (ky + 1) % 2
            for kx in range(3):
                j, dx = (kx + 1) // 2, (kx + 1) % 2
                base = j * 16 + (dy * 2 + dx) * 3
                w2[:w.shape[0], r, 0, base:base + 3] = w[:, ky, kx, :]
        b2 = torch.zeros(cout, dtype=b.dtype)
        b2[:b.shape[0]] = b
        self.input_layout = _lib.INPUT_S2D_ROW2
        self.ops.append(dict(kind=_lib.OP_CONV, src=src, dst=dst, res=-1, cin=64, cout=cout, kh=2, kw=1, stride=1,
                             pad=0, relu=1, tap=-1, src2=-1, cin2=0, stride2=1, pad_hi=-1, w=w2.reshape(cout, -1).contiguous(),
                             b=b2, alg_k=27, alg_cout=w.shape[0], n_convs=1))


    def stem3_s2d16(self, conv, bn, src, dst, cout_pad=64):
        """The CLIP 3x3/2 stem conv over SEMDIFF_INPUT_S2D16: a 2x2 stride-1 conv, padding 1 before / 0 after."""
        self.stem3_s2d(conv, bn, src, dst, cout_pad)
        op = self.ops[-1]
        w = op["w"].reshape(op["cout"], 2, 64)[:, :, :32]          # slots j = 0, 1 of the row window are the real ones
        op.update(cin=16, kh=2, kw=2, pad=1, pad_hi=0, w=w.reshape(op["cout"], -1).contiguous())
        self.input_layout = _lib.INPUT_S2D16


def lower_resnet50(clip: nn.Module, depth: int, s2d_stem=True, program: Program | None = None, raw_stem_tap: bool = False) -> Program:
    """timm resnet50; taps = layer{s}.2.act3 for s in range(4-depth, 5)  (global_eval_models.py:701).
    s2d_stem: "s2d16" (16-bit modes, stem output width <= 125: strip kernel over the compact space-to-depth input),
    True / "row4" (row-window layout, generic im2col path: any even size, fp32 mode), False (channel-padded 7x7 conv
    through the gather producer: odd image sizes)."""
    P = program or Program()
    IN, A, B, T1, T2 = range(5)
    P.n_bufs = max(P.n_bufs, 5)
    c1 = clip.conv1
    first_tap = 0
    if raw_stem_tap:
        # the local-map U-Net hooks the module `conv1` itself (/root/reference/models/local_eval_models.py:196): the raw
        # 7x7 conv output before bn1 / act1 - one more stem launch with the un-folded weights
        if s2d_stem and c1.kernel_size == (7, 7) and c1.stride == (2, 2):
            P.stem7_s2d(c1, None, IN, T2)
            P.ops[-1]["relu"] = 0
        else:
            P.conv(c1, None, IN, T2, cin_pad=8, relu=False)
        P.tap(T2, 0)
        first_tap = 1
    if s2d_stem and c1.kernel_size == (7, 7) and c1.stride == (2, 2) and c1.padding == (3, 3) and c1.in_channels == 3:
        if s2d_stem == "s2d16" and c1.out_channels == 64:
            P.stem7_s2d16(c1, clip.bn1, IN, T1)
        else:
            P.stem7_s2d(c1, clip.bn1, IN, T1)
    else:
        P.conv(c1, clip.bn1, IN, T1, cin_pad=8)
    P.pool(_lib.OP_MAXPOOL3S2, T1, A, 3)
    P.head_ops = len(P.ops)
    x = A
    for li in range(1, 5):
        layer = getattr(clip, f"layer{li}")
        for bi, blk in enumerate(layer.children()):
            y = B if x == A else A
            P.conv(blk.conv1, blk.bn1, x, T1)
            P.conv(blk.conv2, blk.bn2, T1, T2)
            ds = getattr(blk, "downsample", None)
            if ds is not None:   # projection shortcut folded into conv3's launch (K = planes + inplanes)
                dconv, dbn = list(ds.children())[:2]
                P.conv(blk.conv3, blk.bn3, T2, y, second=(dconv, dbn, x))
            else:
                P.conv(blk.conv3, blk.bn3, T2, y, res=x)
            x = y
            if bi == 2 and li >= 4 - depth:
                P.tap(x, first_tap + li - (4 - depth))
    return P


def lower_clip_resnet50(clip: nn.Module, depth: int, s2d_stem: bool = True, taps: dict | None = None,
                        program: Program | None = None, stem_tap: bool = False) -> Program:
    """timm resnet50_clip.openai; taps = stages.{s}.2.act for s in range(3-depth, 4)  (global_eval_models.py:327),
    or an explicit {(stage, block): tap index} map (CLIP_lpips_wperlay_cnn, :832-833)."""
    if taps is None:
        taps = {(s, 2): s - (3 - depth) + (1 if stem_tap else 0) for s in range(3 - depth, 4)}
    P = program or Program()
    IN, A, B, T1, T2, T3, D0 = range(7)
    P.n_bufs = max(P.n_bufs, 7)
    st = clip.stem
    c1 = st.conv1.conv
    if s2d_stem and c1.kernel_size == (3, 3) and c1.stride == (2, 2) and c1.padding == (1, 1) and c1.in_channels == 3:
        # row-window variant: 32-channel stem activations are carried as 64 channels (upper half exactly zero) so that
        # every stem conv runs on the TMA-im2col tensor-core path (64 channels = one 128-byte swizzle row)
        if s2d_stem == "s2d16":
            # strip kernels all the way (csrc/conv3x3_strip.cu): the 32-channel stem activations stay 32 channels wide
            # (64-byte pixel rows, SWIZZLE_64B) - no zero channels through the tensor pipe or shared memory
            P.stem3_s2d16(c1, st.conv1.bn, IN, T1, cout_pad=c1.out_channels)
            P.conv(st.conv2.conv, st.conv2.bn, T1, T2)
            P.conv(st.conv3.conv, st.conv3.bn, T2, T1)
        else:
            P.stem3_s2d(c1, st.conv1.bn, IN, T1, cout_pad=64)
            P.conv(st.conv2.conv, st.conv2.bn, T1, T2, cin_pad=64, cout_pad=64)
            P.conv(st.conv3.conv, st.conv3.bn, T2, T1, cin_pad=64)
    else:
        P.conv(c1, st.conv1.bn, IN, T1, cin_pad=8)
        P.conv(st.conv2.conv, st.conv2.bn, T1, T2)
        P.conv(st.conv3.conv, st.conv3.bn, T2, T1)
    if stem_tap:   # `stem.conv3` (conv + bn + act, before stem.pool): the local-map U-Net's first tap (local_eval_models.py:27)
        P.tap(T1, 0)
    P.pool(_lib.OP_AVGPOOL, T1, A, 2)
    P.head_ops = len(P.ops) if not stem_tap else 0
    x = A
    for si, stage in enumerate(clip.stages.children()):
        for bi, blk in enumerate(stage.children()):
            y = B if x == A else A
            stride = 2 if (bi == 0 and si > 0) else 1
            P.conv(blk.conv1_1x1.conv, blk.conv1_1x1.bn, x, T1)
            P.conv(blk.conv2_kxk.conv, blk.conv2_kxk.bn, T1, T2)
            mid = T2
            if stride > 1:
                P.pool(_lib.OP_AVGPOOL, T2, T3, stride)
                mid = T3
            sc = getattr(blk, "shortcut", None)
            if sc is not None:   # avg-pool + 1x1 projection shortcut: the 1x1 is folded into conv3's launch
                sx = x
                if stride > 1:
                    P.pool(_lib.OP_AVGPOOL, x, D0, stride)
                    sx = D0
                P.conv(blk.conv3_1x1.conv, blk.conv3_1x1.bn, mid, y, second=(sc.conv.conv, sc.conv.bn, sx))
            else:
                P.conv(blk.conv3_1x1.conv, blk.conv3_1x1.bn, mid, y, res=x)
            x = y
            if (si, bi) in taps:
                P.tap(x, taps[(si, bi)])
    return P


def lower_unet(clip: nn.Module, decoder: nn.ModuleList, family: str, s2d_stem=True) -> Program:
    """The local-map model of /root/reference/models/local_eval_models.py:7-171 (CLIP trunk) / :175-339 (ImageNet trunk):
    trunk with five taps (stem + block 2 of every stage) -> (a - b)^2 per tap (:115) -> U-Net decoder (:117-123: two 3x3
    conv + BN + ReLU per level, bilinear x2 upsampling, channel concat with the next shallower difference) -> sigmoid.
    BatchNorm (eval mode) and the conv biases are folded; the 1-channel head conv is carried with 64 output channels."""
    P = Program()
    n_trunk = 5 if family == "resnet50" else 7
    D = [n_trunk + j for j in range(5)]               # squared differences, shallow -> deep
    U, CAT, X1, X2 = (n_trunk + 5 + j for j in range(4))
    P.n_bufs = n_trunk + 9
    P.sqdiff_bufs = D
    if family == "resnet50":
        lower_resnet50(clip, 3, s2d_stem, program=P, raw_stem_tap=True)
    else:
        lower_clip_resnet50(clip, 3, s2d_stem, program=P, stem_tap=True)
    P.head_ops = 0
    levels = list(decoder.children())
    assert len(levels) == 5
    def two_convs(seq, src):
        m = list(seq.children())
        P.conv(m[0], m[1], src, X1)
        if isinstance(m[4], nn.BatchNorm2d):         # conv, bn, relu, conv, bn, relu
            P.conv(m[3], m[4], X1, X2)
        else:                                         # level 0: conv, bn, relu, conv 64 -> 1 (bias), relu
            P.conv(m[3], None, X1, X2, cout_pad=64)
    two_convs(levels[4], D[4])
    for j in (3, 2, 1, 0):
        P.simple(_lib.OP_UPSAMPLE2X, X2, U)
        P.simple(_lib.OP_CONCAT, D[j], CAT, src2=U)   # torch.concat((diff[-j], bottom_pass), dim=1)  (:121)
        two_convs(levels[j], CAT)
    P.simple(_lib.OP_MAP_OUT, X2, -1)                 # upscaler + final_sigmoid (:123-125)
    return P


LOWER = {"resnet50": lower_resnet50, "resnet50_clip.openai": lower_clip_resnet50}


def conv_flops(program: Program, H: int, W: int) -> int:
    """Algorithmic FLOPs (2*MAC, unpadded Cin) of the conv ops for ONE HxW image."""
    pad_rows = {_lib.INPUT_S2D_ROW4: 3, _lib.INPUT_S2D_ROW2: 1, _lib.INPUT_S2D16: 0}.get(program.input_layout)
    shapes = {0: (H // 2 + pad_rows, W // 2) if pad_rows is not None else (H, W)}
    total = 0
    for op in program.ops:
        h, w = shapes[op["src"]]
        if op["kind"] == _lib.OP_CONV:
            ph = op["pad"] + (op["pad"] if op["pad_hi"] < 0 else op["pad_hi"])
            oh = (h + ph - op["kh"]) // op["stride"] + 1
            ow = (w + ph - op["kw"]) // op["stride"] + 1
            total += 2 * oh * ow * op["alg_cout"] * op["alg_k"]
            shapes[op["dst"]] = (oh, ow)
        elif op["kind"] == _lib.OP_MAXPOOL3S2:
            shapes[op["dst"]] = ((h - 1) // 2 + 1, (w - 1) // 2 + 1)
        elif op["kind"] == _lib.OP_AVGPOOL:
            shapes[op["dst"]] = (h // op["stride"], w // op["stride"])
    return total
