"""ctypes binding of libsemdiff_b200.so (include/semdiff_b200.h).  No torch types cross the boundary:
tensors are passed as data_ptr() integers plus sizes.  There is no CPU fallback: if the library is
missing this raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libsemdiff_b200.so")

BF16, FP16, FP32, FP16X3, BF16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"bf16": BF16, "fp16": FP16, "fp32": FP32, "fp16x3": FP16X3, "bf16x3": BF16X3}
SPLIT = {"fp16x3": "fp16", "bf16x3": "bf16"}   # split precisions -> the 16-bit type of their hi / lo halves
CONV_AUTO, CONV_SIMT, CONV_TC_GATHER, CONV_TC_TMA = 0, 1, 2, 3
OP_CONV, OP_MAXPOOL3S2, OP_AVGPOOL, OP_TAP, OP_SQDIFF, OP_CONCAT, OP_UPSAMPLE2X, OP_MAP_OUT = range(8)
MAX_PARTS = 64
INPUT_NHWC8, INPUT_S2D_ROW4, INPUT_S2D_ROW2, INPUT_S2D16 = 0, 1, 2, 3


class SemdiffOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("src", C.c_int32), ("dst", C.c_int32), ("res", C.c_int32),
                ("cin", C.c_int32), ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("stride", C.c_int32), ("pad", C.c_int32), ("relu", C.c_int32), ("tap", C.c_int32),
                ("src2", C.c_int32), ("cin2", C.c_int32), ("stride2", C.c_int32), ("pad_hi", C.c_int32),
                ("weight", C.c_void_p), ("bias", C.c_void_p), ("wscale", C.c_float), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/semdiff_b200.h declares
_P, _I, _L = C.c_void_p, C.c_int32, C.c_int64
SIGNATURES = {
    "semdiff_plan_create": (_I, [C.POINTER(SemdiffOp), _I, _I, _I, _I, _I, C.POINTER(_P)]),
    "semdiff_plan_destroy": (_I, [_P]),
    "semdiff_plan_set_conv_impl": (_I, [_P, _I]),
    "semdiff_workspace_bytes": (_L, [_P, _I, _I, _I]),
    "semdiff_score": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _L, _P, _P, _P, _P]),
    "semdiff_score_map": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _L, _P, _P]),
    "semdiff_decoder_op": (_I, [_I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "semdiff_plan_set_profiling": (_I, [_P, _I]),
    "semdiff_plan_get_profile": (_I, [_P, C.POINTER(C.c_float), C.POINTER(_I), _I, _I]),
    "semdiff_plan_last_launches": (_L, [_P]),
    "semdiff_pack_input": (_I, [_P, _P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "semdiff_conv2d": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "semdiff_conv2d_maxpool": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "semdiff_conv2d_avgpool": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "semdiff_conv1x1_chain": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _I, _I, _I, _I, _I, _I, _I, _P]),
    "semdiff_maxpool3x3s2": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "semdiff_avgpool": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "semdiff_distance_parts": (_I, [_I, _I]),
    "semdiff_layer_distance": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _I, _I, _P]),
    "semdiff_head": (_I, [_P, _I, _I, C.POINTER(_I), C.POINTER(_I), _P, _P, _P, _P]),
    "semdiff_resize_ksize": (_I, [_I, _I]),
    "semdiff_resize_coeffs": (_I, [_I, _I, _P, _P]),
    "semdiff_preprocess_u8": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, C.POINTER(C.c_float),
                                   C.POINTER(C.c_float), _P, _P, _I, _P]),
    "semdiff_last_error": (C.c_char_p, []),
    "semdiff_version": (C.c_char_p, []),
}

_lib = None


class SemdiffError(RuntimeError):
    pass


def load():
    """Load the shared library (building nothing: run __graft_entry__.build() / build.py first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SemdiffError(
            f"{LIB_PATH} is missing. This package has no CPU or PyTorch fallback: build the CUDA library with "
            f"`python __graft_entry__.py build` (needs nvcc) before use.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc < 0:
        raise SemdiffError(f"{what} failed ({rc}): {load().semdiff_last_error().decode()}")
    return rc


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
