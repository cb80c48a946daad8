"""CPU tests of the preprocessing oracle and of the host-side coefficient tables:
oracle/pil_resize.py == Pillow (bit-exact) == torchvision's eval transform; semdiff_resize_coeffs == the oracle tables."""
import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms as T

from oracle.pil_resize import eval_transform, precompute_coeffs, resize_bicubic_u8
from semdiff_b200 import _lib

SHAPES = [(512, 512, 235, 235), (480, 640, 235, 313), (300, 200, 352, 235), (100, 130, 235, 305), (224, 224, 224, 224)]


@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_resize_is_bit_exact_pillow(shape):
    h, w, oh, ow = shape
    img = (np.random.default_rng(h * w).random((h, w, 3)) * 255).astype(np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(resize_bicubic_u8(img, oh, ow), ref)


@pytest.mark.parametrize("cfg", [(235, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
                                 (224, (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])
def test_oracle_transform_equals_torchvision(cfg):
    resize_to, mean, std = cfg
    img = (np.random.default_rng(3).random((512, 640, 3)) * 255).astype(np.uint8)
    tf = T.Compose([T.Resize(resize_to, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(), T.Normalize(mean, std)])
    assert np.array_equal(tf(Image.fromarray(img)).numpy(), eval_transform(img, resize_to, 224, mean, std))


@pytest.mark.parametrize("sizes", [(512, 235), (640, 313), (200, 235), (1024, 224), (224, 224), (100, 305)])
def test_library_coefficient_tables_equal_oracle(sizes):
    in_size, out_size = sizes
    lib = _lib.load()
    bounds_ref, coeffs_ref, ksize_ref = precompute_coeffs(in_size, out_size)
    ksize = lib.semdiff_resize_ksize(in_size, out_size)
    assert ksize == ksize_ref
    bounds = torch.empty(out_size, 2, dtype=torch.int32)
    coeffs = torch.empty(out_size, ksize, dtype=torch.int32)
    assert lib.semdiff_resize_coeffs(in_size, out_size, bounds.data_ptr(), coeffs.data_ptr()) == 0
    assert np.array_equal(bounds.numpy(), bounds_ref) and np.array_equal(coeffs.numpy(), coeffs_ref)
