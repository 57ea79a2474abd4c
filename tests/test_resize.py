"""PIL-exact bicubic Resize(224)+CenterCrop(224): the CPU restatement is pinned bit-exactly
against Pillow itself (the reference's own dependency, present in this image); the CUDA kernel is
pinned against both on the GPU box."""
import numpy as np
import pytest

from oracle import pil_resize_ref as R

SIZES = [(320, 240), (240, 320), (640, 480), (500, 300), (225, 224), (224, 400), (1024, 768), (100, 80),
         (60, 200), (333, 777), (224, 224), (448, 448), (223, 300), (2000, 1500), (224, 225), (7, 9)]


def _pil(arr):
    from PIL import Image
    h, w = arr.shape[:2]
    nw, nh = R.resize_size(w, h)
    im = Image.fromarray(arr)
    if (nw, nh) != (w, h):
        im = im.resize((nw, nh), Image.BICUBIC)
    left, top = R.crop_origin(nw, nh)
    return np.asarray(im.crop((left, top, left + 224, top + 224)))


@pytest.mark.parametrize("w,h", SIZES)
def test_restatement_is_bit_exact_with_pillow(w, h):
    rng = np.random.default_rng(w * 10007 + h)
    arr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(R.resize_center_crop(arr), _pil(arr))


def test_product_transform_uses_the_same_steps():
    from PIL import Image
    from clipb200 import clip
    rng = np.random.default_rng(5)
    arr = rng.integers(0, 256, (300, 500, 3), dtype=np.uint8)
    assert np.array_equal(clip.image_to_u8(Image.fromarray(arr)), _pil(arr))


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", SIZES + [(4000, 3000)])
def test_cuda_resize_is_bit_exact_with_pillow(w, h):
    import torch
    from clipb200 import clip
    rng = np.random.default_rng(w * 31 + h)
    arr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = clip.resize_center_crop_device(torch.from_numpy(arr).cuda()).cpu().numpy()
    assert np.array_equal(got, _pil(arr)), f"{(got.astype(int) - _pil(arr)).__abs__().max()}"
