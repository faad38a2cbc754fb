"""CPU: the NumPy restatement of Pillow's Image.thumbnail((640, 640), LANCZOS) (facet_b200/utils/thumbnail.py,
the tables the CUDA kernels consume) against the installed Pillow itself — the call
utils/image_transforms.py:46-47 of the reference makes."""
import numpy as np
import pytest
from PIL import Image

from facet_b200.utils import thumbnail as th

SHAPES = [(683, 1024), (1000, 1503), (1503, 1000), (300, 400), (641, 700), (2001, 3001), (2500, 323), (97, 3000),
          (1280, 1280), (1279, 1281), (640, 640), (1, 2000)]


def _pil_thumbnail(rgb, size=640):
    t = Image.fromarray(rgb).copy()
    t.thumbnail((size, size), Image.Resampling.LANCZOS)
    return np.asarray(t)


@pytest.mark.parametrize("shape", SHAPES)
def test_numpy_restatement_matches_pillow(shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1])
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    img[::5] //= 3                                     # some structure besides noise
    ref = _pil_thumbnail(img)
    got = th.thumbnail_numpy(img)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_sizes_and_factors():
    assert th.thumbnail_size(4000, 6000) == (427, 640)
    assert th.thumbnail_size(6000, 4000) == (640, 427)
    assert th.thumbnail_size(300, 400) is None
    p = th.plan(4000, 6000)
    assert (p.fx, p.fy, p.red_h, p.red_w) == (4, 4, 1000, 1500)
    assert th.reduce_multiplier(16) == 1 << 20
