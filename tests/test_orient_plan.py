"""The (swap, flip_x, flip_y) table behind fb_orient against PIL's own ImageOps.exif_transpose (CPU)."""
import numpy as np
import pytest
from PIL import Image, ImageOps

from facet_b200.utils.image_loading import EXIF_ORIENTATION_TAG, METHODS, exif_orientation


def apply_method(img, swap, flip_x, flip_y):
    """out(x', y') = in(sx, sy) exactly as csrc/orient.cu indexes it."""
    h, w = img.shape[:2]
    oh, ow = (w, h) if swap else (h, w)
    yy, xx = np.mgrid[0:oh, 0:ow]
    u, v = (yy, xx) if swap else (xx, yy)
    sx = (w - 1 - u) if flip_x else u
    sy = (h - 1 - v) if flip_y else v
    return img[sy, sx]


@pytest.mark.parametrize("code", range(1, 9))
def test_method_table_matches_exif_transpose(code):
    rng = np.random.default_rng(code)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    pil = Image.fromarray(img)
    exif = pil.getexif()
    exif[EXIF_ORIENTATION_TAG] = code
    pil.info["exif"] = exif.tobytes()
    assert exif_orientation(pil) == code
    want = np.asarray(ImageOps.exif_transpose(pil))
    got = apply_method(img, *METHODS[code])
    assert got.shape == want.shape and np.array_equal(got, want)


def test_missing_or_invalid_orientation_is_identity():
    pil = Image.fromarray(np.zeros((4, 5, 3), np.uint8))
    assert exif_orientation(pil) == 1
    exif = pil.getexif()
    exif[EXIF_ORIENTATION_TAG] = 9
    assert exif_orientation(pil) == 1
