"""GPU parity: csrc/thumbnail.cu through the C ABI vs Pillow's Image.thumbnail((640, 640), LANCZOS) — the
pixels of utils/image_transforms.py:32-50 `generate_photo_thumbnail` — bit-exact, and the JPEG bytes."""
from io import BytesIO

import numpy as np
import pytest
from PIL import Image

from facet_b200.synth import synth_image_bgr

pytestmark = pytest.mark.gpu


def _pil_thumbnail(rgb, size=640):
    t = Image.fromarray(rgb).copy()
    t.thumbnail((size, size), Image.Resampling.LANCZOS)
    return t


@pytest.mark.parametrize("shape", [(4000, 6000), (683, 1024), (1000, 1503), (1503, 1000), (2001, 3001), (2500, 323),
                                   (97, 3000), (1279, 1281), (300, 400)])
def test_thumbnail_pixels_bit_exact(shape):
    from facet_b200 import ops
    h, w = shape
    bgr = np.stack([synth_image_bgr(40 + i, h, w) for i in range(2)])
    got = ops.thumbnails(bgr).cpu().numpy()                      # BGR frames -> RGB thumbnails
    for i in range(2):
        ref = np.asarray(_pil_thumbnail(np.ascontiguousarray(bgr[i, :, :, ::-1])))
        assert got[i].shape == ref.shape
        assert np.array_equal(got[i], ref)


def test_noise_frame_and_channel_orders():
    from facet_b200 import ops
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (1, 1777, 2999, 3), dtype=np.uint8)
    ref = np.asarray(_pil_thumbnail(rgb[0]))
    assert np.array_equal(ops.thumbnails(rgb, rgb_order=True).cpu().numpy()[0], ref)
    keep = ops.thumbnails(np.ascontiguousarray(rgb[..., ::-1]), to_rgb=False).cpu().numpy()[0]     # BGR in, BGR out
    assert np.array_equal(keep[..., ::-1], ref)


def test_jpeg_bytes_equal_the_reference_call():
    from facet_b200.utils.image_transforms import generate_photo_thumbnail, generate_photo_thumbnails
    bgr = synth_image_bgr(7, 1200, 1800)
    pil = Image.fromarray(np.ascontiguousarray(bgr[:, :, ::-1]))
    # the reference's function body (utils/image_transforms.py:45-50)
    thumb = pil.copy()
    thumb.thumbnail((640, 640), Image.Resampling.LANCZOS)
    buf = BytesIO()
    thumb.save(buf, format="JPEG", quality=80)
    assert generate_photo_thumbnail(pil) == buf.getvalue()
    assert generate_photo_thumbnails(bgr[None])[0] == buf.getvalue()
