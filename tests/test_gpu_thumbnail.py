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


@pytest.mark.parametrize("shape", [(4000, 6000), (3001, 4504), (2667, 4000), (515, 1032), (402, 600), (403, 1001)])
def test_box_reduction_rides_the_technical_pass(shape):
    """fb_tech_stats_fused: the (4, 4) box reduction (and the luma plane) written by the technical pass itself equal
    Pillow's `reduce(4)`; the statistics are those of the plain pass; the thumbnail finished from the reduced plane
    equals Pillow's `thumbnail((640, 640), LANCZOS)`.  Widths that are not multiples of 8 take the generic kernel and
    the separate reduction pass behind the same call."""
    import torch
    from PIL import Image
    from facet_b200 import ops
    from facet_b200.synth import synth_image_bgr
    h, w = shape
    frames = np.stack([synth_image_bgr(i, h, w) for i in (4, 6)])
    t = torch.from_numpy(frames).cuda()
    for rgb_order in (False, True):
        box = torch.empty((2, (h + 3) // 4, (w + 3) // 4, 3), dtype=torch.uint8, device="cuda")
        luma = torch.empty((2, h, w), dtype=torch.uint8, device="cuda")
        plain = ops.tech_stats_raw(t, rgb_order=rgb_order)
        fused = ops.tech_stats_raw(t, rgb_order=rgb_order, box_out=box)
        both = ops.tech_stats_raw(t, rgb_order=rgb_order, luma_out=luma, box_out=box.clone().zero_()) if w % 8 == 0 else None
        for a, b in zip(plain[:3], fused[:3]):
            assert torch.equal(a, b)
        if both is not None:
            for a, b in zip(plain[:3], both[:3]):
                assert torch.equal(a, b)
        got = box.cpu().numpy()
        for i in range(2):
            want = np.asarray(Image.fromarray(frames[i]).reduce(4))
            assert np.array_equal(got[i], want), (shape, rgb_order, int((got[i] != want).sum()))
        if ops.thumbnail_reduces_by_4(h, w):
            th = ops.thumbnails(t, rgb_order=rgb_order, to_rgb=True, reduced=box).cpu().numpy()
            assert np.array_equal(th, ops.thumbnails(t, rgb_order=rgb_order, to_rgb=True).cpu().numpy())
            for i in range(2):
                pil = Image.fromarray(frames[i] if rgb_order else frames[i][:, :, ::-1].copy())
                pil.thumbnail((640, 640), Image.Resampling.LANCZOS)
                assert np.array_equal(th[i], np.asarray(pil))
