"""JPEG encoding on the device (csrc/jpeg_encode.cu) against Pillow's encoder, byte for byte: the thumbnails the reference stores
(utils/image_transforms.py:32-50, quality 80) and other sizes / qualities, incl. sizes with dummy blocks and replicated edges."""
import io

import numpy as np
import pytest
from PIL import Image

from facet_b200.synth import synth_image_bgr

pytestmark = pytest.mark.gpu


def _pillow(img, quality):
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=quality)
    return buf.getvalue()


@pytest.mark.parametrize("shape", [(427, 640), (640, 427), (480, 640), (16, 16), (17, 33), (100, 150), (8, 8), (1, 1), (31, 47), (426, 640),
                                   (49, 65), (640, 640), (360, 640)])
def test_streams_equal_pillow(shape):
    from facet_b200 import ops
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    imgs = np.stack([synth_image_bgr(3, max(h, 2), max(w, 2))[:h, :w, ::-1], synth_image_bgr(7, max(h, 2), max(w, 2))[:h, :w, ::-1],
                     rng.integers(0, 256, (h, w, 3), dtype=np.uint8), np.full((h, w, 3), 255, np.uint8), np.zeros((h, w, 3), np.uint8)])
    for quality in (80, 95, 30):
        got = ops.jpeg_encode(imgs, quality=quality)
        for i in range(len(imgs)):
            want = _pillow(imgs[i], quality)
            assert got[i] == want, (shape, quality, i, len(got[i]), len(want),
                                    next((k for k, (a, b) in enumerate(zip(got[i], want)) if a != b), None))


def test_thumbnail_jpeg_of_a_24mp_frame_and_timing():
    """generate_photo_thumbnails: thumbnail pixels AND their JPEG stream from the device equal the reference's calls."""
    import torch
    from facet_b200 import ops
    from facet_b200.utils.image_transforms import generate_photo_thumbnails
    frames = np.stack([synth_image_bgr(i, 4000, 6000) for i in (4, 6)])
    got = generate_photo_thumbnails(frames)
    for i in range(2):
        thumb = Image.fromarray(frames[i][:, :, ::-1].copy())
        thumb.thumbnail((640, 640), Image.Resampling.LANCZOS)
        assert got[i] == _pillow(np.asarray(thumb), 80)
    px = ops.thumbnails(torch.from_numpy(frames).cuda().repeat(16, 1, 1, 1))
    ops.jpeg_encode(px, as_device=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.jpeg_encode(px, as_device=True)
    e1.record()
    torch.cuda.synchronize()
    print("JPEG encode of 32 thumbnails (640x427) on the device: %.3f ms" % (e0.elapsed_time(e1) / 5))
