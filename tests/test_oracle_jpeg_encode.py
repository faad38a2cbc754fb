"""The oracle's JPEG encoder restatement against Pillow's own encoder, byte for byte (CPU)."""
import io

import numpy as np
import pytest
from PIL import Image

from facet_b200.synth import synth_image_bgr
from oracle import jpeg_encode_np as je


@pytest.mark.parametrize("shape", [(427, 640), (640, 427), (16, 16), (17, 33), (100, 150), (8, 8), (1, 1), (31, 47), (426, 640), (49, 65)])
def test_encoder_matches_pillow(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    photo = synth_image_bgr(3, max(h, 2), max(w, 2))[:h, :w, ::-1].copy()
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    flat = np.full((h, w, 3), 255, np.uint8)
    for img in (photo, noise, flat):
        for quality in (80, 95) if h < 200 else (80,):
            buf = io.BytesIO()
            Image.fromarray(img).save(buf, format="JPEG", quality=quality)
            ref = buf.getvalue()
            assert je.encode(img, ref) == ref


def test_encoder_tables_of_the_product_match_the_header():
    from facet_b200.utils import jpeg as fj
    header, packed = fj.encoder_tables(427, 640, 80)
    buf = io.BytesIO()
    Image.new("RGB", (640, 427)).save(buf, format="JPEG", quality=80)
    q, huff, sos_end = je.parse_header(buf.getvalue())
    assert header == buf.getvalue()[:sos_end] and len(packed) == 3328
    p = np.frombuffer(packed, np.uint8)
    assert np.array_equal(p[:256].view(np.uint16), np.concatenate([q[0] * 8, q[1] * 8]).astype(np.uint16))
    code = p[256:2304].view(np.uint16).reshape(4, 256)
    size = p[2304:].reshape(4, 256)
    for t, key in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        co, si = je.huff_codes(*huff[key])
        for sym in co:
            assert code[t, sym] == co[sym] and size[t, sym] == si[sym]
