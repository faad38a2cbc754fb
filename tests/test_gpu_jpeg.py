"""GPU JPEG decoding (csrc/jpeg_decode.cu through fb_jpeg_decode) byte-exact against Pillow — what the reference's loader
produces (utils/image_loading.py:90-106) — on 4:4:4 / 4:2:2 / 4:2:0 / grayscale streams, odd sizes, custom Huffman
tables, with and without restart markers, batches with different tables, EXIF orientation, a 24 MP frame, corrupt data."""
import io

import numpy as np
import pytest
from PIL import Image, ImageOps

from facet_b200.synth import synth_frame_int, synth_image_bgr

pytestmark = pytest.mark.gpu


def encode(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, "JPEG", **kw)
    return buf.getvalue()


def pil_rgb(data):
    return np.asarray(ImageOps.exif_transpose(Image.open(io.BytesIO(data))).convert("RGB"))


@pytest.mark.parametrize("shape", [(16, 16), (64, 80), (67, 93), (120, 200), (1, 1), (9, 17), (33, 8), (250, 31), (683, 1024)])
def test_decode_matches_pillow(shape):
    from facet_b200 import ops
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    photo = synth_image_bgr(3, max(h, 2), max(w, 2))[:h, :w, ::-1].copy()
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for kw in ({"quality": 85}, {"quality": 95, "subsampling": 0}, {"quality": 50, "subsampling": 1}, {"quality": 90, "restart_marker_rows": 1},
               {"quality": 75, "restart_marker_blocks": 3, "subsampling": 0}, {"quality": 60, "restart_marker_blocks": 1, "subsampling": 1},
               {"quality": 100, "restart_marker_blocks": 7}, {"quality": 10}, {"quality": 88, "optimize": True, "restart_marker_blocks": 2}):
        # one launch, two streams with different content (and, with optimize, different Huffman tables)
        datas = [encode(photo, **kw), encode(noise, **kw)]
        got = ops.jpeg_decode(datas, bgr=False).cpu().numpy()
        for g, d in zip(got, datas):
            assert np.array_equal(g, pil_rgb(d)), (shape, kw)
    gray = photo[:, :, 0].copy()
    d = encode(gray, quality=80, restart_marker_blocks=4)
    assert np.array_equal(ops.jpeg_decode([d], bgr=True).cpu().numpy()[0], pil_rgb(d)[:, :, ::-1])


def test_bgr_order_orientation_and_errors():
    from facet_b200 import ops
    from facet_b200.utils import jpeg as fj
    rgb = synth_image_bgr(7, 120, 176)[:, :, ::-1].copy()
    for code in (1, 3, 6, 8):
        ex = Image.Exif()
        ex[0x0112] = code
        buf = io.BytesIO()
        Image.fromarray(rgb).save(buf, "JPEG", quality=90, exif=ex, restart_marker_blocks=5)
        want = pil_rgb(buf.getvalue())
        got = ops.jpeg_decode([buf.getvalue()], bgr=True).cpu().numpy()[0]
        assert got.shape == want.shape and np.array_equal(got[:, :, ::-1], want), code
    with pytest.raises(fj.UnsupportedJpeg):
        ops.jpeg_decode([encode(rgb, quality=80, progressive=True)])
    good = encode(rgb, quality=90, restart_marker_blocks=5)
    info = fj.parse(good)
    # a restart marker removed -> the marker count no longer matches the DRI header
    pos = good.index(b"\xff\xd0", info.scan_offset)
    with pytest.raises(RuntimeError, match="restart markers"):
        ops.jpeg_decode([good[:pos] + good[pos + 2:]])
    # streams of different geometry cannot share a launch
    with pytest.raises(ValueError):
        ops.jpeg_decode([good, encode(rgb[:64], quality=90, restart_marker_blocks=5)])


@pytest.mark.parametrize("shape", [(300, 420), (683, 1024), (1200, 1600), (2000, 3008)])
def test_streams_without_restart_markers(shape):
    """No DRI: above 1024 MCUs the self-synchronising scheme runs (unstuff, guessed-state rounds, block prefix, DC integration);
    below, one thread per stream.  Photo-like and pure-noise content, every sampling mode, optimised tables, grayscale."""
    from facet_b200 import ops
    h, w = shape
    rng = np.random.default_rng(h + w)
    photo = synth_image_bgr(9, h, w)[:, :, ::-1].copy()
    noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for kw in ({"quality": 85}, {"quality": 95, "subsampling": 0}, {"quality": 40, "subsampling": 1}, {"quality": 92, "optimize": True}):
        datas = [encode(photo, **kw), encode(noise, **kw)]
        got = ops.jpeg_decode(datas, bgr=False).cpu().numpy()
        for g, d in zip(got, datas):
            assert np.array_equal(g, pil_rgb(d)), (shape, kw)
    d = encode(photo[:, :, 1].copy(), quality=80)
    assert np.array_equal(ops.jpeg_decode([d], bgr=False).cpu().numpy()[0], pil_rgb(d))


def test_full_size_frame_24mp():
    """6000 x 4000 4:2:0 with restart intervals of 25 MCUs (the loader setting the bench uses) against Pillow."""
    from facet_b200 import ops
    rgb = synth_frame_int(1, 4000, 6000)[:, :, ::-1].copy()
    data = encode(rgb, quality=90, restart_marker_blocks=25)
    want = pil_rgb(data)
    got = ops.jpeg_decode([data, data], bgr=True).cpu().numpy()
    assert np.array_equal(got[0][:, :, ::-1], want) and np.array_equal(got[1], got[0])
    # the same frame as a camera would write it: no restart markers (self-synchronising path)
    data = encode(rgb, quality=90)
    got = ops.jpeg_decode([data], bgr=False).cpu().numpy()
    assert np.array_equal(got[0], pil_rgb(data))
