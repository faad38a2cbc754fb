"""JPEG decoding, CPU side: (1) the oracle (oracle/jpeg_np.py) against Pillow's decoder — what the reference gets from
`Image.open(...).convert('RGB')`; (2) the header parser / table packer of the product (facet_b200/utils/jpeg.py) together
with the __host__ __device__ bodies of csrc/jpeg_decode.cu (entropy decoding, inverse DCT, upsampling + colour conversion),
compiled into a host test tool (tests/tools/jpeg_host_check.cu), against Pillow.  The kernels proper run in
tests/test_gpu_jpeg.py."""
import io
import os
import shutil
import subprocess

import numpy as np
import pytest
from PIL import Image

from facet_b200.synth import synth_image_bgr
from facet_b200.utils import jpeg as fj

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def encode(arr, **kw):
    buf = io.BytesIO()
    Image.fromarray(arr).save(buf, "JPEG", **kw)
    return buf.getvalue()


def fixtures():
    rng = np.random.default_rng(0)
    out = []
    for (h, w) in [(16, 16), (64, 80), (67, 93), (120, 200), (1, 1), (9, 17), (33, 8), (250, 31)]:
        photo = synth_image_bgr(3, max(h, 2), max(w, 2))[:h, :w, ::-1].copy()
        noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for arr in (photo, noise):
            for kw in ({"quality": 85}, {"quality": 95, "subsampling": 0}, {"quality": 50, "subsampling": 1},
                       {"quality": 90, "restart_marker_rows": 1}, {"quality": 75, "restart_marker_blocks": 3, "subsampling": 0},
                       {"quality": 60, "restart_marker_blocks": 1, "subsampling": 1}, {"quality": 100}, {"quality": 10},
                       {"quality": 88, "optimize": True, "restart_marker_blocks": 2}):
                out.append((arr, kw))
    gray = synth_image_bgr(5, 50, 70)[:, :, 0].copy()
    out.append((gray, {"quality": 80}))
    out.append((gray, {"quality": 92, "restart_marker_blocks": 4}))
    return out


def pil_rgb(data):
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def test_oracle_matches_pillow():
    from oracle import jpeg_np
    for arr, kw in fixtures():
        data = encode(arr, **kw)
        want = pil_rgb(data)
        got = jpeg_np.decode_rgb(data)
        assert got.shape == want.shape and np.array_equal(got, want), (arr.shape, kw)


def test_parser_matches_oracle_header_and_rejects_unsupported():
    from oracle import jpeg_np
    for arr, kw in fixtures()[:40]:
        data = encode(arr, **kw)
        a, b = fj.parse(data), jpeg_np.parse_header(data)
        assert (a.width, a.height, a.ncomp, a.restart_interval, a.scan_offset, a.scan_end) == \
               (b.width, b.height, len(b.components), b.restart_interval, b.scan_offset, b.scan_end)
        if a.ncomp == 3:
            assert a.hs == [c["h"] for c in b.components] and a.td == [c["td"] for c in b.components]
        assert len(fj.pack_tables(a)) == fj.TABLESET_BYTES
    rgb = synth_image_bgr(1, 40, 40)
    with pytest.raises(fj.UnsupportedJpeg):
        fj.parse(encode(rgb, quality=80, progressive=True))
    with pytest.raises(fj.UnsupportedJpeg):
        fj.parse(b"\x89PNG....")
    buf = io.BytesIO()
    Image.fromarray(rgb).convert("CMYK").save(buf, "JPEG")
    with pytest.raises(fj.UnsupportedJpeg):
        fj.parse(buf.getvalue())
    # EXIF orientation is read from APP1 (utils/image_loading.py:101 applies ImageOps.exif_transpose)
    ex = Image.Exif()
    ex[0x0112] = 6
    buf = io.BytesIO()
    Image.fromarray(rgb).save(buf, "JPEG", exif=ex)
    assert fj.parse(buf.getvalue()).orientation == 6 and fj.parse(encode(rgb)).orientation == 1


@pytest.fixture(scope="module")
def host_tool():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out_dir = os.path.join(HERE, "tools", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "jpeg_host_check")
    src = os.path.join(HERE, "tools", "jpeg_host_check.cu")
    dep = os.path.join(ROOT, "facet_b200", "csrc", "jpeg_decode.cu")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-w", "-I", os.path.join(ROOT, "facet_b200", "csrc"),
                               "-o", exe, src])
    return exe


def run_host_tool(exe, data, tmp_path, bgr=0, selfsync=0):
    info = fj.parse(data)
    pad = lambda v: (list(v) + [1, 1, 1])[:3] if v is info.hs or v is info.vs else (list(v) + [0, 0, 0])[:3]
    hdr = np.zeros(32, np.int32)
    hdr[0:3] = (info.width, info.height, info.ncomp)
    hdr[3:6], hdr[6:9], hdr[9:12], hdr[12:15] = pad(info.hs), pad(info.vs), pad(info.tq), pad(info.td)
    hdr[15], hdr[16], hdr[17] = info.restart_interval, bgr, info.scan_end - info.scan_offset
    hdr[18:21] = pad(info.ta)
    hdr[21] = selfsync
    req, out = tmp_path / "req.bin", tmp_path / "out.bin"
    req.write_bytes(hdr.tobytes() + fj.pack_tables(info) + data[info.scan_offset:info.scan_end])
    subprocess.check_call([exe, str(req), str(out)])
    return np.frombuffer(out.read_bytes(), np.uint8).reshape(info.height, info.width, 3)


def test_device_function_bodies_match_pillow_on_the_host(host_tool, tmp_path):
    for arr, kw in fixtures():
        data = encode(arr, **kw)
        want = pil_rgb(data)
        got = run_host_tool(host_tool, data, tmp_path)
        assert np.array_equal(got, want), (arr.shape, kw, int(np.abs(got.astype(int) - want).max()))
    data = encode(synth_image_bgr(2, 97, 131)[:, :, ::-1].copy(), quality=90, restart_marker_blocks=5)
    assert np.array_equal(run_host_tool(host_tool, data, tmp_path, bgr=1), pil_rgb(data)[:, :, ::-1])


def test_self_synchronising_decode_matches_pillow_on_the_host(host_tool, tmp_path):
    """Streams WITHOUT restart markers through the self-synchronising scheme (subsequences decoded from guessed states, rounds
    until the boundary states are stable, block-index prefix sum, DC differences integrated afterwards), executed sequentially
    by the host tool with the kernels' own functions."""
    rng = np.random.default_rng(5)
    for (h, w) in [(64, 80), (250, 331), (683, 1024), (1200, 1600)]:
        photo = synth_image_bgr(4, h, w)[:, :, ::-1].copy()
        noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for arr in (photo, noise):
            for kw in ({"quality": 85}, {"quality": 95, "subsampling": 0}, {"quality": 40, "subsampling": 1}, {"quality": 92, "optimize": True}):
                data = encode(arr, **kw)
                assert fj.parse(data).restart_interval == 0
                got = run_host_tool(host_tool, data, tmp_path, selfsync=1)
                assert np.array_equal(got, pil_rgb(data)), (arr.shape, kw)
    gray = synth_image_bgr(5, 500, 700)[:, :, 0].copy()
    data = encode(gray, quality=80)
    assert np.array_equal(run_host_tool(host_tool, data, tmp_path, selfsync=1), pil_rgb(data))
