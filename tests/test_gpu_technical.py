"""GPU parity: csrc/tech_stats.cu through the C ABI vs the NumPy oracle and the reference goldens."""
import numpy as np
import pytest

from conftest import approx_rel
from facet_b200.synth import synth_image_bgr

pytestmark = pytest.mark.gpu

INT_KEYS = ("shadow_clipped", "highlight_clipped", "is_silhouette", "is_monochrome")


def _check_stats(st, ref):
    assert st.hist256.tolist() == ref["hist256"].tolist(), "hist256 differs"
    assert np.array_equal(st.hs_hist.astype(np.int64), ref["hs_hist"]), "hs_hist differs"
    assert st.sum_lap == ref["sum_lap"]
    assert st.sum_lap_sq == ref["sum_lap_sq"]
    assert st.sum_abs_noise == ref["sum_abs_noise"]


@pytest.mark.parametrize("shape", [(64, 64), (128, 192), (96, 512), (50, 1040), (683, 1024), (300, 1552),
                                   (2, 8), (3, 16), (5, 264), (130, 8), (257, 776)])
def test_fast_kernel_bit_exact_vs_oracle(shape):
    from facet_b200 import ops
    from oracle import technical_np as onp
    h, w = shape
    imgs = np.stack([synth_image_bgr(i, h, w) for i in range(8)])
    stats = ops.tech_stats(imgs, want_hs=True)
    for i, st in enumerate(stats):
        ref = onp.tech_stats(imgs[i])
        _check_stats(st, ref)
        assert st.sum_saturation == float(np.dot(np.arange(256), ref["hs_hist"].sum(axis=0)))


@pytest.mark.parametrize("shape", [(2, 2), (3, 517), (260, 5), (97, 131), (1024, 683), (33, 48)])
def test_generic_kernel_bit_exact_vs_oracle(shape):
    from facet_b200 import ops
    from oracle import technical_np as onp
    h, w = shape
    imgs = np.stack([synth_image_bgr(20 + i, h, w) for i in range(3)])
    for st, img in zip(ops.tech_stats(imgs, want_hs=True), imgs):
        _check_stats(st, onp.tech_stats(img))


def test_fast_and_generic_agree_and_rgb_order():
    from facet_b200 import ops
    imgs = np.stack([synth_image_bgr(i, 200, 640) for i in range(4)])
    a = ops.tech_stats(imgs, want_hs=True)
    b = ops.tech_stats(imgs, want_hs=True, force_generic=True)
    c = ops.tech_stats(np.ascontiguousarray(imgs[..., ::-1]), rgb_order=True, want_hs=True)
    for x, y, z in zip(a, b, c):
        for other in (y, z):
            assert x.hist256.tolist() == other.hist256.tolist()
            assert np.array_equal(x.hs_hist, other.hs_hist)
            assert (x.sum_lap, x.sum_lap_sq, x.sum_abs_noise) == (other.sum_lap, other.sum_lap_sq, other.sum_abs_noise)


def test_all_colours_hsv_gray_exhaustive():
    """All 2^24 colours through the kernel: histograms must equal the oracle's (cv2-pinned) ones."""
    from facet_b200 import ops
    from oracle import technical_np as onp
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v & 255), (v >> 8) & 255, (v >> 16) & 255], axis=-1).astype(np.uint8).reshape(4096, 4096, 3)
    st = ops.tech_stats(img, want_hs=True)[0]
    ref = onp.tech_stats(img)
    _check_stats(st, ref)


def test_metric_dicts_match_reference_golden(technical_golden):
    """Drop-in surface: ImageCache + TechnicalAnalyzer dicts vs the reference's own outputs."""
    from facet_b200.analyzers import ImageCache, TechnicalAnalyzer as TA
    for rec in technical_golden["cases"]:
        img = synth_image_bgr(rec["index"], rec["height"], rec["width"])
        cache = ImageCache(img)
        assert cache.height == rec["height"] and cache.width == rec["width"]
        assert cache.stats.hist256.tolist() == rec["hist256"]
        assert approx_rel(cache.laplacian_variance, rec["laplacian_variance"], rel=1e-9, abs_=1e-9)
        got = {
            "sharpness": TA.get_sharpness_data(img, cache=cache),
            "color": TA.get_color_harmony_data(img, cache=cache),
            "histogram": TA.get_histogram_data(img, cache=cache),
            "monochrome": TA.detect_monochrome(img, threshold=0.10, cache=cache),
            "dynamic_range": TA.get_dynamic_range(img, cache=cache),
            "noise": TA.get_noise_estimate(img, cache=cache),
            "contrast": TA.get_contrast_score(img, cache=cache),
        }
        for name, want in ((k, rec[k]) for k in got):
            for key, w in want.items():
                g = got[name][key]
                if key == "histogram_bytes":
                    assert g.hex() == w, "histogram_bytes must be bit-exact"
                elif key in INT_KEYS:
                    assert int(g) == int(w), (name, key)
                else:
                    # north_star: float metrics within 1e-3 relative (abs floor = half a rounding step)
                    assert approx_rel(float(g), float(w), rel=1e-3, abs_=5e-5), (rec["index"], name, key, g, w)


def test_none_image_defaults_and_cache_planes():
    from facet_b200.analyzers import ImageCache, TechnicalAnalyzer as TA
    from oracle import technical_np as onp
    assert TA.get_sharpness_data(None) == {"raw_variance": 0, "normalized": 0}
    assert TA.get_histogram_data(None)["exposure_score"] == 5.0
    assert TA.get_noise_estimate(None) == {"noise_sigma": 0}
    img = synth_image_bgr(5, 77, 130)
    cache = ImageCache(img)
    assert np.array_equal(cache.gray, onp.gray_u8(img))
    h, s, v = onp.hsv_u8(img)
    assert np.array_equal(cache.hsv, np.stack([h, s, v], axis=-1).astype(np.uint8))


def test_host_buffer_entry_point_and_roi():
    from facet_b200 import ops
    from oracle import technical_np as onp
    imgs = np.stack([synth_image_bgr(i, 120, 320) for i in range(3)])
    for st, img in zip(ops.tech_stats_host(imgs, want_hs=True), imgs):
        _check_stats(st, onp.tech_stats(img))
    boxes = [(0, 0, 320, 120), (10, 5, 100, 60), (300, 100, 320, 120), (7, 7, 9, 9)]
    got = ops.roi_laplacian(imgs[1], boxes)
    for row, box in zip(got, boxes):
        assert tuple(int(x) for x in row) == onp.roi_laplacian_sums(imgs[1], box)


def test_full_size_properties_24mp():
    """BASELINE config-2 size: properties that do not need the (slow) oracle at 24 MP —
    histogram mass, linearity over a vertical split, and agreement of two row decompositions."""
    import torch
    from facet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    img = torch.randint(0, 256, (1, 4000, 6000, 3), dtype=torch.uint8, device="cuda", generator=g)
    st = ops.tech_stats(img, want_hs=True)[0]
    assert int(st.hist256.sum()) == 24_000_000
    assert int(st.hs_hist.astype(np.int64).sum()) == 24_000_000
    # top/bottom halves: histograms add up exactly (stencil sums do not: borders differ)
    top = ops.tech_stats(img[:, :2000].contiguous(), want_hs=True)[0]
    bot = ops.tech_stats(img[:, 2000:].contiguous(), want_hs=True)[0]
    assert np.array_equal(top.hist256 + bot.hist256, st.hist256)
    assert np.array_equal(top.hs_hist.astype(np.int64) + bot.hs_hist.astype(np.int64), st.hs_hist.astype(np.int64))
    # a batch of 3 copies exercises a different unit decomposition: results must not change
    rep = ops.tech_stats(img.expand(3, -1, -1, -1).contiguous(), want_hs=False)
    for r in rep:
        assert (r.sum_lap, r.sum_lap_sq, r.sum_abs_noise) == (st.sum_lap, st.sum_lap_sq, st.sum_abs_noise)
        assert np.array_equal(r.hist256, st.hist256)
    # oracle on a 1/16 crop of the same frame
    from oracle import technical_np as onp
    crop = img[0, 1000:2000, 1504:3008].contiguous()
    cs = ops.tech_stats(crop, want_hs=True)[0]
    _check_stats(cs, onp.tech_stats(crop.cpu().numpy()))


def test_crop_sharpness_matches_cv2_on_the_crop():
    """FaceAnalyzer._get_crop_sharpness (analyzers/face.py:272-279): the reference's own cv2 calls on the cropped
    array, including boxes that stick out of the frame and empty ones; isolation bonus of batch_processor.py:254-260."""
    import cv2
    from facet_b200.analyzers.face import crop_sharpness, crop_sharpness_batch, isolation_bonus, mean_face_sharpness
    from facet_b200.analyzers import ImageCache
    img = synth_image_bgr(5, 240, 360)
    boxes = [(20, 30, 140, 200), (-15, -4, 60, 50), (300, 200, 500, 400), (0, 0, 360, 240), (50, 50, 50, 90), (400, 10, 420, 30),
             (100, 100, 101, 103), (10, 10, 12, 12)]

    def ref(b):
        h, w = img.shape[:2]
        y1, y2, x1, x2 = max(0, b[1]), min(h, b[3]), max(0, b[0]), min(w, b[2])
        crop = img[y1:y2, x1:x2]
        if crop.size == 0:
            return 0
        return cv2.Laplacian(cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY), cv2.CV_64F).var()

    got = crop_sharpness_batch(img, boxes)
    for g, b in zip(got, boxes):
        want = ref(b)
        assert abs(g - want) <= 1e-9 * max(1.0, abs(want)), (b, g, want)
    assert crop_sharpness(img, boxes[0]) == got[0]
    assert mean_face_sharpness(img, boxes[:2]) == float(np.mean(got[:2])) and mean_face_sharpness(img, []) == 0
    full = ImageCache(img).laplacian_variance
    assert isolation_bonus(got[0], full) == max(1.0, got[0] / (full + 1)) and isolation_bonus(99.0, full, face_count=0) == 1.0
