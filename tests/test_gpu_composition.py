"""Rule-based composition analyzer on the GPU (csrc/canny.cu + analyzers/composition.py) against the installed
OpenCV (edge maps, bit-exact) and against goldens written by the unmodified reference (dicts, boxes)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import GOLDEN_DIR


def _gold():
    with open(os.path.join(GOLDEN_DIR, "composition_golden.json")) as f:
        return json.load(f)


def _planes(h, w, rng):
    import cv2
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = ((np.sin(xx / 9.0) + np.cos(yy / 7.0)) * 60 + 128 + rng.normal(0, 6, (h, w))).clip(0, 255).astype(np.uint8)
    boxes = np.zeros((h, w), np.uint8)
    cv2.rectangle(boxes, (w // 4, h // 4), (3 * w // 4, 3 * h // 4), 200, -1)
    cv2.line(boxes, (0, 0), (w - 1, h - 1), 90, 3)
    boxes = (boxes + rng.integers(0, 20, (h, w))).astype(np.uint8)
    return [rng.integers(0, 256, (h, w), dtype=np.uint8), smooth, boxes, np.full((h, w), 255, np.uint8),
            ((xx + yy) % 2 * 255).astype(np.uint8)]


@pytest.mark.parametrize("shape", [(64, 80), (3, 3), (7, 5), (2, 9), (1, 6), (5, 1), (240, 333), (400, 601), (31, 65), (33, 64),
                                   (683, 1024), (1500, 2100)])
def test_edge_maps_equal_opencv_and_oracle(shape):
    import cv2
    from facet_b200 import ops
    from oracle import composition_np as co
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for k, g in enumerate(_planes(*shape, rng)):
        med = float(np.median(g))
        for blur, lo, hi in ((True, 50, 150), (False, int(max(0, 0.5 * med)), int(min(255, 1.5 * med))), (False, 0, 0), (True, 254, 255)):
            want = cv2.Canny(cv2.GaussianBlur(g, (5, 5), 0) if blur else g, lo, hi)
            got, cnt = ops.canny_edges(g, lo, hi, blur=blur, want_count=True)
            got = got.cpu().numpy()
            assert np.array_equal(got, want), (shape, k, blur, lo, hi, int((got != want).sum()))
            assert int(cnt.item()) == int((want > 0).sum())
            if shape[0] * shape[1] <= 400 * 601:
                src = co.gaussian_blur5(g) if blur else g
                assert np.array_equal(co.canny(src, lo, hi), want)


def test_gray_plane_and_median():
    import cv2
    from facet_b200 import ops
    from facet_b200.analyzers.composition import _median_from_hist
    from facet_b200.synth import synth_image_bgr
    for idx, h, w in ((0, 97, 131), (6, 333, 501), (3, 2, 2), (5, 683, 1024)):
        img = synth_image_bgr(idx, h, w)
        gray, hist = ops.gray_plane(img, want_hist=True)
        want = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(gray.cpu().numpy(), want)
        assert np.array_equal(hist.cpu().numpy().view(np.uint32), np.bincount(want.ravel(), minlength=256))
        assert _median_from_hist(hist.cpu().numpy()) == float(np.median(want))
        gray_rgb = ops.gray_plane(np.ascontiguousarray(img[:, :, ::-1]), rgb_order=True)
        assert np.array_equal(gray_rgb.cpu().numpy(), want)


def test_composition_dicts_match_reference_golden():
    from facet_b200 import ops
    from facet_b200.analyzers import ImageCache
    from facet_b200.analyzers.composition import CompositionAnalyzer, _median_from_hist
    from facet_b200.synth import synth_image_bgr
    gold = _gold()
    for case in gold["cases"]:
        h, w = case["height"], case["width"]
        img = synth_image_bgr(case["index"], h, w)
        gray, hist = ops.gray_plane(img, want_hist=True)
        assert _median_from_hist(hist.cpu().numpy()) == case["median"]
        e1 = ops.canny_edges(gray, 50, 150, blur=True).cpu().numpy()
        e2 = ops.canny_edges(gray, case["lower"], case["upper"], blur=False).cpu().numpy()
        assert hashlib.sha256(e1.tobytes()).hexdigest() == case["edges_lines_sha256"], (case["index"], h, w)
        assert hashlib.sha256(e2.tobytes()).hexdigest() == case["edges_subject_sha256"], (case["index"], h, w)
        cache = ImageCache(img) if min(h, w) >= 2 else None
        lead = CompositionAnalyzer.detect_leading_lines(img, cache=cache)
        assert lead == case["leading_lines"], (case["index"], lead, case["leading_lines"])
        assert CompositionAnalyzer.detect_subject_region(img) == case["subject_bbox"]
        place = CompositionAnalyzer.get_placement_data(None, w, h, None, img_cv=img)
        assert place == case["placement"]
        assert CompositionAnalyzer.integrate_leading_lines(place["score"], lead["leading_lines_score"], False) == case["integrated"]
    assert CompositionAnalyzer.detect_leading_lines(None) == {"leading_lines_score": 0, "line_count": 0}
    assert CompositionAnalyzer.detect_subject_region(None) is None


def test_device_frame_input_and_timing():
    """The analyzer takes a frame that is already in device memory; the device part of a 24 MP frame is timed."""
    import torch
    from facet_b200 import ops
    from facet_b200.analyzers.composition import CompositionAnalyzer
    from facet_b200.synth import synth_image_bgr
    img = synth_image_bgr(4, 4000, 6000)
    t = torch.from_numpy(img).cuda()
    gold = [c for c in _gold()["cases"] if c["height"] == 4000][0]
    assert CompositionAnalyzer.detect_leading_lines(t) == gold["leading_lines"]
    gray = ops.gray_plane(t)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        ops.canny_edges(ops.gray_plane(t), 50, 150, blur=True)
    ev[1].record()
    torch.cuda.synchronize()
    print("gray + blur + Canny on the device, 24 MP frame: %.3f ms" % (ev[0].elapsed_time(ev[1]) / 5))
    del gray
