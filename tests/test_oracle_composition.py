"""The oracle's blur / Canny / median restatement against the installed OpenCV / NumPy (CPU)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import composition_np as co


def _planes(h, w, rng):
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = ((np.sin(xx / 9.0) + np.cos(yy / 7.0)) * 60 + 128 + rng.normal(0, 6, (h, w))).clip(0, 255).astype(np.uint8)
    boxes = np.zeros((h, w), np.uint8)
    cv2.rectangle(boxes, (w // 4, h // 4), (3 * w // 4, 3 * h // 4), 200, -1)
    cv2.line(boxes, (0, 0), (w - 1, h - 1), 90, 3)
    boxes = (boxes + rng.integers(0, 20, (h, w))).astype(np.uint8)
    return [rng.integers(0, 256, (h, w), dtype=np.uint8), smooth, boxes, np.full((h, w), 255, np.uint8),
            ((xx + yy) % 2 * 255).astype(np.uint8)]


@pytest.mark.parametrize("shape", [(64, 80), (3, 3), (7, 5), (2, 9), (1, 6), (5, 1), (240, 333), (400, 601)])
def test_blur_and_canny_match_opencv(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for g in _planes(*shape, rng):
        b = cv2.GaussianBlur(g, (5, 5), 0)
        assert np.array_equal(b, co.gaussian_blur5(g))
        med = float(np.median(g))
        assert med == co.median_from_hist(np.bincount(g.ravel(), minlength=256))
        for src, lo, hi in ((b, 50, 150), (g, int(max(0, 0.5 * med)), int(min(255, 1.5 * med)))):
            assert np.array_equal(cv2.Canny(src, lo, hi), co.canny(src, lo, hi))
