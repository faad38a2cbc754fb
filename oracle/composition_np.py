"""TEST INFRASTRUCTURE ONLY — NumPy restatement of the OpenCV calls of the reference's composition analyzer.

Restates, for uint8 gray planes (analyzers/composition.py of the reference):
  * `cv2.GaussianBlur(gray, (5, 5), 0)`        (:215)  -> gaussian_blur5
  * `cv2.Canny(src, low, high)`                (:36, :218)  aperture 3, L1 gradient -> canny
  * `np.median(gray)` from a 256-bin histogram (:33)   -> median_from_hist
OpenCV's source is not under /root/reference (third-party wheel, cv2 4.13 installed here): the restatement follows
the published algorithm (fixed-point binomial kernel; Sobel with replicated borders, |dx| + |dy|, non-maximum
suppression with 15-bit fixed-point tangents, hysteresis) and is PINNED against the installed cv2 itself on random,
smooth and structured planes of many sizes by tests/test_oracle_composition.py.  Only tests / smoke / the bench's CPU
legs may import this module; the product path (facet_b200/csrc/canny.cu) never does.
"""
import numpy as np


def gaussian_blur5(gray: np.ndarray) -> np.ndarray:
    h, w = gray.shape
    p = np.pad(gray.astype(np.int32), 2, mode="reflect") if min(h, w) > 2 else _pad_reflect101(gray.astype(np.int32), 2)
    k = (1, 4, 6, 4, 1)
    hor = sum(k[i] * p[:, i:i + w] for i in range(5))
    ver = sum(k[i] * hor[i:i + h, :] for i in range(5))
    return ((ver + 128) >> 8).astype(np.uint8)


def _pad_reflect101(a, r):
    def idx(n):
        out = []
        for i in range(-r, n + r):
            if n == 1:
                out.append(0)
                continue
            while i < 0 or i >= n:
                i = -i if i < 0 else 2 * n - 2 - i
            out.append(i)
        return np.array(out)
    return a[idx(a.shape[0])][:, idx(a.shape[1])]


def canny_classes(src: np.ndarray, low: int, high: int) -> np.ndarray:
    """0 = not an edge candidate, 1 = candidate (local maximum with magnitude > low), 2 = strong (> high)."""
    p = np.pad(src.astype(np.int32), 1, mode="edge")
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    mag = np.abs(dx) + np.abs(dy)
    m = np.pad(mag, 1, mode="constant")
    c = m[1:-1, 1:-1]
    tg22 = int(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
    x = np.abs(dx).astype(np.int64)
    y = np.abs(dy).astype(np.int64) << 15
    t22 = x * tg22
    t67 = t22 + (x << 16)
    left, right, up, down = m[1:-1, :-2], m[1:-1, 2:], m[:-2, 1:-1], m[2:, 1:-1]
    same_sign = (dx ^ dy) >= 0
    up_d = np.where(same_sign, m[:-2, :-2], m[:-2, 2:])
    dn_d = np.where(same_sign, m[2:, 2:], m[2:, :-2])
    keep = np.where(y < t22, (c > left) & (c >= right), np.where(y > t67, (c > up) & (c >= down), (c > up_d) & (c > dn_d)))
    cand = (c > low) & keep
    return cand.astype(np.uint8) + (cand & (c > high)).astype(np.uint8)


def canny(src: np.ndarray, low: int, high: int) -> np.ndarray:
    import scipy.ndimage as ndi
    cls = canny_classes(src, low, high)
    lab, n = ndi.label(cls > 0, structure=np.ones((3, 3), np.int32))
    keep = np.zeros(n + 1, bool)
    keep[np.unique(lab[cls == 2])] = True
    keep[0] = False
    return (keep[lab] * 255).astype(np.uint8)


def median_from_hist(hist256) -> float:
    h = np.asarray(hist256, dtype=np.int64)
    n = int(h.sum())
    cum = np.cumsum(h)
    return (int(np.searchsorted(cum, (n - 1) // 2 + 1)) + int(np.searchsorted(cum, n // 2 + 1))) / 2.0
