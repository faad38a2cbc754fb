"""ORACLE (test infrastructure, never shipped): NumPy restatement of the reference's
technical metrics.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product path
(``facet_b200``) must never call it.

What is restated, and where it lives in the reference (paths under /root/reference):

* ``analyzers/image_cache.py:22-32``  ImageCache: gray / hsv / Laplacian variance
* ``analyzers/technical.py:39-58``    get_sharpness_data
* ``analyzers/technical.py:79-113``   get_color_harmony_data
* ``analyzers/technical.py:126-215``  get_histogram_data
* ``analyzers/technical.py:219-242``  detect_monochrome
* ``analyzers/technical.py:245-273``  get_dynamic_range
* ``analyzers/technical.py:276-305``  get_noise_estimate
* ``analyzers/technical.py:308-342``  get_contrast_score

The pixel arithmetic under those functions belongs to third-party wheels that
are not vendored in the reference: OpenCV (``opencv-python>=4.8.0``,
requirements.txt:23; 4.13.0 installed here), NumPy (2.3.5 here) and SciPy (1.18.1
here).  Their published integer algorithms are restated below:

* BGR2GRAY 8-bit: 15-bit fixed point, ``(3735 B + 19235 G + 9798 R + 2^14) >> 15``
* BGR2HSV 8-bit, hue range 180: ``sdiv/hdiv`` tables with 12-bit shift
* Laplacian ksize=1 (4-neighbour) and filter2D, both BORDER_REFLECT_101
* calcHist: integer bin counts returned as float32
* np.percentile(method="linear"), np.std (population), scipy kurtosis (biased, Fisher)

Parity is pinned by ``tests/golden/*.json`` — outputs of the reference modules
themselves, produced in the build container by ``tests/golden/make_golden.py`` —
and by ``tests/test_oracle_technical.py`` which also checks the integer stages
against the installed cv2 directly (all 2^24 colours for gray/HSV).
"""
from __future__ import annotations

import math
import struct

import numpy as np

HSV_SHIFT = 12


def sdiv_table() -> np.ndarray:
    """OpenCV ``sdiv_table``: round(255 * 4096 / i), entry 0 is 0."""
    t = np.zeros(256, np.int32)
    i = np.arange(1, 256, dtype=np.float64)
    t[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int32)
    return t


def hdiv_table180() -> np.ndarray:
    """OpenCV ``hdiv_table180``: round(180 * 4096 / (6 i)), entry 0 is 0."""
    t = np.zeros(256, np.int32)
    i = np.arange(1, 256, dtype=np.float64)
    t[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int32)
    return t


_SDIV = sdiv_table()
_HDIV = hdiv_table180()


def gray_u8(img_bgr: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(BGR2GRAY) on uint8 (image_cache.py:30)."""
    b = img_bgr[..., 0].astype(np.int32)
    g = img_bgr[..., 1].astype(np.int32)
    r = img_bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def hsv_u8(img_bgr: np.ndarray):
    """cv2.cvtColor(BGR2HSV) on uint8 (image_cache.py:31) -> (h, s, v) int32 planes."""
    b = img_bgr[..., 0].astype(np.int32)
    g = img_bgr[..., 1].astype(np.int32)
    r = img_bgr[..., 2].astype(np.int32)
    v = np.maximum(np.maximum(b, g), r)
    vmin = np.minimum(np.minimum(b, g), r)
    d = v - vmin
    s = (d * _SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    hraw = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (hraw * _HDIV[d] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(h < 0, h + 180, h)
    return h, s, v


def _reflect101_pad(a: np.ndarray) -> np.ndarray:
    return np.pad(a, 1, mode="reflect")


def laplacian_i32(gray: np.ndarray) -> np.ndarray:
    """cv2.Laplacian(gray, CV_64F), ksize=1, as exact integers."""
    p = _reflect101_pad(gray.astype(np.int32))
    return p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4 * p[1:-1, 1:-1]


def immerkaer_i32(gray: np.ndarray) -> np.ndarray:
    """cv2.filter2D(gray, -1, [[1,-2,1],[-2,4,-2],[1,-2,1]]) as exact integers."""
    p = _reflect101_pad(gray.astype(np.int32))
    dxx = p[:, :-2] - 2 * p[:, 1:-1] + p[:, 2:]          # (H+2, W)
    return dxx[:-2] - 2 * dxx[1:-1] + dxx[2:]


def tech_stats(img_bgr: np.ndarray) -> dict:
    """The sufficient statistics one pass over the image has to produce.

    Returns hist256 (int64[256]), hs_hist (int64[180,256]), sum_lap, sum_lap_sq,
    sum_abs_noise (Python ints), height, width.
    """
    if img_bgr.shape[0] < 2 or img_bgr.shape[1] < 2:
        raise ValueError("reflect-101 borders need at least 2 rows and 2 columns")
    gray = gray_u8(img_bgr)
    h, s, _ = hsv_u8(img_bgr)
    lap = laplacian_i32(gray).astype(np.int64)
    nz = immerkaer_i32(gray).astype(np.int64)
    hist256 = np.bincount(gray.ravel(), minlength=256).astype(np.int64)
    hs = np.bincount((h * 256 + s).ravel(), minlength=180 * 256).astype(np.int64).reshape(180, 256)
    return {
        "hist256": hist256,
        "hs_hist": hs,
        "sum_lap": int(lap.sum()),
        "sum_lap_sq": int((lap * lap).sum()),
        "sum_abs_noise": int(np.abs(nz).sum()),
        "height": int(img_bgr.shape[0]),
        "width": int(img_bgr.shape[1]),
    }


# ---------------------------------------------------------------------------
# closed-form metrics from the sufficient statistics
# ---------------------------------------------------------------------------

def laplacian_variance(st: dict) -> float:
    """``cv2.Laplacian(gray, CV_64F).var()`` (image_cache.py:32): population variance."""
    n = st["height"] * st["width"]
    mean = st["sum_lap"] / n
    return st["sum_lap_sq"] / n - mean * mean


def sharpness_data(st: dict) -> dict:
    var = laplacian_variance(st)
    return {"raw_variance": var, "normalized": float(min(10.0, var / 50.0))}


def color_harmony_data(st: dict) -> dict:
    """technical.py:94-113.  calcHist returns float32, so p and p*log2(p) are float32."""
    hist = st["hs_hist"].astype(np.float32)
    total = hist.sum()
    if total > 0:
        p = hist / total
        nz = p[p > 0]
        ent = -np.sum(nz * np.log2(nz))
    else:
        ent = 0
    return {"raw_entropy": ent, "normalized": float(min(10.0, ent * 10.0 / 15.5))}


def _kurtosis_fisher_biased(x: np.ndarray) -> float:
    x = np.asarray(x)
    m = x.mean()
    d = x - m
    m2 = np.mean(d * d)
    m4 = np.mean(d * d * d * d)
    if m2 == 0:
        return float("nan")
    return float(m4 / (m2 * m2) - 3.0)


def histogram_data(st: dict, shadow_threshold: float = 0.15, highlight_threshold: float = 0.10) -> dict:
    """technical.py:152-215."""
    hist = st["hist256"].astype(np.float32)
    total = hist.sum()
    hn = hist / total if total > 0 else hist
    histogram_bytes = struct.pack("256f", *hn)
    bins = np.arange(256)
    mean_val = np.sum(bins * hn)
    spread = np.sqrt(np.sum(((bins - mean_val) ** 2) * hn))
    mean_luminance = mean_val / 255.0
    shadow_mass = np.sum(hn[:30])
    highlight_mass = np.sum(hn[225:])
    shadow_clipped = 1 if shadow_mass > shadow_threshold else 0
    highlight_clipped = 1 if highlight_mass > highlight_threshold else 0
    lower_third = np.sum(hn[:85])
    upper_third = np.sum(hn[170:])
    is_silhouette = 1 if (lower_third > 0.35 and upper_third > 0.25) else 0
    bimodality = -_kurtosis_fisher_biased(hn * 256)
    luminance_penalty = abs(mean_luminance - 0.5) * 8
    spread_bonus = min(4.0, spread / 20.0)
    bimodality_penalty = max(0, bimodality - 1.0) * 0.6
    clipping_penalty = 0
    if not is_silhouette:
        clipping_penalty = shadow_mass * 4.0 + highlight_mass * 5.0
    exposure = max(0, min(10.0, 7.0 - luminance_penalty + spread_bonus - bimodality_penalty - clipping_penalty))
    return {
        "histogram_bytes": histogram_bytes,
        "spread": round(float(spread), 4),
        "mean_luminance": round(float(mean_luminance), 4),
        "bimodality": round(float(bimodality), 4),
        "exposure_score": round(float(exposure), 2),
        "shadow_clipped": shadow_clipped,
        "highlight_clipped": highlight_clipped,
        "is_silhouette": is_silhouette,
    }


def monochrome_data(st: dict, threshold: float = 0.1) -> dict:
    """technical.py:237-242: mean of the S plane == sum_s s * colsum(hs_hist) / N."""
    n = st["height"] * st["width"]
    s_counts = st["hs_hist"].sum(axis=0)
    mean_sat = float(np.dot(np.arange(256, dtype=np.int64), s_counts)) / n / 255.0
    return {"is_monochrome": 1 if mean_sat < threshold else 0, "mean_saturation": round(mean_sat, 4)}


def percentile_from_hist(hist256: np.ndarray, q: float) -> float:
    """np.percentile(gray, q) (method='linear') from the 256-bin counts.

    NumPy: virtual index (n-1)*q/100, neighbours a<=b from the sorted data,
    ``a + (b-a)*t`` and, for t >= 0.5, ``b - (b-a)*(1-t)``.
    """
    n = int(hist256.sum())
    cum = np.cumsum(hist256)
    pos = (n - 1) * (q / 100.0)
    lo = math.floor(pos)
    t = pos - lo
    hi = min(lo + 1, n - 1)
    a = float(np.searchsorted(cum, lo + 1, side="left"))
    b = float(np.searchsorted(cum, hi + 1, side="left"))
    diff = b - a
    if t >= 0.5:
        return b - diff * (1 - t)
    return a + diff * t


def dynamic_range_data(st: dict) -> dict:
    p2 = percentile_from_hist(st["hist256"], 2)
    p98 = percentile_from_hist(st["hist256"], 98)
    if p2 < 1:
        p2 = 1
    return {"dynamic_range_stops": round(float(np.log2(max(p98, 1) / p2)), 2)}


def noise_data(st: dict) -> dict:
    """technical.py:302-305 (sum over all H*W outputs, divisor (W-2)(H-2))."""
    w, h = st["width"], st["height"]
    sigma = float(st["sum_abs_noise"]) * np.sqrt(0.5 * np.pi) / (6 * (w - 2) * (h - 2))
    return {"noise_sigma": round(float(sigma), 2)}


def contrast_data(st: dict) -> dict:
    """technical.py:326-342; np.std(gray) from the histogram (population std)."""
    hist = st["hist256"]
    p5 = percentile_from_hist(hist, 5)
    p95 = percentile_from_hist(hist, 95)
    pc = (p95 - p5) / 255.0
    n = int(hist.sum())
    k = np.arange(256, dtype=np.int64)
    s1 = int(np.dot(k, hist))
    s2 = int(np.dot(k * k, hist))
    var = s2 / n - (s1 / n) ** 2
    rms = math.sqrt(max(var, 0.0)) / 255.0
    score = min(10.0, pc * 5.0 + rms * 20.0)
    return {
        "contrast_score": round(float(score), 2),
        "percentile_contrast": round(float(pc), 4),
        "rms_contrast": round(float(rms), 4),
    }


def all_metrics(img_bgr: np.ndarray, mono_threshold: float = 0.1) -> dict:
    """Everything `_process_batch` (batch_processor.py:198-233) asks the analyzer for."""
    st = tech_stats(img_bgr)
    return {
        "stats": st,
        "sharpness": sharpness_data(st),
        "color": color_harmony_data(st),
        "histogram": histogram_data(st),
        "monochrome": monochrome_data(st, mono_threshold),
        "dynamic_range": dynamic_range_data(st),
        "noise": noise_data(st),
        "contrast": contrast_data(st),
    }


def roi_laplacian_sums(img_bgr: np.ndarray, box) -> tuple:
    """analyzers/face.py:272-279 `_get_crop_sharpness`: Laplacian variance of a crop.

    The crop is converted to gray and filtered on its own, so the reflect-101
    border is the border of the crop.  Returns (n, sum_lap, sum_lap_sq).
    """
    x1, y1, x2, y2 = [int(v) for v in box]
    crop = img_bgr[y1:y2, x1:x2]
    if crop.size == 0:
        return (0, 0, 0)
    g = gray_u8(crop)
    if g.shape[0] < 2 or g.shape[1] < 2:
        # reflect-101 of an extent of one pixel is the pixel itself (checked against cv2.Laplacian): that axis
        # contributes nothing, the other one its second difference
        gi = g.astype(np.int64)
        lap = np.zeros_like(gi)
        for axis in (0, 1):
            if g.shape[axis] >= 2:
                pad = [(1, 1) if a == axis else (0, 0) for a in (0, 1)]
                e = np.pad(gi, pad, mode="reflect")
                lo = e[:-2] if axis == 0 else e[:, :-2]
                hi = e[2:] if axis == 0 else e[:, 2:]
                lap += lo + hi - 2 * gi
        return (int(lap.size), int(lap.sum()), int((lap * lap).sum()))
    lap = laplacian_i32(g).astype(np.int64)
    return (int(lap.size), int(lap.sum()), int((lap * lap).sum()))
