"""Metric dicts of ``analyzers/technical.py`` as closed forms of the GPU sufficient statistics.

The CUDA pass (csrc/tech_stats.cu) returns, per image: the 256-bin luminance histogram, three
exact integer sums (Laplacian, Laplacian^2, |Immerkaer|) and two reductions of the H-S
histogram (entropy, sum of saturation).  Everything the reference computes from pixels after
that point is a function of those numbers plus (H, W); this module evaluates those functions
on the host in float64/float32 exactly where the reference does.

Reference lines are cited per function (paths under rlorenzo/facet).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass

import numpy as np

_BINS = np.arange(256)
_BINS_I64 = np.arange(256, dtype=np.int64)


@dataclass
class TechStats:
    """Sufficient statistics of one image (all exact integers except entropy)."""
    height: int
    width: int
    hist256: np.ndarray          # int64[256]
    sum_lap: int
    sum_lap_sq: int
    sum_abs_noise: int
    hs_entropy: float            # -sum p log2 p over the 180x256 H-S histogram
    sum_saturation: float        # sum of the S plane
    hs_hist: np.ndarray | None = None   # uint32[180,256] when the caller asked for it

    @property
    def n_pixels(self) -> int:
        return self.height * self.width


def laplacian_variance(st: TechStats) -> float:
    """image_cache.py:32 — population variance of the CV_64F Laplacian."""
    n = st.n_pixels
    mean = st.sum_lap / n
    return st.sum_lap_sq / n - mean * mean


def sharpness(st: TechStats) -> dict:
    """technical.py:55-58."""
    var = laplacian_variance(st)
    return {"raw_variance": var, "normalized": float(min(10.0, var / 50.0))}


def color_harmony(st: TechStats) -> dict:
    """technical.py:97-113 (entropy itself comes from hs_derive_kernel)."""
    ent = st.hs_entropy
    return {"raw_entropy": ent, "normalized": float(min(10.0, ent * 10.0 / 15.5))}


def histogram(st: TechStats, shadow_threshold: float = 0.15, highlight_threshold: float = 0.10) -> dict:
    """technical.py:152-215.  calcHist output is float32, so the normalised histogram is too."""
    counts = st.hist256.astype(np.float32)
    total = counts.sum()
    p = counts / total if total > 0 else counts
    blob = p.astype("<f4", copy=False).tobytes()      # = struct.pack('256f', *p) of technical.py:158, 20x cheaper
    mu = np.sum(_BINS * p)
    spread = np.sqrt(np.sum(((_BINS - mu) ** 2) * p))
    lum = mu / 255.0
    shadow = np.sum(p[:30])
    highlight = np.sum(p[225:])
    low, high = np.sum(p[:85]), np.sum(p[170:])
    silhouette = 1 if (low > 0.35 and high > 0.25) else 0
    # scipy.stats.kurtosis(p*256, fisher=True) with bias: m4/m2^2 - 3 over the 256 bin heights
    x = p * 256
    dev = x - x.mean()
    m2 = np.mean(dev * dev)
    m4 = np.mean(dev * dev * dev * dev)
    bimodality = -(float(m4 / (m2 * m2)) - 3.0) if m2 != 0 else float("nan")
    score = 7.0 - abs(lum - 0.5) * 8 + min(4.0, spread / 20.0) - max(0, bimodality - 1.0) * 0.6
    if not silhouette:
        score -= shadow * 4.0 + highlight * 5.0
    score = max(0, min(10.0, score))
    return {
        "histogram_bytes": blob,
        "spread": round(float(spread), 4),
        "mean_luminance": round(float(lum), 4),
        "bimodality": round(float(bimodality), 4),
        "exposure_score": round(float(score), 2),
        "shadow_clipped": 1 if shadow > shadow_threshold else 0,
        "highlight_clipped": 1 if highlight > highlight_threshold else 0,
        "is_silhouette": silhouette,
    }


def monochrome(st: TechStats, threshold: float = 0.1) -> dict:
    """technical.py:237-242."""
    mean_sat = st.sum_saturation / st.n_pixels / 255.0
    return {"is_monochrome": 1 if mean_sat < threshold else 0, "mean_saturation": round(mean_sat, 4)}


def _order_statistic(cum: np.ndarray, rank: int) -> float:
    """Value of the rank-th (0-based) smallest pixel given cumulative bin counts."""
    return float(np.searchsorted(cum, rank + 1, side="left"))


def percentile(hist256: np.ndarray, q: float) -> float:
    """np.percentile(gray, q) — 'linear' method incl. NumPy's two-sided lerp."""
    cum = np.cumsum(hist256)
    n = int(cum[-1])
    virtual = (n - 1) * (q / 100.0)
    k = math.floor(virtual)
    t = virtual - k
    a = _order_statistic(cum, k)
    b = _order_statistic(cum, min(k + 1, n - 1))
    return b - (b - a) * (1 - t) if t >= 0.5 else a + (b - a) * t


def dynamic_range(st: TechStats) -> dict:
    """technical.py:263-273."""
    p2, p98 = percentile(st.hist256, 2), percentile(st.hist256, 98)
    p2 = 1 if p2 < 1 else p2
    return {"dynamic_range_stops": round(float(np.log2(max(p98, 1) / p2)), 2)}


def noise(st: TechStats) -> dict:
    """technical.py:302-305."""
    denom = 6 * (st.width - 2) * (st.height - 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma = np.float64(st.sum_abs_noise) * np.sqrt(0.5 * np.pi) / denom
    return {"noise_sigma": round(float(sigma), 2)}


def contrast(st: TechStats) -> dict:
    """technical.py:326-342: percentiles and np.std(gray) from the histogram."""
    h = st.hist256
    p5, p95 = percentile(h, 5), percentile(h, 95)
    pc = (p95 - p5) / 255.0
    n = int(h.sum())
    s1 = int(np.dot(_BINS_I64, h))
    s2 = int(np.dot(_BINS_I64 * _BINS_I64, h))
    var = s2 / n - (s1 / n) ** 2
    rms = math.sqrt(var if var > 0 else 0.0) / 255.0
    return {
        "contrast_score": round(float(min(10.0, pc * 5.0 + rms * 20.0)), 2),
        "percentile_contrast": round(float(pc), 4),
        "rms_contrast": round(float(rms), 4),
    }


# ---------------------------------------------------------------------------------------------
# The same closed forms over a whole batch (one image per row).  Every array operation is the
# scalar function's operation applied row-wise in the same dtype and order (reductions run along
# the contiguous axis, i.e. NumPy's pairwise sum per row), so the dicts are bit-identical to the
# scalar ones (tests/test_closed_form_batch.py); only the Python-level work per image shrinks to
# the rounding of the final numbers.
# ---------------------------------------------------------------------------------------------
def _percentiles_batch(hists: np.ndarray, qs) -> list:
    """np.percentile(gray, q) per row for every q in qs -> list of float64 [n] arrays."""
    cum = np.cumsum(hists, axis=1)
    n = cum[:, -1]
    out = []
    for q in qs:
        virtual = (n - 1) * (q / 100.0)
        k = np.floor(virtual)
        t = virtual - k
        ki = k.astype(np.int64)
        a = (cum < (ki + 1)[:, None]).sum(axis=1).astype(np.float64)           # searchsorted(cum, k + 1, 'left')
        b = (cum < (np.minimum(ki + 1, n - 1) + 1)[:, None]).sum(axis=1).astype(np.float64)
        out.append(np.where(t >= 0.5, b - (b - a) * (1 - t), a + (b - a) * t))
    return out


def all_metrics_batch(height: int, width: int, hists: np.ndarray, sum_lap, sum_lap_sq, sum_abs_noise, hs_entropy,
                      sum_saturation, mono_threshold: float = 0.1, shadow_threshold: float = 0.15,
                      highlight_threshold: float = 0.10) -> list:
    """Per image: {'sharpness', 'color', 'histogram', 'monochrome', 'dynamic_range', 'noise', 'contrast'} with
    the dicts of the scalar functions above.  hists: int64 [n,256]; the other arguments are length-n sequences."""
    hists = np.ascontiguousarray(hists, dtype=np.int64)
    n_img = hists.shape[0]
    if n_img == 0:
        return []
    npx = height * width
    # sharpness / colour / monochrome / noise: a handful of scalar operations per image
    counts = hists.astype(np.float32)
    total = counts.sum(axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = np.where((total > 0)[:, None], counts / total[:, None], counts)
    p = np.ascontiguousarray(p, dtype=np.float32)
    bp = _BINS * p                                                # float64 [n,256]
    mu = np.sum(bp, axis=1)
    spread = np.sqrt(np.sum(((_BINS - mu[:, None]) ** 2) * p, axis=1))
    lum = mu / 255.0
    shadow = np.sum(p[:, :30], axis=1)
    highlight = np.sum(p[:, 225:], axis=1)
    low, high = np.sum(p[:, :85], axis=1), np.sum(p[:, 170:], axis=1)
    x = p * 256
    dev = x - x.mean(axis=1)[:, None]
    d2 = dev * dev
    m2 = np.mean(d2, axis=1)
    m4 = np.mean(d2 * dev * dev, axis=1)
    p2, p98, p5, p95 = _percentiles_batch(hists, (2, 98, 5, 95))
    ntot = hists.sum(axis=1)
    s1 = hists @ _BINS_I64
    s2 = hists @ (_BINS_I64 * _BINS_I64)
    denom = 6 * (width - 2) * (height - 2)
    root_half_pi = np.sqrt(0.5 * np.pi)
    out = []
    for i in range(n_img):
        mean = int(sum_lap[i]) / npx
        var = int(sum_lap_sq[i]) / npx - mean * mean
        ent = float(hs_entropy[i])
        silhouette = 1 if (low[i] > 0.35 and high[i] > 0.25) else 0
        bimodality = -(float(m4[i] / (m2[i] * m2[i])) - 3.0) if m2[i] != 0 else float("nan")
        score = 7.0 - abs(lum[i] - 0.5) * 8 + min(4.0, spread[i] / 20.0) - max(0, bimodality - 1.0) * 0.6
        if not silhouette:
            score -= shadow[i] * 4.0 + highlight[i] * 5.0
        score = max(0, min(10.0, score))
        mean_sat = float(sum_saturation[i]) / npx / 255.0
        lo2 = 1 if p2[i] < 1 else p2[i]
        with np.errstate(divide="ignore", invalid="ignore"):
            sigma = np.float64(int(sum_abs_noise[i])) * root_half_pi / denom
        pc = (p95[i] - p5[i]) / 255.0
        nn = int(ntot[i])
        v = int(s2[i]) / nn - (int(s1[i]) / nn) ** 2
        rms = math.sqrt(v if v > 0 else 0.0) / 255.0
        out.append({
            "sharpness": {"raw_variance": var, "normalized": float(min(10.0, var / 50.0))},
            "color": {"raw_entropy": ent, "normalized": float(min(10.0, ent * 10.0 / 15.5))},
            "histogram": {
                "histogram_bytes": p[i].tobytes(),
                "spread": round(float(spread[i]), 4),
                "mean_luminance": round(float(lum[i]), 4),
                "bimodality": round(float(bimodality), 4),
                "exposure_score": round(float(score), 2),
                "shadow_clipped": 1 if shadow[i] > shadow_threshold else 0,
                "highlight_clipped": 1 if highlight[i] > highlight_threshold else 0,
                "is_silhouette": silhouette,
            },
            "monochrome": {"is_monochrome": 1 if mean_sat < mono_threshold else 0, "mean_saturation": round(mean_sat, 4)},
            "dynamic_range": {"dynamic_range_stops": round(float(np.log2(max(p98[i], 1) / lo2)), 2)},
            "noise": {"noise_sigma": round(float(sigma), 2)},
            "contrast": {"contrast_score": round(float(min(10.0, pc * 5.0 + rms * 20.0)), 2),
                         "percentile_contrast": round(float(pc), 4), "rms_contrast": round(float(rms), 4)},
        })
    return out
