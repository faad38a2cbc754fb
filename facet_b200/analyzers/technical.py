"""``TechnicalAnalyzer`` with the reference's static-method interface
(analyzers/technical.py:13-342), computed from one GPU pass per image.

Every ``get_*`` accepts the same arguments as the reference (``image_cv`` BGR uint8 array,
optional ``cache``) and returns a dict with the same keys, rounding and ``None``-image
defaults (technical.py:46, :86, :135-145, :229, :254, :286, :317).
"""
from __future__ import annotations

import numpy as np

from . import _closed_form as cf
from .image_cache import ImageCache


def _stats(image_cv, cache):
    if cache is not None and getattr(cache, "stats", None) is not None:
        return cache.stats
    return ImageCache(image_cv).stats


class TechnicalAnalyzer:
    """Objective image metrics; the pixel work runs in csrc/tech_stats.cu."""

    @staticmethod
    def get_iso_adjusted_sharpness(raw_variance, iso):
        # technical.py:17-26
        if iso is None or iso <= 100:
            return raw_variance
        return raw_variance * (1.0 + 0.15 * np.log2(iso / 100))

    @staticmethod
    def get_sharpness_score(image_cv):
        if image_cv is None:
            return 0
        return cf.sharpness(_stats(image_cv, None))["normalized"]

    @staticmethod
    def get_sharpness_data(image_cv, cache=None):
        if image_cv is None:
            return {"raw_variance": 0, "normalized": 0}
        return cf.sharpness(_stats(image_cv, cache))

    @staticmethod
    def get_color_harmony(image_cv):
        return cf.color_harmony(_stats(image_cv, None))["normalized"]

    @staticmethod
    def get_color_harmony_data(image_cv, cache=None):
        if image_cv is None:
            return {"raw_entropy": 0, "normalized": 0}
        return cf.color_harmony(_stats(image_cv, cache))

    @staticmethod
    def get_exposure_score(image_cv):
        # technical.py:116-123: clipped-pixel penalty from the luminance histogram
        h = _stats(image_cv, None).hist256
        total = int(h.sum())
        penalty = (int(h[:6].sum()) / total + int(h[250:].sum()) / total) * 10
        return max(0, 10 - penalty)

    @staticmethod
    def get_histogram_data(image_cv, shadow_threshold=0.15, highlight_threshold=0.10, cache=None):
        if image_cv is None:
            return {"histogram_bytes": None, "spread": 0, "mean_luminance": 0.5, "bimodality": 0,
                    "exposure_score": 5.0, "shadow_clipped": 0, "highlight_clipped": 0, "is_silhouette": 0}
        return cf.histogram(_stats(image_cv, cache), shadow_threshold, highlight_threshold)

    @staticmethod
    def detect_monochrome(image_cv, threshold=0.1, cache=None):
        if image_cv is None:
            return {"is_monochrome": 0, "mean_saturation": 0}
        return cf.monochrome(_stats(image_cv, cache), threshold)

    @staticmethod
    def get_dynamic_range(image_cv, cache=None):
        if image_cv is None:
            return {"dynamic_range_stops": 0}
        return cf.dynamic_range(_stats(image_cv, cache))

    @staticmethod
    def get_noise_estimate(image_cv, cache=None):
        if image_cv is None:
            return {"noise_sigma": 0}
        return cf.noise(_stats(image_cv, cache))

    @staticmethod
    def get_contrast_score(image_cv, cache=None):
        if image_cv is None:
            return {"contrast_score": 0, "percentile_contrast": 0, "rms_contrast": 0}
        return cf.contrast(_stats(image_cv, cache))
