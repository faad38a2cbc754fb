"""GPU-backed ``ImageCache`` (mirrors analyzers/image_cache.py:8-32 of the reference).

The reference computes gray, HSV and the Laplacian variance on the CPU when the cache is
constructed.  Here construction launches the single technical pass on the GPU and keeps the
sufficient statistics (``stats``); ``laplacian_variance``, ``height`` and ``width`` are plain
attributes as in the reference.  ``gray`` / ``hsv`` are not needed by any technical metric any
more, so they are materialised lazily (by the same integer formulas, on the GPU) only if a
caller such as the composition analyzer asks for them.
"""
from __future__ import annotations

import numpy as np

from .. import ops
from . import _closed_form as cf


class ImageCache:
    __slots__ = ["_img", "_gray", "_hsv", "laplacian_variance", "height", "width", "stats", "_rgb"]

    def __init__(self, img_cv, *, rgb_order: bool = False, stats: cf.TechStats | None = None):
        """img_cv: BGR uint8 array [H,W,3] (numpy, as the reference passes) or a CUDA uint8 tensor."""
        self._img = img_cv
        self._rgb = rgb_order
        self.height, self.width = int(img_cv.shape[0]), int(img_cv.shape[1])
        self.stats = stats if stats is not None else ops.tech_stats(img_cv, rgb_order=rgb_order)[0]
        self.laplacian_variance = cf.laplacian_variance(self.stats)
        self._gray = None
        self._hsv = None

    @classmethod
    def from_batch(cls, images, *, rgb_order: bool = False):
        """One GPU launch for a same-shaped batch; returns a list of caches."""
        stats = ops.tech_stats(images, rgb_order=rgb_order)
        return [cls(images[i], rgb_order=rgb_order, stats=s) for i, s in enumerate(stats)]

    def _planes(self):
        if self._gray is None:
            self._gray, self._hsv = ops.gray_hsv_planes(self._img, rgb_order=self._rgb)
        return self._gray, self._hsv

    @property
    def gray(self) -> np.ndarray:
        return self._planes()[0]

    @property
    def hsv(self) -> np.ndarray:
        return self._planes()[1]
