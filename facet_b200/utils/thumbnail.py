"""Plan of `generate_photo_thumbnail` (utils/image_transforms.py:32-50 of the reference):
``thumb = pil_img.copy(); thumb.thumbnail((size, size), Image.Resampling.LANCZOS)`` followed by a JPEG save.

Pillow is third-party to the reference (pillow>=10.0.0, requirements.txt) and its source is not under
/root/reference; this module restates the published algorithm of ``Image.thumbnail`` (aspect-preserving
size, ``reducing_gap=2.0``), ``Image.resize`` (integer box reduction by ``int(scale / reducing_gap)`` first),
``ImagingReduce`` (8-bit: ``((sum + n/2) * multiplier) >> 24`` with a float-computed multiplier, separate
edge boxes) and the two-pass 8-bit Lanczos resampler on the reduced image with a fractional source box.
Parity is pinned against the installed Pillow itself (tests/test_thumbnail_plan.py on the CPU,
tests/test_gpu_thumbnail.py through the CUDA kernels).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

from .resample import PRECISION_BITS, precompute_coeffs

LANCZOS_FILTER_SUPPORT = 3.0
REDUCING_GAP = 2.0


def thumbnail_size(height: int, width: int, size: int = 640):
    """Final (height, width) of Image.thumbnail((size, size)); None when the image already fits."""
    x, y = int(math.floor(size)), int(math.floor(size))
    if x >= width and y >= height:
        return None

    def round_aspect(number, key):
        return max(min(math.floor(number), math.ceil(number), key=key), 1)

    aspect = width / height
    if x / y >= aspect:
        x = round_aspect(y * aspect, key=lambda n: abs(aspect - n / y))
    else:
        y = round_aspect(x / aspect, key=lambda n: 0 if n == 0 else abs(aspect - x / n))
    return int(y), int(x)


def reduce_multiplier(count: int) -> int:
    """division_UINT32(count, 8) of Pillow's Reduce.c: (UINT32)(2^32 (as float) / (256 * count)) in float32."""
    return int(np.float32(4294967296.0) / np.float32(256 * count))


@dataclass(frozen=True)
class ThumbnailPlan:
    height: int
    width: int
    out_h: int
    out_w: int
    fx: int                 # box-reduction factors (1 = no reduction)
    fy: int
    red_h: int              # size of the reduced image
    red_w: int
    hbounds: np.ndarray     # Lanczos taps of the horizontal pass over the reduced image
    hcoef: np.ndarray
    hk: int
    vbounds: np.ndarray
    vcoef: np.ndarray
    vk: int


@lru_cache(maxsize=64)
def plan(height: int, width: int, size: int = 640):
    """Tables for thumbnail((size, size), LANCZOS) of an HxW image, or None when Pillow leaves it unchanged."""
    final = thumbnail_size(height, width, size)
    if final is None or final == (height, width):
        return None
    out_h, out_w = final
    fx = int(width / out_w / REDUCING_GAP) or 1
    fy = int(height / out_h / REDUCING_GAP) or 1
    if fx > 1 or fy > 1:
        # _get_safe_box of the full-image box is the full image, so the reduced image covers everything and the
        # resize box is the (fractional) image extent in reduced coordinates
        red_w, red_h = (width + fx - 1) // fx, (height + fy - 1) // fy
        box_w, box_h = width / fx, height / fy
    else:
        red_w, red_h, box_w, box_h = width, height, float(width), float(height)
    hb, hc, hk = precompute_coeffs(red_w, out_w, "lanczos", 0.0, box_w)
    vb, vc, vk = precompute_coeffs(red_h, out_h, "lanczos", 0.0, box_h)
    return ThumbnailPlan(height, width, out_h, out_w, fx, fy, red_h, red_w, hb, hc, hk, vb, vc, vk)


def box_reduce_numpy(img: np.ndarray, fx: int, fy: int) -> np.ndarray:
    """ImagingReduce of an [H,W,C] uint8 image (whole-image box): interior boxes and the narrower edge boxes."""
    h, w, c = img.shape
    oh, ow = (h + fy - 1) // fy, (w + fx - 1) // fx
    out = np.empty((oh, ow, c), np.uint8)
    a = img.astype(np.uint64)
    for (y0, y1, ys) in ((0, h // fy, fy), (h // fy, oh, h % fy)):
        for (x0, x1, xs) in ((0, w // fx, fx), (w // fx, ow, w % fx)):
            if y1 <= y0 or x1 <= x0:
                continue
            yy0, xx0 = y0 * fy, x0 * fx
            blk = a[yy0:yy0 + (y1 - y0) * ys, xx0:xx0 + (x1 - x0) * xs]
            ssum = blk.reshape(y1 - y0, ys, x1 - x0, xs, c).sum(axis=(1, 3))
            n = ys * xs
            out[y0:y1, x0:x1] = (((ssum + n // 2) * reduce_multiplier(n)) & 0xFFFFFFFF) >> 24
    return out


def resample_u8_numpy(img: np.ndarray, hb, hc, vb, vc) -> np.ndarray:
    """Pillow's two-pass 8-bit resampler (horizontal first, uint8 intermediate)."""
    def one_pass(a, bounds, coef, axis):
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        out = np.empty((bounds.shape[0],) + a.shape[1:], np.uint8)
        for i, (first, cnt) in enumerate(bounds):
            acc = np.tensordot(coef[i, :cnt].astype(np.int64), a[first:first + cnt], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
            out[i] = np.clip(acc >> PRECISION_BITS, 0, 255)
        return np.moveaxis(out, 0, axis)
    return one_pass(one_pass(img, hb, hc, 1), vb, vc, 0)


def thumbnail_numpy(img: np.ndarray, size: int = 640) -> np.ndarray:
    """NumPy restatement of Image.thumbnail((size, size), LANCZOS) on an [H,W,3] uint8 image."""
    p = plan(img.shape[0], img.shape[1], size)
    if p is None:
        return img.copy()
    red = box_reduce_numpy(img, p.fx, p.fy) if (p.fx > 1 or p.fy > 1) else img
    return resample_u8_numpy(red, p.hbounds, p.hcoef, p.vbounds, p.vcoef)
