"""Coefficient tables for the CLIP preprocess kernels.

`scorer.preprocess` in the reference (processing/scorer.py:508-510) is open_clip's inference
transform: torchvision ``Resize(224, BICUBIC)`` (shorter side, on a PIL image, hence Pillow's
antialiased resampler) -> ``CenterCrop(224)`` -> ``ToTensor`` -> ``Normalize``.  Pillow and
torchvision are third-party to the reference (pillow>=10.0.0, torchvision; requirements.txt)
and their sources are not under /root/reference; this module restates Pillow's published
algorithm (src/libImaging/Resample.c: ``precompute_coeffs`` + ``normalize_coeffs_8bpc``) in
float64, tap by tap, so the integer taps are identical.  Parity is pinned in
tests/test_preprocess.py against the installed Pillow/torchvision themselves.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

PRECISION_BITS = 32 - 8 - 2
BICUBIC_SUPPORT = 2.0

# open_clip's pretrained config for ViT-L-14 / laion2b_s32b_b82k; OpenAI CLIP
# statistics are kept for callers that load OpenAI weights.
LAION_MEAN = (0.5, 0.5, 0.5)
LAION_STD = (0.5, 0.5, 0.5)
OPENAI_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def _lanczos(x: np.ndarray) -> np.ndarray:
    """Pillow lanczos_filter (support 3): sinc(x) * sinc(x / 3) on [-3, 3)."""
    def sinc(v):
        out = np.ones_like(v)
        nz = v != 0.0
        out[nz] = np.sin(np.pi * v[nz]) / (np.pi * v[nz])
        return out
    x = np.asarray(x, dtype=np.float64)
    return np.where((x >= -3.0) & (x < 3.0), sinc(x) * sinc(x / 3.0), 0.0)


LANCZOS_SUPPORT = 3.0


def precompute_coeffs(in_size: int, out_size: int, kernel: str = "bicubic", in0: float = 0.0, in1: float | None = None):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the box (in0, in1) of an axis of in_size pixels
    (default: the whole axis), bicubic or lanczos.  The box ends are C floats, as in ImagingResample.

    Returns (bounds int32 [out,2] = first tap / tap count, coef int32 [out,ksize], ksize).
    """
    filt, base_support = (_bicubic, BICUBIC_SUPPORT) if kernel == "bicubic" else (_lanczos, LANCZOS_SUPPORT)
    in0 = float(np.float32(in0))
    in1 = float(np.float32(in_size if in1 is None else in1))
    scale = (in1 - in0) / out_size
    filterscale = max(scale, 1.0)
    support = base_support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        cnt = xmax - xmin
        w = filt((np.arange(cnt, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:            # sequential double accumulation, as in C
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        kk[xx, :cnt] = w
        bounds[xx] = (xmin, cnt)
    scaled = kk * float(1 << PRECISION_BITS)
    coef = np.where(kk < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int32)
    return bounds, coef, ksize


def resized_size(height: int, width: int, size: int):
    """torchvision Resize(int): shorter side -> size, longer side int(size*long/short)."""
    short, long_ = (width, height) if width <= height else (height, width)
    new_short, new_long = size, int(size * long_ / short)
    return (new_long, new_short) if width <= height else (new_short, new_long)   # (new_h, new_w)


@dataclass(frozen=True)
class ResamplePlan:
    height: int
    width: int
    out: int
    hbounds: np.ndarray
    hcoef: np.ndarray
    hk: int
    hp0: np.ndarray        # int32 [out]: first tap of each column rounded down to a multiple of 4 pixels
    hcpad: np.ndarray      # int32 [4*hgroups, out]: coefficient of pixel hp0[xo]+i, zero padded
    hgroups: int
    h_px_lo: int           # staged pixel range of a row (multiples of 16)
    h_span_px: int
    vbounds: np.ndarray
    vcoef: np.ndarray
    vk: int
    row0: int
    rows: int
    # tensor-core form of the horizontal pass (csrc/resample_tc.cu): per block of 8 output columns an
    # int8 matrix [96][tc_kw]: row = limb*24 + (xo - 8j)*3 + channel, column = byte - tc_kb0[j]
    tc_coef: np.ndarray | None = None    # int8 [out/8 * 96, tc_kw]
    tc_kb0: np.ndarray | None = None     # int32 [out/8]
    tc_kw: int = 0
    tc_limbs: int = 0


@lru_cache(maxsize=64)
def plan(height: int, width: int, out: int = 224) -> ResamplePlan:
    """Tables for Resize(out) + CenterCrop(out) of an HxW image, cropped to the kept outputs."""
    new_h, new_w = resized_size(height, width, out)
    if new_h < out or new_w < out:
        raise ValueError("image too small for the crop")
    top = int(round((new_h - out) / 2.0))
    left = int(round((new_w - out) / 2.0))
    hb, hc, hk = precompute_coeffs(width, new_w)
    vb, vc, vk = precompute_coeffs(height, new_h)
    hb, hc = hb[left:left + out].copy(), hc[left:left + out].copy()
    vb, vc = vb[top:top + out].copy(), vc[top:top + out].copy()
    row0 = int(vb[:, 0].min())
    row1 = int((vb[:, 0] + vb[:, 1]).max())
    # horizontal taps re-based to 4-pixel (12-byte, word-aligned) groups for the kernel
    hp0 = (hb[:, 0] & ~3).astype(np.int32)
    hgroups = int(np.max((hb[:, 0] + hb[:, 1] - hp0 + 3) // 4))
    hcpad = np.zeros((4 * hgroups, out), np.int32)
    for xo in range(out):
        off = int(hb[xo, 0] - hp0[xo])
        cnt = int(hb[xo, 1])
        hcpad[off:off + cnt, xo] = hc[xo, :cnt]
    px_lo = int(hp0.min()) // 16 * 16
    px_hi = -(-int((hp0 + 4 * hgroups).max()) // 16) * 16
    tc_coef, tc_kb0, tc_kw, tc_limbs = _tc_tables(hb, hc, out)
    return ResamplePlan(height, width, out, np.ascontiguousarray(hb), np.ascontiguousarray(hc), hk,
                        np.ascontiguousarray(hp0), np.ascontiguousarray(hcpad), hgroups, px_lo, px_hi - px_lo,
                        np.ascontiguousarray(vb), np.ascontiguousarray(vc), vk, row0, row1 - row0,
                        tc_coef, tc_kb0, tc_kw, tc_limbs)


def split_limbs(c: np.ndarray, limbs: int) -> np.ndarray:
    """Signed base-128 digits of int coefficients: c == sum_i d_i * 128**i, d_i in [-64, 63] except the top one."""
    c = c.astype(np.int64).copy()
    out = []
    for i in range(limbs - 1):
        d = ((c + 64) % 128) - 64
        out.append(d)
        c = (c - d) // 128
    out.append(c)
    return np.stack(out)


def _tc_tables(hb, hc, out, nb: int = 8, n_rows: int = 96, channels: int = 3):
    """Banded int8 coefficient matrices for the tcgen05 horizontal pass (None if the layout does not fit)."""
    if out % nb:
        return None, None, 0, 0
    nblk = out // nb
    first = hb[:, 0].astype(np.int64)
    last = (hb[:, 0] + hb[:, 1]).astype(np.int64)
    kb0 = np.array([int(first[nb * j]) * channels // 16 * 16 for j in range(nblk)], np.int32)
    span = max(int(last[nb * j:nb * j + nb].max()) * channels - int(kb0[j]) for j in range(nblk))
    kw = -(-span // 128) * 128
    cmax = int(np.abs(hc).max())
    limbs = 3 if cmax < 63 * 16384 else 4
    if np.abs(split_limbs(hc, limbs)[-1]).max() > 127 or nb * channels * limbs > n_rows:
        return None, None, 0, 0
    digits = split_limbs(hc, limbs)                      # [limbs, out, ksize]
    table = np.zeros((nblk, n_rows, kw), np.int8)
    for j in range(nblk):
        for xl in range(nb):
            xo = nb * j + xl
            f, cnt = int(hb[xo, 0]), int(hb[xo, 1])
            for c in range(channels):
                cols = (f + np.arange(cnt)) * channels + c - int(kb0[j])
                for L in range(limbs):
                    table[j, L * nb * channels + xl * channels + c, cols] = digits[L, xo, :cnt]
    return np.ascontiguousarray(table.reshape(nblk * n_rows, kw)), kb0, kw, limbs


def resample_reference_numpy(img_rgb: np.ndarray, out: int = 224) -> np.ndarray:
    """Host evaluation of the same integer plan (used by CPU tests of the tables, not by the
    product path): returns the uint8 [out,out,3] crop Pillow would produce."""
    p = plan(img_rgb.shape[0], img_rgb.shape[1], out)
    src = img_rgb.astype(np.int64)
    tmp = np.zeros((p.rows, out, 3), np.uint8)
    for xo in range(out):
        f, c = p.hbounds[xo]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[p.row0:p.row0 + p.rows, f:f + c, :],
                                                         p.hcoef[xo, :c].astype(np.int64), axes=([1], [0]))
        tmp[:, xo, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    res = np.zeros((out, out, 3), np.uint8)
    t64 = tmp.astype(np.int64)
    for yo in range(out):
        f, c = p.vbounds[yo]
        f -= p.row0
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(p.vcoef[yo, :c].astype(np.int64), t64[f:f + c], axes=([0], [0]))
        res[yo] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return res


@lru_cache(maxsize=64)
def phash_plan(height: int, width: int, size: int = 32):
    """Lanczos tables of imagehash.phash's `resize((32, 32), ANTIALIAS)` for an HxW luma image:
    (hbounds, hcoef, hk, vbounds, vcoef, vk)."""
    hb, hc, hk = precompute_coeffs(width, size, "lanczos")
    vb, vc, vk = precompute_coeffs(height, size, "lanczos")
    return (np.ascontiguousarray(hb), np.ascontiguousarray(hc), hk, np.ascontiguousarray(vb), np.ascontiguousarray(vc), vk)


@lru_cache(maxsize=64)
def phash_tc_tables(height: int, width: int, size: int = 32):
    """int8 limb tables of the horizontal Lanczos pass on the luma plane (one channel, 32 rows per block)."""
    hb, hc, _ = precompute_coeffs(width, size, "lanczos")
    return _tc_tables(hb, hc, size, nb=8, n_rows=32, channels=1)
