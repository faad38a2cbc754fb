"""Seeded synthetic inputs for the scoring pass.

There is no dataset and no network: every test, smoke run and bench line uses
images / hashes / embeddings produced here from an integer seed, so the CUDA
path, the oracle and the reference see byte-identical inputs.

Image layout follows ``utils/image_loading.py:106`` of the reference: an
``[H, W, 3]`` uint8 array in **BGR** channel order, C-contiguous.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "synth_image_bgr",
    "synth_hashes",
    "synth_embeddings",
    "synth_timestamps",
    "IMAGE_KINDS",
]

# index % 8 selects the branch of the reference the frame is meant to exercise
# (technical.py:165-177 clipping / silhouette, technical.py:239 monochrome).
IMAGE_KINDS = ("mono", "lowkey", "highkey", "backlit", "plain", "plain", "noise", "plain")


def _smooth_field(rng: np.random.Generator, h: int, w: int, octaves: int = 3) -> np.ndarray:
    """Band-limited field in [0,1]: a few random low-frequency cosines."""
    yy = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None]
    xx = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :]
    f = np.zeros((h, w), np.float32)
    for o in range(octaves):
        fx, fy = rng.uniform(0.5, 3.0 * (o + 1), size=2)
        ph = rng.uniform(0, 2 * np.pi, size=2)
        f += (np.cos(2 * np.pi * fx * xx + ph[0]) * np.cos(2 * np.pi * fy * yy + ph[1])) / (o + 1)
    f -= f.min()
    f /= max(float(f.max()), 1e-6)
    return f


def synth_image_bgr(index: int, height: int, width: int, seed: int = 1000) -> np.ndarray:
    """Deterministic ``[H,W,3]`` uint8 BGR frame number ``index``.

    gradient + band-limited texture + filled rectangles + sensor-like noise,
    then a per-kind tone curve (see IMAGE_KINDS).
    """
    rng = np.random.default_rng(seed + index)
    kind = IMAGE_KINDS[index % 8]
    h, w = int(height), int(width)
    if kind == "noise":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)

    base = _smooth_field(rng, h, w)
    img = np.empty((h, w, 3), np.float32)
    tint = rng.uniform(0.55, 1.0, size=3).astype(np.float32)
    for c in range(3):
        img[:, :, c] = base * tint[c] + 0.25 * _smooth_field(rng, h, w, octaves=2)
    # filled rectangles (hard edges feed the Laplacian / Immerkaer sums)
    for _ in range(int(rng.integers(3, 9))):
        y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        y1 = min(h, y0 + int(rng.integers(1, max(2, h // 3))))
        x1 = min(w, x0 + int(rng.integers(1, max(2, w // 3))))
        img[y0:y1, x0:x1, :] = rng.uniform(0.0, 1.0, size=3).astype(np.float32)
    img /= max(float(img.max()), 1e-6)

    if kind == "lowkey":
        img = img ** 2.6 * 0.55
    elif kind == "highkey":
        img = 1.0 - (1.0 - img) ** 2.4 * 0.6
    elif kind == "backlit":
        mask = _smooth_field(rng, h, w, octaves=1) > 0.5
        img = np.where(mask[:, :, None], 0.04 + 0.08 * img, 0.82 + 0.18 * img)
    else:
        gain, gamma = rng.uniform(0.8, 1.05), rng.uniform(0.7, 1.4)
        img = np.clip(img * gain, 0, 1) ** gamma

    sigma = rng.uniform(0.5, 6.0)
    noise = rng.standard_normal(size=(h, w, 3), dtype=np.float32) * (sigma / 255.0)
    out = np.clip((img + noise) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    if kind == "mono":
        g = out[:, :, 1].copy()
        out[:, :, 0] = g
        out[:, :, 2] = g
    return np.ascontiguousarray(out)


def synth_hashes(n: int, seed: int = 11, dup_fraction: float = 0.2, max_flip: int = 8) -> np.ndarray:
    """``n`` uint64 perceptual hashes with planted near-duplicates.

    A ``dup_fraction`` of the entries are copies of an earlier entry with
    0..max_flip random bits flipped, so both sides of the ``<= 6`` threshold of
    ``utils/duplicate.py:61-63`` are populated.
    """
    rng = np.random.default_rng(seed)
    hs = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    n_dup = int(n * dup_fraction)
    if n > 1 and n_dup:
        dst = rng.choice(np.arange(1, n), size=min(n_dup, n - 1), replace=False)
        for d in np.sort(dst):
            src = int(rng.integers(0, d))
            v = int(hs[src])
            for b in rng.choice(64, size=int(rng.integers(0, max_flip + 1)), replace=False):
                v ^= 1 << int(b)
            hs[d] = np.uint64(v)
    return hs


def synth_timestamps(n: int, seed: int = 13) -> np.ndarray:
    """Sorted integer second offsets with bursty gaps (0 s .. 120 s)."""
    rng = np.random.default_rng(seed)
    gaps = rng.choice(np.array([0, 0, 1, 1, 2, 5, 20, 47, 48, 49, 90, 120]), size=n)
    gaps[0] = 0
    return np.cumsum(gaps).astype(np.int64)


def synth_embeddings(n: int, dim: int = 768, seed: int = 11, cluster_fraction: float = 0.15,
                     jitter: float = 0.012) -> np.ndarray:
    """``[n, dim]`` float32 L2-normalised embeddings with planted near-duplicate clusters.

    With jitter 0.012 per coordinate (|noise| ~ 0.33 at dim 768) a copy sits at
    cosine ~0.95 from its source; unrelated rows sit near 0.
    """
    rng = np.random.default_rng(seed)
    e = rng.standard_normal(size=(n, dim), dtype=np.float32)
    e /= np.linalg.norm(e, axis=1, keepdims=True)
    n_dup = int(n * cluster_fraction)
    if n > 1 and n_dup:
        dst = np.sort(rng.choice(np.arange(1, n), size=min(n_dup, n - 1), replace=False))
        src = (rng.random(len(dst)) * dst).astype(np.int64)
        scale = rng.uniform(0.2, 2.0, size=(len(dst), 1)).astype(np.float32)
        for d, s, sc in zip(dst, src, scale):
            v = e[s] + jitter * sc * rng.standard_normal(dim).astype(np.float32)
            e[d] = v / np.linalg.norm(v)
    return np.ascontiguousarray(e)


# ----------------------------------------------------------------------------------------------------
# Integer-only frames (bit-identical on every host: no floating point in the generator) for the
# full-size (24 MP) parity goldens, and the extremal frames that reach the stated maxima of the
# technical kernel's exactness arguments (|Laplacian| = 1020, |Immerkaer response| = 4080 per pixel).
# ----------------------------------------------------------------------------------------------------
EXTREMAL_KINDS = ("checker", "black", "white", "red", "green", "blue", "vstripes", "hstripes", "checker2", "halfplane")


def synth_frame_int(index: int, height: int, width: int, seed: int = 5000) -> np.ndarray:
    """Deterministic ``[H,W,3]`` uint8 BGR frame built from integer arithmetic only: per-channel planar
    gradients, a coarse block pattern, filled rectangles and uniform integer noise."""
    rng = np.random.default_rng(seed + index)
    h, w = int(height), int(width)
    yy = np.arange(h, dtype=np.int64)[:, None]
    xx = np.arange(w, dtype=np.int64)[None, :]
    out = np.empty((h, w, 3), np.uint8)
    amp = int(rng.integers(2, 24))
    blocks = rng.integers(0, 64, size=(h // 64 + 1, w // 64 + 1, 3), dtype=np.int64)
    for c in range(3):
        a, b, o = (int(v) for v in rng.integers(0, 256, size=3))
        plane = (o + (a * xx) // max(w, 1) + (b * yy) // max(h, 1)) % 256
        plane = plane + blocks[:, :, c].repeat(64, axis=0)[:h].repeat(64, axis=1)[:, :w]
        plane = plane + rng.integers(-amp, amp + 1, size=(h, w), dtype=np.int64)
        out[:, :, c] = np.clip(plane, 0, 255).astype(np.uint8)
    for _ in range(int(rng.integers(3, 9))):
        y0, x0 = int(rng.integers(0, h)), int(rng.integers(0, w))
        y1 = min(h, y0 + int(rng.integers(1, max(2, h // 3))))
        x1 = min(w, x0 + int(rng.integers(1, max(2, w // 3))))
        out[y0:y1, x0:x1, :] = rng.integers(0, 256, size=3, dtype=np.int64).astype(np.uint8)
    return np.ascontiguousarray(out)


def extremal_frame(kind: str, height: int, width: int) -> np.ndarray:
    """Frames that sit on the extremes of the per-pixel responses (BGR uint8)."""
    h, w = int(height), int(width)
    yy = np.arange(h)[:, None]
    xx = np.arange(w)[None, :]
    out = np.zeros((h, w, 3), np.uint8)
    if kind == "checker":          # 0 / 255 at one-pixel pitch: |L| = 1020 and |N| = 4080 in the interior
        out[:] = (((yy + xx) & 1) * 255).astype(np.uint8)[:, :, None]
    elif kind == "checker2":       # two-pixel pitch, saturated colours alternating with their complements
        m = (((yy >> 1) + (xx >> 1)) & 1).astype(bool)
        out[:] = np.where(m[:, :, None], np.array([255, 0, 255], np.uint8), np.array([0, 255, 0], np.uint8))
    elif kind == "black":
        pass
    elif kind == "white":
        out[:] = 255
    elif kind == "blue":
        out[:, :, 0] = 255
    elif kind == "green":
        out[:, :, 1] = 255
    elif kind == "red":
        out[:, :, 2] = 255
    elif kind == "vstripes":       # one-pixel vertical stripes
        out[:] = ((xx & 1) * 255).astype(np.uint8)[:, :, None] * np.ones((h, 1, 1), np.uint8)
    elif kind == "hstripes":
        out[:] = ((yy & 1) * 255).astype(np.uint8)[:, :, None] * np.ones((1, w, 1), np.uint8)
    elif kind == "halfplane":      # silhouette-like: left half black, right half white, one hard edge
        out[:, w // 2:, :] = 255
    else:
        raise ValueError(f"unknown extremal frame kind {kind!r}")
    return np.ascontiguousarray(out)
