"""Thumbnail generation with the reference's interface (utils/image_transforms.py:32-50).

`generate_photo_thumbnail(img, size=640, quality=80) -> bytes` returns the same JPEG bytes as the reference:
the pixels come from the CUDA kernels (csrc/thumbnail.cu, bit-exact with Pillow's `thumbnail`) and so does the JPEG stream
(csrc/jpeg_encode.cu, byte-exact with Pillow's encoder at the same quality).  `generate_photo_thumbnails` is the batched form used by the pipeline.
"""
from __future__ import annotations

import numpy as np

from .. import ops


def generate_photo_thumbnails(images, size: int = 640, quality: int = 80, rgb_order: bool = False) -> list[bytes]:
    """images: uint8 [n,H,W,3] (BGR as utils/image_loading.py:106 hands them over, unless rgb_order) on the host
    or the device -> one JPEG per image."""
    thumbs = ops.thumbnails(images, size=size, rgb_order=rgb_order, to_rgb=True)
    return ops.jpeg_encode(thumbs, quality=quality)          # csrc/jpeg_encode.cu: the bytes Pillow's encoder would write


def generate_photo_thumbnail(pil_img, size: int = 640, quality: int = 80) -> bytes:
    """Reference signature: a PIL image (RGB) or an [H,W,3] RGB array -> JPEG bytes."""
    rgb = np.asarray(pil_img.convert("RGB") if hasattr(pil_img, "convert") else pil_img, dtype=np.uint8)
    return generate_photo_thumbnails(rgb[None], size=size, quality=quality, rgb_order=True)[0]
