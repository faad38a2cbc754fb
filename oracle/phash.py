"""ORACLE (test infrastructure): `imagehash.phash` restated.

imagehash (requirements.txt:25, `imagehash>=4.3.0`) is third-party, not vendored under
/root/reference and not installed in this image; the reference calls it at
processing/batch_processor.py:216, processing/scorer.py:972 and processing/multi_pass.py:449 and has
no test that pins its output: **parity unpinned** against imagehash itself.  Its published algorithm
(imagehash/__init__.py `phash`, hash_size=8, highfreq_factor=4) is reproduced with the very libraries
it calls — Pillow and scipy.fftpack, both installed here and on the GPU box:

    image.convert('L').resize((32, 32), Image.LANCZOS) -> scipy.fftpack.dct(dct(pixels, axis=0), axis=1)
    -> dct[:8, :8] > median -> bits row-major, most significant first -> '%016x'
"""
from __future__ import annotations

import numpy as np


def phash_parts(img_bgr: np.ndarray):
    import scipy.fftpack
    from PIL import Image
    pil = Image.fromarray(np.ascontiguousarray(img_bgr[..., ::-1]))
    small = np.asarray(pil.convert("L").resize((32, 32), Image.LANCZOS))
    dct = scipy.fftpack.dct(scipy.fftpack.dct(small, axis=0), axis=1)
    low = dct[:8, :8]
    med = np.median(low)
    bits = (low > med).flatten()
    value = 0
    for b in bits:
        value = (value << 1) | int(b)
    return small, low, value


def phash_hex(img_bgr: np.ndarray) -> str:
    return "%016x" % phash_parts(img_bgr)[2]
