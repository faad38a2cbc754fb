"""Generate golden vectors by running the UNMODIFIED reference modules.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Writes tests/golden/technical_golden.json: for each seeded synthetic frame the
exact dicts returned by analyzers/image_cache.py + analyzers/technical.py, in
the call order of processing/batch_processor.py:198-233.  The GPU box has no
/root/reference; tests there compare against this file.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from facet_b200.synth import synth_image_bgr  # noqa: E402

# (index, height, width): odd sizes, tiny sizes, the config-1 size
CASES = [(i, 128, 192) for i in range(8)] + [
    (8, 97, 131), (9, 2, 2), (10, 3, 517), (11, 260, 5), (12, 683, 1024), (13, 683, 1024),
    (14, 1024, 683), (15, 171, 256), (16, 64, 64), (19, 333, 500), (22, 96, 96),
]


def _clean(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, (bytes, bytearray)):
            out[k] = v.hex()
        elif isinstance(v, (np.floating, np.integer)):
            out[k] = v.item()
        else:
            out[k] = v
    return out


def main():
    import cv2
    from analyzers.image_cache import ImageCache
    from analyzers.technical import TechnicalAnalyzer as TA

    cases = []
    for idx, h, w in CASES:
        img = synth_image_bgr(idx, h, w)
        cache = ImageCache(img)
        hs = cv2.calcHist([cache.hsv], [0, 1], None, [180, 256], [0, 180, 0, 256])
        nzb = np.flatnonzero(hs.ravel())
        lap = cv2.Laplacian(cache.gray, cv2.CV_64F)
        M = np.array([[1, -2, 1], [-2, 4, -2], [1, -2, 1]])
        nz = cv2.filter2D(cache.gray.astype(np.float64), -1, M)
        rec = {
            "index": idx, "height": h, "width": w,
            "laplacian_variance": float(cache.laplacian_variance),
            "hist256": np.bincount(cache.gray.ravel(), minlength=256).tolist(),
            "hs_nonzero_bins": nzb.tolist(),
            "hs_nonzero_counts": hs.ravel()[nzb].astype(np.int64).tolist(),
            "sum_lap": int(lap.sum()), "sum_lap_sq": int((lap * lap).sum()),
            "sum_abs_noise": int(np.abs(nz).sum()),
            "sharpness": _clean(TA.get_sharpness_data(img, cache=cache)),
            "color": _clean(TA.get_color_harmony_data(img, cache=cache)),
            "histogram": _clean(TA.get_histogram_data(img, cache=cache)),
            "monochrome": _clean(TA.detect_monochrome(img, threshold=0.10, cache=cache)),
            "dynamic_range": _clean(TA.get_dynamic_range(img, cache=cache)),
            "noise": _clean(TA.get_noise_estimate(img, cache=cache)),
            "contrast": _clean(TA.get_contrast_score(img, cache=cache)),
        }
        cases.append(rec)
    meta = {
        "generator": "tests/golden/make_golden.py",
        "reference": "rlorenzo/facet analyzers/technical.py + analyzers/image_cache.py (unmodified)",
        "versions": {"cv2": cv2.__version__, "numpy": np.__version__,
                     "scipy": __import__("scipy").__version__},
        "cases": cases,
    }
    with open(os.path.join(HERE, "technical_golden.json"), "w") as f:
        json.dump(meta, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
