"""Full-size (6000 x 4000 = 24 MP, BASELINE configs[1] / [4]) goldens from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden_24mp.py

Writes tests/golden/technical_24mp_golden.json.  Frames come from integer-only generators
(facet_b200/synth.py: synth_frame_int, extremal_frame), so the GPU box rebuilds byte-identical
inputs.  Per frame: the integer sufficient statistics (hist256, the five sums, sha256 + non-zero
count of the 180x256 H-S histogram) computed with the reference's own cv2 calls
(analyzers/image_cache.py:30-32, analyzers/technical.py:94,153,302), and the seven dicts of
analyzers/technical.py.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from facet_b200.synth import EXTREMAL_KINDS, extremal_frame, synth_frame_int  # noqa: E402

H, W = 4000, 6000
FRAMES = [("int", 0), ("int", 1)] + [(k, 0) for k in EXTREMAL_KINDS]


def make_frame(kind, index, h=H, w=W):
    return synth_frame_int(index, h, w) if kind == "int" else extremal_frame(kind, h, w)


def _clean(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, (bytes, bytearray)):
            out[k] = v.hex()
        elif isinstance(v, (np.floating, np.integer, np.bool_)):
            out[k] = v.item()
        else:
            out[k] = v
    return out


def main():
    import cv2
    from analyzers.image_cache import ImageCache
    from analyzers.technical import TechnicalAnalyzer as TA

    cases = []
    for kind, idx in FRAMES:
        img = make_frame(kind, idx)
        cache = ImageCache(img)
        hs = cv2.calcHist([cache.hsv], [0, 1], None, [180, 256], [0, 180, 0, 256])
        hs_i = hs.astype(np.int64)      # counts < 2^24 are exact in calcHist's float32; 24e6 > 2^24 is checked below
        hs_exact = np.bincount(cache.hsv[:, :, 0].astype(np.int64).ravel() * 256 + cache.hsv[:, :, 1].ravel(), minlength=180 * 256)
        lap = cv2.Laplacian(cache.gray, cv2.CV_64F)
        M = np.array([[1, -2, 1], [-2, 4, -2], [1, -2, 1]])
        nz = cv2.filter2D(cache.gray.astype(np.float64), -1, M)
        rec = {
            "kind": kind, "index": idx, "height": H, "width": W,
            "frame_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
            "laplacian_variance": float(cache.laplacian_variance),
            "hist256": np.bincount(cache.gray.ravel(), minlength=256).tolist(),
            "hs_sha256": hashlib.sha256(hs_exact.astype("<u4").tobytes()).hexdigest(),
            "hs_nonzero": int(np.count_nonzero(hs_exact)),
            "hs_max_bin": int(hs_exact.max()),
            "hs_calchist_matches_exact": bool(np.array_equal(hs_i.ravel(), hs_exact)),
            "sum_saturation": int(cache.hsv[:, :, 1].astype(np.int64).sum()),
            "sum_lap": int(lap.sum()), "sum_lap_sq": int((lap * lap).sum()),
            "sum_abs_noise": int(np.abs(nz).sum()),
            "max_abs_lap": int(np.abs(lap).max()), "max_abs_noise": int(np.abs(nz).max()),
            "sharpness": _clean(TA.get_sharpness_data(img, cache=cache)),
            "color": _clean(TA.get_color_harmony_data(img, cache=cache)),
            "histogram": _clean(TA.get_histogram_data(img, cache=cache)),
            "monochrome": _clean(TA.detect_monochrome(img, threshold=0.10, cache=cache)),
            "dynamic_range": _clean(TA.get_dynamic_range(img, cache=cache)),
            "noise": _clean(TA.get_noise_estimate(img, cache=cache)),
            "contrast": _clean(TA.get_contrast_score(img, cache=cache)),
        }
        print(kind, idx, rec["max_abs_lap"], rec["max_abs_noise"], rec["hs_nonzero"], rec["hs_calchist_matches_exact"], flush=True)
        cases.append(rec)
    meta = {
        "generator": "tests/golden/make_golden_24mp.py",
        "reference": "rlorenzo/facet analyzers/technical.py + analyzers/image_cache.py (unmodified)",
        "versions": {"cv2": cv2.__version__, "numpy": np.__version__, "scipy": __import__("scipy").__version__},
        "cases": cases,
    }
    with open(os.path.join(HERE, "technical_24mp_golden.json"), "w") as f:
        json.dump(meta, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
