"""Golden vectors of the rule-based composition analyzer from the UNMODIFIED reference.

    python tests/golden/make_golden_composition.py        (build container: /root/reference must exist)

Writes tests/golden/composition_golden.json: for seeded synthetic frames the dicts / boxes returned by
analyzers/composition.py `detect_leading_lines`, `detect_subject_region`, `get_placement_data(None, w, h, None,
img_cv)` and `integrate_leading_lines`, plus the sha256 of the two edge maps the reference computes on the way
(cv2.Canny of the blurred plane at 50 / 150; cv2.Canny of the gray plane at the median-derived thresholds).
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

import cv2  # noqa: E402
from analyzers.composition import CompositionAnalyzer  # noqa: E402
from analyzers.image_cache import ImageCache  # noqa: E402

from facet_b200.synth import synth_image_bgr  # noqa: E402

CASES = [(i, 683, 1024) for i in range(8)] + [(8, 97, 131), (9, 2, 2), (10, 3, 517), (11, 260, 5), (12, 1024, 683),
                                                (13, 333, 500), (14, 1200, 1800), (20, 2000, 3000), (4, 4000, 6000)]


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    return v


def main():
    out = {"versions": {"cv2": cv2.__version__, "numpy": np.__version__}, "cases": []}
    for idx, h, w in CASES:
        img = synth_image_bgr(idx, h, w)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        med = np.median(gray)
        lower, upper = int(max(0, 0.5 * med)), int(min(255, 1.5 * med))
        e_lines = cv2.Canny(cv2.GaussianBlur(gray, (5, 5), 0), 50, 150)
        e_subj = cv2.Canny(gray, lower, upper)
        lead = CompositionAnalyzer.detect_leading_lines(img, cache=ImageCache(img))
        lead_nocache = CompositionAnalyzer.detect_leading_lines(img)
        assert lead == lead_nocache
        bbox = CompositionAnalyzer.detect_subject_region(img)
        place = CompositionAnalyzer.get_placement_data(None, w, h, None, img_cv=img)
        out["cases"].append({
            "index": idx, "height": h, "width": w, "median": float(med), "lower": lower, "upper": upper,
            "edges_lines_sha256": hashlib.sha256(e_lines.tobytes()).hexdigest(), "edges_lines_count": int((e_lines > 0).sum()),
            "edges_subject_sha256": hashlib.sha256(e_subj.tobytes()).hexdigest(), "edges_subject_count": int((e_subj > 0).sum()),
            "leading_lines": _plain(lead), "subject_bbox": _plain(bbox), "placement": _plain(place),
            "integrated": _plain(CompositionAnalyzer.integrate_leading_lines(place["score"], lead["leading_lines_score"], False)),
        })
        print(idx, h, w, lead, bbox, place)
    with open(os.path.join(HERE, "composition_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
