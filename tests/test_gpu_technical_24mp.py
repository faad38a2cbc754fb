"""GPU parity at BASELINE's full frame size (6000 x 4000): the technical pass through the C ABI vs goldens of
the UNMODIFIED reference (tests/golden/make_golden_24mp.py) on two integer-synthesised frames and on the
extremal frames that reach the maxima behind the kernel's exactness arguments (csrc/tech_stats.cu: per-row
fp32 sums, fp16 stencil values): 0/255 checkerboard (|L| = 1020, |N| = 2040 everywhere), all-0, all-255,
saturated primaries, one-pixel stripes, complementary 2-px checker, a half plane."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, approx_rel
from facet_b200.synth import extremal_frame, synth_frame_int

pytestmark = pytest.mark.gpu

INT_KEYS = ("shadow_clipped", "highlight_clipped", "is_silhouette", "is_monochrome")

with open(os.path.join(GOLDEN_DIR, "technical_24mp_golden.json")) as _f:
    CASES = json.load(_f)["cases"]


def _frame(rec):
    h, w = rec["height"], rec["width"]
    return synth_frame_int(rec["index"], h, w) if rec["kind"] == "int" else extremal_frame(rec["kind"], h, w)


@pytest.mark.parametrize("rec", CASES, ids=[f'{c["kind"]}{c["index"]}' for c in CASES])
def test_24mp_frame_matches_reference_golden(rec):
    from facet_b200 import ops
    from facet_b200.analyzers import ImageCache, TechnicalAnalyzer as TA
    img = _frame(rec)
    assert hashlib.sha256(img.tobytes()).hexdigest() == rec["frame_sha256"], "frame generator is not reproducible on this host"
    st = ops.tech_stats(img, want_hs=True)[0]
    cache = ImageCache(img, stats=st)
    assert st.hist256.tolist() == rec["hist256"]
    assert (st.sum_lap, st.sum_lap_sq, st.sum_abs_noise) == (rec["sum_lap"], rec["sum_lap_sq"], rec["sum_abs_noise"])
    hs = np.ascontiguousarray(st.hs_hist).astype("<u4")
    assert int(np.count_nonzero(hs)) == rec["hs_nonzero"] and int(hs.max()) == rec["hs_max_bin"]
    assert hashlib.sha256(hs.tobytes()).hexdigest() == rec["hs_sha256"], "H-S histogram differs"
    assert st.sum_saturation == float(rec["sum_saturation"])
    assert approx_rel(cache.laplacian_variance, rec["laplacian_variance"], rel=1e-9, abs_=1e-9)
    got = {
        "sharpness": TA.get_sharpness_data(img, cache=cache),
        "color": TA.get_color_harmony_data(img, cache=cache),
        "histogram": TA.get_histogram_data(img, cache=cache),
        "monochrome": TA.detect_monochrome(img, threshold=0.10, cache=cache),
        "dynamic_range": TA.get_dynamic_range(img, cache=cache),
        "noise": TA.get_noise_estimate(img, cache=cache),
        "contrast": TA.get_contrast_score(img, cache=cache),
    }
    for name in got:
        for key, w in rec[name].items():
            g = got[name][key]
            if key == "histogram_bytes":
                assert g.hex() == w, "histogram_bytes must be bit-exact"
            elif key in INT_KEYS:
                assert int(g) == int(w), (name, key)
            else:
                assert approx_rel(float(g), float(w), rel=1e-3, abs_=5e-5), (rec["kind"], name, key, g, w)


def test_24mp_batch_of_extremal_frames_one_launch():
    """All extremal frames in ONE launch (a CTA's unit range then crosses image boundaries between frames whose
    histograms are single bins) + both kernels (fast and generic) on the checkerboard."""
    from facet_b200 import ops
    recs = [c for c in CASES if c["kind"] != "int"][:6]
    imgs = np.stack([_frame(r) for r in recs])
    for st, rec in zip(ops.tech_stats(imgs, want_hs=True), recs):
        assert st.hist256.tolist() == rec["hist256"]
        assert (st.sum_lap, st.sum_lap_sq, st.sum_abs_noise) == (rec["sum_lap"], rec["sum_lap_sq"], rec["sum_abs_noise"])
        assert hashlib.sha256(np.ascontiguousarray(st.hs_hist).astype("<u4").tobytes()).hexdigest() == rec["hs_sha256"]
    chk = next(c for c in CASES if c["kind"] == "checker")
    g = ops.tech_stats(imgs[:1], want_hs=False, force_generic=True)[0]
    assert (g.sum_lap, g.sum_lap_sq, g.sum_abs_noise) == (chk["sum_lap"], chk["sum_lap_sq"], chk["sum_abs_noise"])
