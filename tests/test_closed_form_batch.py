"""The batched closed forms return exactly the dicts of the per-image functions (which are pinned against the
reference through tests/golden/technical_golden.json)."""
import math

import numpy as np

from facet_b200.analyzers import _closed_form as cf


def _same(a, b):
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        return True
    return type(a) is type(b) and a == b


def _hists(rng, h, w):
    npx = h * w
    out = []
    for k in range(40):
        alpha = rng.dirichlet(np.ones(256) * [0.02, 0.2, 1.0, 5.0][k % 4])
        out.append(rng.multinomial(npx, alpha))
    one = np.zeros(256, np.int64); one[17] = npx; out.append(one)                      # a flat frame
    two = np.zeros(256, np.int64); two[0] = npx // 2; two[255] = npx - npx // 2; out.append(two)
    dark = np.zeros(256, np.int64); dark[0] = npx; out.append(dark)
    ramp = np.full(256, npx // 256, np.int64); ramp[-1] += npx - ramp.sum(); out.append(ramp)
    sil = np.zeros(256, np.int64); sil[10] = int(npx * 0.6); sil[240] = npx - sil[10]; out.append(sil)
    return np.array(out, dtype=np.int64)


def test_batch_equals_scalar_bit_for_bit():
    rng = np.random.default_rng(0)
    for (h, w) in [(4000, 6000), (683, 1024), (2, 2), (37, 53)]:
        hists = _hists(rng, h, w)
        n = len(hists)
        sl = rng.integers(-10**6, 10**6, n)
        sq = rng.integers(0, 10**13, n)
        sn = rng.integers(0, 10**10, n)
        ent = rng.uniform(0, 15.5, n)
        ss = rng.uniform(0, 255.0 * h * w, n)
        got = cf.all_metrics_batch(h, w, hists, sl, sq, sn, ent, ss, mono_threshold=0.1)
        for i in range(n):
            st = cf.TechStats(h, w, hists[i], int(sl[i]), int(sq[i]), int(sn[i]), float(ent[i]), float(ss[i]))
            want = {"sharpness": cf.sharpness(st), "color": cf.color_harmony(st), "histogram": cf.histogram(st),
                    "monochrome": cf.monochrome(st, 0.1), "dynamic_range": cf.dynamic_range(st), "noise": cf.noise(st),
                    "contrast": cf.contrast(st)}
            assert _same(got[i], want), (h, w, i, got[i], want)
    assert cf.all_metrics_batch(8, 8, np.zeros((0, 256), np.int64), [], [], [], [], []) == []
