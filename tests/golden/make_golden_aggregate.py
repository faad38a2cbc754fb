"""Golden vectors for the aggregate score / category, produced by the UNMODIFIED reference
(`Facet(lightweight=True).calculate_aggregate_logic`, processing/scorer.py:769) on seeded random metric dicts.

    python tests/golden/make_golden_aggregate.py   ->  tests/golden/aggregate_golden.json

Two configs: the scoring sections of the reference's own scoring_config.json (tag prompt lists dropped: they do
not enter the aggregate), and a mutated copy (no penalties / exif / exposure sections, silhouette detection off,
weights that do not sum to 100, a duplicated category name, custom limits, a default category without an entry).
"""
import copy
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

SECTIONS = ("categories", "scoring", "thresholds", "composition", "exif_adjustments", "exposure", "penalties", "viewer",
            "monochrome_detection", "tagging", "burst_detection", "duplicate_detection")
TAGS = ["portrait", "landscape", "mountain", "concert", "street", "candid", "animal", "food", "macro", "flower",
        "architecture", "city", "silhouette", "group", "painting", "fog", "Sunset", " sky ", "minimalist", "vintage",
        "aerial", "sports", "vehicle", "fashion", "cinematic", "texture", "dramatic", "travel"]


def scoring_sections(cfg):
    out = {k: copy.deepcopy(cfg[k]) for k in SECTIONS if k in cfg}
    for cat in out["categories"]:
        cat.pop("tags", None)
    out["viewer"] = {"default_category": cfg.get("viewer", {}).get("default_category", "default")}
    return out


def mutate(cfg):
    c = copy.deepcopy(cfg)
    for k in ("penalties", "exif_adjustments", "thresholds"):
        c.pop(k, None)
    c["exposure"] = {"silhouette_detection": False}
    c["scoring"] = {"score_min": 1.0, "score_max": 9.0}
    c["viewer"] = {"default_category": "misc"}
    cats = [x for x in c["categories"] if x["name"] != "default"]
    for i, cat in enumerate(cats):
        w = cat.get("weights", {})
        if i % 3 == 0 and w:
            k = sorted(w)[0]
            w[k] = w[k] + 7                       # no longer sums to 100 -> renormalised
        if i % 4 == 1:
            cat.pop("modifiers", None)
        if i % 5 == 2:
            cat.setdefault("modifiers", {})["_apply_blink_penalty"] = True
            cat["modifiers"]["_clipping_multiplier"] = 2.0
    dup = copy.deepcopy(cats[-1])                 # same name as an earlier entry, different filter and weights
    dup["priority"] = 1
    dup["filters"] = {"iso_min": 50, "iso_max": 90}
    dup["weights"] = {"aesthetic_percent": 50, "noise_percent": 30, "saturation_percent": 20}
    cats.append(dup)
    c["categories"] = cats
    return c


def rand_metric(rng, i):
    def maybe(v, p_none=0.1):
        return None if rng.random() < p_none else v

    def score():
        r = rng.random()
        if r < 0.04:
            return float(rng.uniform(-150, 150))     # out of [-100, 100] -> default
        if r < 0.07:
            return "%.3f" % rng.uniform(0, 10)        # numeric string
        if r < 0.09:
            return "n/a"
        if r < 0.11:
            return {"__bytes__": "00ff"}
        if r < 0.13:
            return int(rng.integers(0, 11))
        return float(np.round(rng.uniform(0, 10), int(rng.integers(1, 6))))

    m = {}
    for k in ("aesthetic", "face_quality", "eye_sharpness", "tech_sharpness", "color_score", "exposure_score",
              "comp_score", "contrast_score", "face_sharpness", "power_point_score", "quality_score"):
        if rng.random() < 0.85:
            m[k] = maybe(score(), 0.05)
    faces = rng.random() < 0.5
    m["face_count"] = int(rng.integers(1, 6)) if faces else 0
    m["face_ratio"] = float(rng.uniform(0, 0.4)) if faces else 0.0
    if rng.random() < 0.1:
        m["face_ratio"] = None
    m["isolation_bonus"] = float(rng.uniform(1, 4)) if faces else 1.0
    m["is_blink"] = int(rng.random() < 0.3) if faces else 0
    m["is_group_portrait"] = int(faces and rng.random() < 0.3)
    m["shadow_clipped"] = [0, 1, True, False][int(rng.integers(0, 4))]
    m["highlight_clipped"] = [0, 1, True, False][int(rng.integers(0, 4))]
    m["is_silhouette"] = int(rng.random() < 0.15)
    if rng.random() < 0.8:
        m["histogram_spread"] = float(rng.uniform(0, 90))
    if rng.random() < 0.7:
        m["is_monochrome"] = int(rng.random() < 0.2)
    if rng.random() < 0.7:
        m["mean_luminance"] = maybe(float(rng.uniform(0, 1) ** 2))
    if rng.random() < 0.7:
        m["mean_saturation"] = maybe(float(rng.uniform(0, 1)))
    if rng.random() < 0.7:
        m["noise_sigma"] = maybe(float(np.round(rng.uniform(0, 16), 2)))
    if rng.random() < 0.7:
        m["histogram_bimodality"] = maybe(float(rng.uniform(-2, 5)))
    if rng.random() < 0.6:
        m["leading_lines_score"] = maybe(float(np.round(rng.uniform(0, 8), 2)))
    if rng.random() < 0.6:
        m["iso"] = maybe([50, 64, 80, 100, 400, 1600, 6400, 90.5][int(rng.integers(0, 8))])
    if rng.random() < 0.6:
        m["f_stop"] = maybe([1.4, 1.8, 2.0, 2.8, 4.0, 8.0, 0, -1.0, 2][int(rng.integers(0, 9))])
    if rng.random() < 0.5:
        m["shutter_speed"] = maybe(["1/500", "1/30", "2", 15, 30.0, "1/0", "bulb", 1.0, "5"][int(rng.integers(0, 9))])
    if rng.random() < 0.3:
        m["focal_length"] = maybe(float(rng.choice([24, 50, 85, 200])))
    if rng.random() < 0.7:
        k = int(rng.integers(0, 4))
        m["tags"] = ",".join(rng.choice(TAGS, size=k, replace=False).tolist()) if k else maybe("", 0.5)
    if i % 97 == 0:
        m["aesthetic"] = float("nan")
    if i % 101 == 0:
        m["noise_sigma"] = float("nan")
    m["scoring_model"] = "clip-mlp"
    return m


def decode(m):
    return {k: (bytes.fromhex(v["__bytes__"]) if isinstance(v, dict) else v) for k, v in m.items()}


def main():
    from processing.scorer import Facet
    ref_cfg = json.load(open("/root/reference/scoring_config.json"))
    base = scoring_sections(ref_cfg)
    cases = []
    for name, cfg, seed, n in (("reference_config", base, 11, 1500), ("mutated_config", mutate(base), 12, 1500)):
        rng = np.random.default_rng(seed)
        metrics = [rand_metric(rng, i) for i in range(n)]
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "cfg.json")
            json.dump(cfg, open(path, "w"))
            ref = Facet(db_path=os.path.join(td, "t.db"), config_path=path, lightweight=True)
            weights = {c["name"]: ref.config.get_weights(c["name"]) for c in cfg["categories"]}
            out = []
            for m in metrics:
                s, c = ref.calculate_aggregate_logic(decode(m))
                out.append([float(s).hex(), c])
        cases.append({"name": name, "config": cfg, "metrics": metrics, "result": out, "weights": weights})
    # placement data of CompositionAnalyzer (analyzers/composition.py:115) for a few boxes
    from analyzers.composition import CompositionAnalyzer
    rng = np.random.default_rng(5)
    boxes = []
    for _ in range(64):
        w, h = int(rng.integers(100, 6000)), int(rng.integers(100, 4000))
        x1, y1 = int(rng.integers(0, w - 10)), int(rng.integers(0, h - 10))
        box = [x1, y1, int(rng.integers(x1 + 1, w)), int(rng.integers(y1 + 1, h))]
        boxes.append({"bbox": box, "w": w, "h": h, "data": CompositionAnalyzer.get_placement_data(box, w, h, None),
                      "score": CompositionAnalyzer.get_placement_score(box, w, h)})
    boxes.append({"bbox": None, "w": 100, "h": 100, "data": CompositionAnalyzer.get_placement_data(None, 100, 100, None),
                  "score": CompositionAnalyzer.get_placement_score(None, 100, 100)})
    json.dump({"cases": cases, "placement": boxes}, open(os.path.join(HERE, "aggregate_golden.json"), "w"))
    print("wrote", sum(len(c["metrics"]) for c in cases), "metric dicts")


if __name__ == "__main__":
    main()
