"""Aggregate score / category / config validation against outputs of the unmodified reference
(tests/golden/aggregate_golden.json, made by tests/golden/make_golden_aggregate.py): bit-exact float64."""
import json
import os

import numpy as np
import pytest

from facet_b200.config import ScoringConfig
from facet_b200.processing.aggregate import AggregateScorer, calculate_aggregate_logic

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "aggregate_golden.json")


def _decode(m):
    return {k: (bytes.fromhex(v["__bytes__"]) if isinstance(v, dict) else v) for k, v in m.items()}


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def _config(case, tmp_path_factory):
    path = tmp_path_factory.mktemp("cfg") / (case["name"] + ".json")
    path.write_text(json.dumps(case["config"]))
    return ScoringConfig(str(path))          # validates (and rewrites) like the reference's constructor


@pytest.mark.parametrize("idx", [0, 1])
def test_weights_after_validation_match_reference(golden, tmp_path_factory, idx):
    case = golden["cases"][idx]
    cfg = _config(case, tmp_path_factory)
    for name, want in case["weights"].items():
        assert cfg.get_weights(name) == want, name
    assert cfg.get_weights("no such category") == {}


@pytest.mark.parametrize("idx", [0, 1])
def test_batch_aggregate_bit_exact(golden, tmp_path_factory, idx):
    case = golden["cases"][idx]
    cfg = _config(case, tmp_path_factory)
    rows = [_decode(m) for m in case["metrics"]]
    scores, cats = AggregateScorer(cfg).score_batch(rows)
    assert len(set(cats)) >= 10                  # the vectors exercise many categories
    for i, (want_hex, want_cat) in enumerate(case["result"]):
        assert cats[i] == want_cat, (i, rows[i])
        assert float(scores[i]).hex() == want_hex, (i, cats[i], rows[i])


def test_scalar_entry_equals_batch(golden, tmp_path_factory):
    case = golden["cases"][0]
    cfg = _config(case, tmp_path_factory)
    for m, (want_hex, want_cat) in list(zip(case["metrics"], case["result"]))[:200]:
        s, c = calculate_aggregate_logic(_decode(m), cfg)
        assert isinstance(s, float) and s.hex() == want_hex and c == want_cat
    assert AggregateScorer(cfg).score_batch([])[1] == []


def test_placement_data_matches_reference(golden):
    from facet_b200.analyzers.composition import CompositionAnalyzer
    for p in golden["placement"]:
        assert CompositionAnalyzer.get_placement_data(p["bbox"], p["w"], p["h"], None) == p["data"]
        assert CompositionAnalyzer.get_placement_score(p["bbox"], p["w"], p["h"]) == p["score"]


def test_detect_silhouette_truth_table():
    from facet_b200.utils.detection import detect_silhouette
    assert detect_silhouette({"is_silhouette": True}, None, 0) == 0          # no human
    assert detect_silhouette({"is_silhouette": True}, None, 2) == 1
    assert detect_silhouette({"is_silhouette": False}, "sunset,silhouette", 1) == 1
    assert detect_silhouette({}, "silhouette,group", 0) == 1                 # tag-only human
    assert detect_silhouette({"is_silhouette": 1}, "landscape", 0) == 0


def test_config_rejects_old_format(tmp_path):
    p = tmp_path / "c.json"
    p.write_text(json.dumps({"weights": {}}))
    with pytest.raises(ValueError):
        ScoringConfig(str(p))
    with pytest.raises(FileNotFoundError):
        ScoringConfig(str(tmp_path / "missing.json"))


# result columns of the reference's single-pass consumer (processing/batch_processor.py:298-355)
REFERENCE_COLUMNS = {
    "path", "filename", "category", "image_width", "image_height", "aesthetic", "face_count", "face_quality",
    "eye_sharpness", "face_sharpness", "face_ratio", "tech_sharpness", "color_score", "exposure_score", "comp_score",
    "isolation_bonus", "is_blink", "phash", "aggregate", "clip_embedding", "raw_sharpness_variance", "histogram_data",
    "histogram_spread", "mean_luminance", "histogram_bimodality", "power_point_score", "raw_color_entropy",
    "raw_eye_sharpness", "config_version", "shadow_clipped", "highlight_clipped", "is_silhouette", "is_group_portrait",
    "leading_lines_score", "face_confidence", "is_monochrome", "mean_saturation", "dynamic_range_stops", "noise_sigma",
    "contrast_score", "tags", "quality_score", "composition_explanation", "scoring_model", "composition_pattern",
    "face_details"}


class _FakeScorer:
    """Stands in for the device pass: returns analyzer-shaped dicts so the host-side assembly runs without a GPU."""

    def __init__(self, config, rows):
        self.config, self.rows = config, rows

    def score_images(self, frames, **_):
        out = []
        for k in range(len(frames)):
            r = dict(self.rows[int(frames[k, 0, 0, 0])])
            r.update(image_height=frames.shape[1], image_width=frames.shape[2])
            out.append(r)
        return out


def test_batch_processor_assembles_reference_columns(golden, tmp_path_factory):
    from facet_b200.processing.batch_processor import BatchProcessor
    cfg = _config(golden["cases"][0], tmp_path_factory)
    rng = np.random.default_rng(3)
    rows, items = [], []
    for i in range(24):
        aest, sharp, col, expo = (float(x) for x in rng.uniform(0, 10, 4))
        rows.append({
            "aesthetic_unrounded": aest, "tech_sharpness_unrounded": sharp, "color_score_unrounded": col,
            "exposure_score_unrounded": expo, "aesthetic": round(aest, 2), "tech_sharpness": round(sharp, 2),
            "color_score": round(col, 2), "exposure_score": round(expo, 2), "clip_embedding": b"\0" * 3072,
            "raw_sharpness_variance": float(rng.uniform(0, 900)), "histogram_data": b"\0" * 1024,
            "histogram_spread": float(rng.uniform(0, 90)), "mean_luminance": float(rng.uniform(0, 1)),
            "histogram_bimodality": float(rng.uniform(-2, 4)), "raw_color_entropy": float(rng.uniform(0, 15)),
            "shadow_clipped": int(rng.random() < 0.3), "highlight_clipped": int(rng.random() < 0.3),
            "is_silhouette": int(rng.random() < 0.4), "is_monochrome": int(rng.random() < 0.3),
            "mean_saturation": 0.3, "dynamic_range_stops": 5.1, "noise_sigma": 2.2, "contrast_score": 6.0,
            "tags": [None, "portrait,street", "landscape", "silhouette,group"][i % 4], "quality_score": None,
            "scoring_model": "clip-mlp", "phash": "%016x" % i})
        img = np.zeros((40, 60, 3), np.uint8)
        img[0, 0, 0] = i
        item = {"path": f"/photos/a/img{i}.jpg", "img_cv": img}
        if i % 3 == 0:      # what a face analyzer outside this path would attach
            item["face_res"] = {"face_count": 1 + i % 2, "face_quality": 7.5, "eye_sharpness": 6.0, "is_blink": i % 2,
                                "face_area": 300 + 40 * i, "bbox": [10, 5, 30, 25], "face_sharpness": 400.0 + i,
                                "raw_eye_sharpness": 88.0, "is_group_portrait": int(i % 6 == 0), "max_face_confidence": 0.9,
                                "face_details": [{"bbox": [10, 5, 30, 25]}]}
        if i % 4 == 1:
            item["exif_data"] = {"iso": [64, 1600][i % 2], "f_stop": [1.8, 2.8, 8.0][i % 3], "camera_model": "X"}
        items.append(item)
    res = list(BatchProcessor(_FakeScorer(cfg, rows), batch_size=7).process_items(items))
    seen = set()
    for i, (item, r) in enumerate(zip(items, res)):
        assert "error" not in r, r
        assert REFERENCE_COLUMNS <= set(r) and not any(k.endswith("_unrounded") for k in r)
        face = item.get("face_res")
        exif = item.get("exif_data", {})
        face_count = face["face_count"] if face else 0
        sil = 1 if ((rows[i]["is_silhouette"] or "silhouette" in (rows[i]["tags"] or ""))
                    and (face_count > 0 or any(t in (rows[i]["tags"] or "") for t in ("portrait", "group")))) else 0
        assert r["is_silhouette"] == sil
        iso_bonus = max(1.0, face["face_sharpness"] / (rows[i]["raw_sharpness_variance"] + 1)) if face else 1.0
        m = {"aesthetic": rows[i]["aesthetic_unrounded"], "face_count": face_count,
             "face_quality": face["face_quality"] if face else 0, "eye_sharpness": face["eye_sharpness"] if face else 0,
             "tech_sharpness": rows[i]["tech_sharpness_unrounded"], "color_score": rows[i]["color_score_unrounded"],
             "exposure_score": rows[i]["exposure_score_unrounded"],
             "face_ratio": (face["face_area"] / 2400) if face else 0.0, "comp_score": r["comp_score"],
             "isolation_bonus": iso_bonus, "is_blink": face["is_blink"] if face else 0,
             "shadow_clipped": rows[i]["shadow_clipped"], "highlight_clipped": rows[i]["highlight_clipped"],
             "is_silhouette": sil, "histogram_spread": rows[i]["histogram_spread"], "iso": exif.get("iso"),
             "f_stop": exif.get("f_stop"), "quality_score": None, "scoring_model": "clip-mlp"}
        want, cat = calculate_aggregate_logic(m, cfg)
        assert (r["aggregate"], r["category"]) == (round(want, 2), cat)
        assert r["comp_score"] == (7.0 if not face else r["comp_score"]) and r["power_point_score"] == (5.0 if not face else r["power_point_score"])
        assert r["config_version"] == cfg.version_hash and r["filename"] == f"img{i}.jpg"
        if exif:
            assert r["camera_model"] == "X"
        seen.add(cat)
    assert {"default", "portrait"} <= seen or len(seen) >= 3


def test_refresh_picks_up_in_place_edits(golden):
    cfg = ScoringConfig.from_dict(json.loads(json.dumps(golden["cases"][0]["config"])))
    m = {"aesthetic": 8.0, "tech_sharpness": 3.0, "exposure_score": 6.0, "comp_score": 5.0, "color_score": 4.0}
    before, cat = calculate_aggregate_logic(m, cfg)
    assert cat == "default"
    default = next(c for c in cfg.config["categories"] if c["name"] == "default")
    default["weights"] = {"aesthetic_percent": 100}
    default["modifiers"] = {}
    assert cfg.get_weights("default") == {"aesthetic": 1.0}           # the getter reads the live config
    old_hash = cfg.version_hash
    cfg.refresh()
    after, _ = calculate_aggregate_logic(m, cfg)
    assert cfg.version_hash != old_hash and after == 8.0 and after != before
