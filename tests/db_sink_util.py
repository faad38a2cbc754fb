"""Helpers shared by tests/golden/make_golden_db_sink.py and tests/test_db_sink.py."""
import numpy as np

from facet_b200.processing.db_sink import PHOTO_COLUMNS


def synth_result(i, rng):
    emb = rng.standard_normal(768).astype(np.float32)
    res = {
        "path": f"/photos/2024/img_{i:04d}.jpg", "filename": f"img_{i:04d}.jpg",
        "category": ["default", "portrait", "landscape", "night"][i % 4],
        "image_width": int(rng.integers(100, 6000)), "image_height": int(rng.integers(100, 4000)),
        "date_taken": f"2024:05:{1 + i:02d} 10:{i:02d}:00", "camera_model": "X-T5", "lens_model": None,
        "iso": int(rng.integers(64, 6400)), "f_stop": float(rng.choice([1.8, 2.8, 4.0])), "shutter_speed": "1/250",
        "focal_length": 35.0, "focal_length_35mm": 53, "aesthetic": round(float(rng.uniform(0, 10)), 2),
        "face_count": i % 3, "face_quality": round(float(rng.uniform(0, 10)), 2), "eye_sharpness": round(float(rng.uniform(0, 10)), 2),
        "face_sharpness": float(rng.uniform(0, 500)), "face_ratio": float(rng.uniform(0, 0.5)),
        "tech_sharpness": round(float(rng.uniform(0, 10)), 2), "color_score": round(float(rng.uniform(0, 10)), 2),
        "exposure_score": round(float(rng.uniform(0, 10)), 2), "comp_score": round(float(rng.uniform(0, 10)), 2),
        "isolation_bonus": round(float(rng.uniform(1, 3)), 2), "is_blink": int(i % 2), "phash": "%016x" % int(rng.integers(0, 2**63)),
        "aggregate": round(float(rng.uniform(0, 10)), 2), "clip_embedding": emb.tobytes(),
        "raw_sharpness_variance": float(rng.uniform(0, 3000)),
        "histogram_data": rng.random(256).astype(np.float32).tobytes(), "histogram_spread": float(rng.uniform(0, 90)),
        "mean_luminance": float(rng.uniform(0, 1)), "histogram_bimodality": float(rng.uniform(-3, 3)),
        "power_point_score": float(rng.uniform(0, 10)), "raw_color_entropy": float(rng.uniform(0, 15)),
        "raw_eye_sharpness": float(rng.uniform(0, 100)), "config_version": "abc123",
        "shadow_clipped": int(i % 2), "highlight_clipped": 0, "is_silhouette": 0, "is_group_portrait": 0,
        "leading_lines_score": float(rng.uniform(0, 10)), "face_confidence": float(rng.uniform(0, 1)),
        "is_monochrome": 0, "mean_saturation": round(float(rng.uniform(0, 1)), 4),
        "dynamic_range_stops": round(float(rng.uniform(0, 8)), 2), "noise_sigma": round(float(rng.uniform(0, 9)), 2),
        "contrast_score": round(float(rng.uniform(0, 10)), 2), "tags": "sunset,beach" if i % 2 else None,
        "quality_score": None, "topiq_score": None, "composition_explanation": None, "scoring_model": "clip-mlp",
        "composition_pattern": None, "face_details": [],
    }
    for k in range(res["face_count"]):
        res["face_details"].append({
            "index": k, "embedding": rng.standard_normal(512).astype(np.float32).tobytes() if (i + k) % 4 else None,
            "bbox": [int(v) for v in rng.integers(0, 300, size=4)], "confidence": float(rng.uniform(0.5, 1)),
            "thumbnail": bytes(rng.integers(0, 256, size=64, dtype=np.uint8)), "landmark_2d_106": None})
    return res


def _enc(v):
    if isinstance(v, (bytes, bytearray)):
        return {"hex": bytes(v).hex()}
    if isinstance(v, list):
        return [_enc(x) for x in v]
    if isinstance(v, dict):
        return {k: _enc(x) for k, x in v.items()}
    return v


def _dec(v):
    if isinstance(v, dict) and set(v) == {"hex"}:
        return bytes.fromhex(v["hex"])
    if isinstance(v, list):
        return [_dec(x) for x in v]
    if isinstance(v, dict):
        return {k: _dec(x) for k, x in v.items()}
    return v


def encode_result(res):
    return _enc({k: v for k, v in res.items() if k != "thumbnail"})


def decode_result(enc):
    return _dec(enc)


def dump_rows(conn):
    """All photo / face rows in a canonical order, BLOBs as hex; the thumbnail only as its length + sha256."""
    import hashlib
    out = {"photos": [], "faces": []}
    cols = ", ".join(PHOTO_COLUMNS)
    for row in conn.execute(f"SELECT {cols} FROM photos ORDER BY path"):
        rec = {}
        for c, v in zip(PHOTO_COLUMNS, row):
            if c == "thumbnail":
                rec[c] = None if v is None else {"len": len(v), "sha256": hashlib.sha256(v).hexdigest()}
            else:
                rec[c] = {"hex": v.hex()} if isinstance(v, (bytes, bytearray)) else v
        out["photos"].append(rec)
    for row in conn.execute("SELECT photo_path, face_index, embedding, bbox_x1, bbox_y1, bbox_x2, bbox_y2, confidence, "
                            "face_thumbnail, landmark_2d_106 FROM faces ORDER BY photo_path, face_index"):
        out["faces"].append([{"hex": v.hex()} if isinstance(v, (bytes, bytearray)) else v for v in row])
    return out
