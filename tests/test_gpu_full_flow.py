"""The whole flow a user of the reference runs (`photos.py DIR --single-pass`), on this library:
items -> BatchProcessor.process_items_streamed -> PhotoSink (reference schema, reference columns) -> process_bursts ->
detect_duplicates on the same SQLite file.  Duplicated frames must come out as duplicate groups / bursts, every row
must satisfy the reference validator's invariants (validation/database_validator.py:89-117, 282-376)."""
import json
import os
import sqlite3

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from facet_b200.synth import synth_embeddings, synth_image_bgr

pytestmark = pytest.mark.gpu


def test_single_pass_flow_into_sqlite(tmp_path):
    from facet_b200.config import ScoringConfig
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.processing.batch_processor import BatchProcessor
    from facet_b200.processing.bursts import process_bursts
    from facet_b200.processing.db_sink import PhotoSink
    from facet_b200.processing.scorer import Facet
    from facet_b200.utils.duplicate import detect_duplicates
    with open(os.path.join(GOLDEN_DIR, "aggregate_golden.json")) as f:
        cfg_dict = json.load(f)["cases"][0]["config"]
    cfg_path = tmp_path / "scoring_config.json"
    cfg_path.write_text(json.dumps(cfg_dict))
    cfg = ScoringConfig(str(cfg_path))
    tags = synth_embeddings(24, seed=7, cluster_fraction=0.0)
    sc = Facet(random_state_dict(0), config=cfg, text_embeddings=tags, tag_names=[f"tag{i // 2}" for i in range(24)])
    # 10 distinct frames; frames 3 and 7 appear three times each (shot within a second: a burst and a duplicate group)
    base = [synth_image_bgr(80 + i, 240, 360) for i in range(10)]
    order = list(range(10)) + [3, 3, 7, 7]
    items = []
    for k, i in enumerate(order):
        sec = 10 * i + (k - 9 if k >= 10 else 0)          # the copies follow their original within a few seconds
        items.append({"path": str(tmp_path / f"img_{k:03d}.jpg"), "img_cv": base[i],
                      "exif_data": {"date_taken": "2024:06:01 10:%02d:%02d" % (sec // 60, sec % 60), "camera_model": "synthetic", "lens_model": None,
                                    "iso": 200, "f_stop": 4.0, "shutter_speed": "1/250", "focal_length": 35.0, "focal_length_35mm": 35}})
    db = str(tmp_path / "photos.db")
    with open(os.path.join(GOLDEN_DIR, "db_sink_golden.json")) as f:
        schema = json.load(f)["schema"]
    with sqlite3.connect(db) as conn:
        for sql in schema:
            conn.execute(sql)
    results = BatchProcessor(sc, batch_size=8).process_items_streamed(items, chunk=4, vit_batch=8)
    assert all("error" not in r for r in results)
    with PhotoSink(db, batch_save_size=5) as sink:
        for r, it in zip(results, items):
            sink.add(r, it["img_cv"])                      # BGR frame -> GPU thumbnail pixels + JPEG encode
    assert sink.saved == len(items)
    process_bursts(db, str(cfg_path))
    detect_duplicates(db, str(cfg_path))
    with sqlite3.connect(db) as conn:
        rows = conn.execute("SELECT path, aggregate, aesthetic, mean_luminance, length(histogram_data), length(clip_embedding), "
                            "length(thumbnail), phash, duplicate_group_id, is_duplicate_lead, is_burst_lead, category, tags, "
                            "is_monochrome, mean_saturation FROM photos ORDER BY path").fetchall()
    assert len(rows) == len(items)
    for r in rows:
        assert 0.0 <= r[1] <= 10.0 and 0.0 <= r[2] <= 10.0 and 0.0 <= r[3] <= 1.0      # database_validator.py:89-117, 282-305
        assert r[4] == 1024 and r[5] == 3072 and r[6] > 500                             # :331-376 + a real JPEG thumbnail
        assert len(r[7]) == 16 and r[11]
        assert not (r[13] and r[14] >= 0.1)                                             # :584-607
    by_path = {os.path.basename(r[0]): r for r in rows}
    for members in (["img_003.jpg", "img_010.jpg", "img_011.jpg"], ["img_007.jpg", "img_012.jpg", "img_013.jpg"]):
        gids = {by_path[m][8] for m in members}
        assert len(gids) == 1 and None not in gids, "identical frames must share a duplicate group"
        assert sum(by_path[m][9] for m in members) == 1, "exactly one duplicate lead per group"
        assert sum(by_path[m][10] for m in members) == 1, "identical frames seconds apart are one burst with one lead"


def test_process_files_from_paths_to_rows(tmp_path):
    """`BatchProcessor.process_files(paths, db_path)`: JPEG files (with and without restart markers, one progressive, one
    rotated by EXIF), a PNG, an unreadable file -> rows in a database with the reference's schema."""
    from PIL import Image
    from facet_b200.models.clip_vit import random_state_dict
    from facet_b200.processing.batch_processor import BatchProcessor
    from facet_b200.processing.scorer import Facet
    sc = Facet(random_state_dict(0))
    paths = []
    for i, kw in enumerate([{"quality": 90}, {"quality": 85, "restart_marker_blocks": 4}, {"quality": 80, "progressive": True}, {"quality": 92}]):
        rgb = synth_image_bgr(100 + i, 300, 420)[:, :, ::-1].copy()
        ex = Image.Exif()
        ex[0x0112] = 6 if i == 3 else 1
        p = tmp_path / f"photo_{i}.jpg"
        Image.fromarray(rgb).save(p, "JPEG", exif=ex, **kw)
        paths.append(str(p))
    p = tmp_path / "drawing.png"
    Image.fromarray(synth_image_bgr(110, 200, 200)[:, :, ::-1].copy()).save(p)
    paths.append(str(p))
    bad = tmp_path / "broken.jpg"
    bad.write_bytes(b"not an image")
    paths.append(str(bad))
    db = str(tmp_path / "photos.db")
    with open(os.path.join(GOLDEN_DIR, "db_sink_golden.json")) as f:
        schema = json.load(f)["schema"]
    with sqlite3.connect(db) as conn:
        for sql in schema:
            conn.execute(sql)
    bp = BatchProcessor(sc, batch_size=8, num_workers=3)
    assert bp.process_files(paths, db_path=db, show_metrics=False, batch_save_size=2, chunk=2, vit_batch=4) == 5
    assert bp.metrics.get("host_decoded") == 1              # the progressive file
    with sqlite3.connect(db) as conn:
        rows = dict((os.path.basename(r[0]), r[1:]) for r in conn.execute(
            "SELECT path, image_width, image_height, length(thumbnail), length(clip_embedding), phash FROM photos"))
    assert set(rows) == {"photo_0.jpg", "photo_1.jpg", "photo_2.jpg", "photo_3.jpg", "drawing.png"}
    assert rows["photo_0.jpg"][:2] == (420, 300) and rows["photo_3.jpg"][:2] == (300, 420)      # EXIF 6: rotated upright
    assert all(r[2] > 500 and r[3] == 3072 and len(r[4]) == 16 for r in rows.values())
    # without a database the dicts come back, in input order, the unreadable file as an error item
    res = bp.process_files(paths, show_metrics=False)
    assert [os.path.basename(r["path"]) for r in res] == [os.path.basename(p) for p in paths] and "error" in res[-1]
    # process_stream (batch_processor.py:458): None when done; the remaining paths when the calibration callback wants new workers
    seen = []
    assert bp.process_stream(iter(paths), len(paths), calibration_callback=lambda m: seen.append(m) or False, calibration_size=1,
                             show_metrics=False) is None
    assert len(seen) == 1 and [os.path.basename(r["path"]) for r in bp.last_results] == [os.path.basename(p) for p in paths]
    assert bp.process_stream(iter(paths), len(paths), calibration_callback=lambda m: True, calibration_size=1, show_metrics=False) == paths[2:]
    assert bp.process_stream(iter([]), 0) is None
