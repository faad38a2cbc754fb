"""DB sink (processing/db_sink.py) vs rows written by the UNMODIFIED reference's `Facet.save_photos_batch`
(tests/golden/make_golden_db_sink.py): same inputs, same schema, every column of every row equal — including the JPEG
thumbnail bytes (Pillow path) — over two batches, the second replacing rows of the first."""
import json
import os
import sqlite3

import numpy as np

from conftest import GOLDEN_DIR
from db_sink_util import decode_result, dump_rows


def _golden():
    with open(os.path.join(GOLDEN_DIR, "db_sink_golden.json")) as f:
        return json.load(f)


def _images(seed, n):
    """The PIL images of make_results(): same generator sequence as the golden script."""
    from PIL import Image
    from db_sink_util import synth_result
    rng = np.random.default_rng(seed)
    imgs = []
    for i in range(n):
        synth_result(i, rng)
        h, w = (int(v) for v in rng.integers(40, 900, size=2))
        imgs.append(Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)))
    return imgs


def test_rows_equal_the_reference(tmp_path):
    from facet_b200.processing.db_sink import PhotoSink, save_photos_batch
    g = _golden()
    db = str(tmp_path / "sink.db")
    with sqlite3.connect(db) as conn:
        for sql in g["schema"]:
            conn.execute(sql)
    for batch in g["batches"]:
        results = [decode_result(r) for r in batch["inputs"]]
        n = save_photos_batch(db, list(zip(results, _images(batch["seed"], batch["n"]))))
        assert n == batch["n"]
    with sqlite3.connect(db) as conn:
        rows = dump_rows(conn)
    assert rows["faces"] == g["rows"]["faces"]
    assert len(rows["photos"]) == len(g["rows"]["photos"])
    for got, want in zip(rows["photos"], g["rows"]["photos"]):
        assert got == want, {k: (got[k], want[k]) for k in want if got[k] != want[k]}

    # PhotoSink: flush every 3 rows on one connection, error items skipped, a result without topiq_score -> NULL
    db2 = str(tmp_path / "sink2.db")
    with sqlite3.connect(db2) as conn:
        for sql in g["schema"]:
            conn.execute(sql)
    batch = g["batches"][0]
    results = [decode_result(r) for r in batch["inputs"]]
    results[2].pop("topiq_score")
    with PhotoSink(db2, batch_save_size=3) as sink:
        for r, img in zip(results, _images(batch["seed"], batch["n"])):
            sink.add(r, img)
        sink.add({"path": "/x/bad.jpg", "error": "Failed to load image"})
    assert sink.saved == batch["n"]
    with sqlite3.connect(db2) as conn:
        assert conn.execute("SELECT COUNT(*), COUNT(thumbnail), COUNT(topiq_score) FROM photos").fetchone() == (batch["n"], batch["n"], 0)


def test_atomic_batch_and_thumbnail_forms(tmp_path):
    """A failing row rolls the whole batch back; ready-made JPEG bytes and None are accepted as the image."""
    import pytest
    from facet_b200.processing.db_sink import save_photos_batch
    g = _golden()
    db = str(tmp_path / "sink.db")
    with sqlite3.connect(db) as conn:
        for sql in g["schema"]:
            conn.execute(sql)
    results = [decode_result(r) for r in g["batches"][0]["inputs"]][:3]
    results[1]["aesthetic"] = {"not": "bindable"}
    with pytest.raises(Exception):
        save_photos_batch(db, [(r, None) for r in results])
    with sqlite3.connect(db) as conn:
        assert conn.execute("SELECT COUNT(*) FROM photos").fetchone()[0] == 0
    results[1]["aesthetic"] = 5.0
    assert save_photos_batch(db, [(results[0], b"\xff\xd8jpeg"), (results[1], None), (results[2], None)]) == 3
    with sqlite3.connect(db) as conn:
        got = dict(conn.execute("SELECT path, thumbnail FROM photos").fetchall())
    assert got[results[0]["path"]] == b"\xff\xd8jpeg" and got[results[1]["path"]] is None
