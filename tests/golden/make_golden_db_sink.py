"""Golden rows for the DB sink from the UNMODIFIED reference (`Facet.save_photos_batch`, processing/scorer.py:1670-1749).

Run in the build container:   python tests/golden/make_golden_db_sink.py
Writes tests/golden/db_sink_golden.json: the reference's CREATE statements for `photos` / `faces` (db/schema.py via
init_database), the seeded input results, and every row the reference wrote (BLOBs as hex).
"""
import json
import os
import sqlite3
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")


def make_results(n=7, seed=3):
    """Complete result dicts (every bound column, as the multi-pass path produces them) + PIL images."""
    from PIL import Image
    from db_sink_util import synth_result
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        res = synth_result(i, rng)
        h, w = (int(v) for v in rng.integers(40, 900, size=2))
        img = Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8))
        out.append((res, img))
    return out


def main():
    import types
    # `_load_image_modules` (scorer.py:71) imports the third-party imagehash, absent here and unused by the sink
    sys.modules.setdefault("imagehash", types.ModuleType("imagehash"))
    import processing.scorer as ref_scorer
    from processing.scorer import Facet
    ref_scorer._load_image_modules()
    from db_sink_util import dump_rows, encode_result
    with tempfile.TemporaryDirectory() as td:
        db = os.path.join(td, "g.db")
        facet = Facet(db_path=db, config_path="/root/reference/scoring_config.json", lightweight=True)
        pairs = make_results()
        inputs = [encode_result(r) for r, _ in pairs]
        facet.save_photos_batch(pairs)
        # a second batch that replaces two rows (INSERT OR REPLACE) and adds one
        pairs2 = make_results(3, seed=4)
        pairs2[0][0]["path"] = pairs[1][0]["path"]
        pairs2[1][0]["path"] = pairs[5][0]["path"]
        inputs2 = [encode_result(r) for r, _ in pairs2]
        facet.save_photos_batch(pairs2)
        with sqlite3.connect(db) as conn:
            schema = [r[0] for r in conn.execute("SELECT sql FROM sqlite_master WHERE name IN ('photos', 'faces') AND type = 'table'")]
            rows = dump_rows(conn)
    golden = {"generator": "tests/golden/make_golden_db_sink.py", "reference": "Facet.save_photos_batch (unmodified)",
              "schema": schema, "batches": [{"seed": 3, "n": 7, "inputs": inputs}, {"seed": 4, "n": 3, "inputs": inputs2}], "rows": rows}
    with open(os.path.join(HERE, "db_sink_golden.json"), "w") as f:
        json.dump(golden, f)
    print("photos", len(rows["photos"]), "faces", len(rows["faces"]))


if __name__ == "__main__":
    main()
