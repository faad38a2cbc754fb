"""Golden vectors for duplicate / burst grouping, produced by the UNMODIFIED reference functions
(utils/duplicate.detect_duplicates, processing/scorer.process_bursts) on temporary SQLite DBs.

    python tests/golden/make_golden_grouping.py   ->  tests/golden/grouping_golden.json
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
sys.path.insert(0, "/root/reference")

from facet_b200.synth import synth_hashes, synth_timestamps  # noqa: E402

CONFIG = "/root/reference/scoring_config.json"


def make_rows(n, seed):
    rng = np.random.default_rng(seed)
    hashes = synth_hashes(n, seed=seed, dup_fraction=0.3, max_flip=24)
    ts = synth_timestamps(n, seed=seed + 1)
    agg = np.round(rng.uniform(0, 10, size=n), 2)
    agg[rng.random(n) < 0.15] = 5.0            # ties exercise "first max wins"
    perm = rng.permutation(n)                    # path order != time order
    rows = []
    for k in range(n):
        sec = int(ts[k])
        date = "2024:03:%02d %02d:%02d:%02d" % (1 + sec // 86400, (sec // 3600) % 24, (sec // 60) % 60, sec % 60)
        if seed % 2 == 1 and k % 37 == 5:
            date = None if k % 2 else "not a date"
        rows.append({"path": "/p/img_%06d.jpg" % int(perm[k]), "phash": "%016x" % int(hashes[k]),
                     "aggregate": None if (k % 53 == 7) else float(agg[k]), "date_taken": date})
    return rows


def run_reference(rows, with_faces):
    from db import init_database
    from utils.duplicate import detect_duplicates
    from processing.scorer import process_bursts
    with tempfile.TemporaryDirectory() as td:
        db = os.path.join(td, "t.db")
        init_database(db)
        with sqlite3.connect(db) as conn:
            for r in rows:
                conn.execute("INSERT INTO photos (path, filename, phash, aggregate, date_taken) VALUES (?,?,?,?,?)",
                             (r["path"], os.path.basename(r["path"]), r["phash"], r["aggregate"], r["date_taken"]))
            persons = {}
            if with_faces:
                rng = np.random.default_rng(99)
                for r in rows:
                    if rng.random() < 0.5:
                        pid = int(rng.integers(1, 4))
                        conn.execute("INSERT OR IGNORE INTO persons (id, name) VALUES (?, ?)", (pid, "p%d" % pid))
                        conn.execute("INSERT INTO faces (photo_path, face_index, embedding, person_id) VALUES (?,?,?,?)",
                                     (r["path"], 0, b"\x00" * 2048, pid))
                        persons.setdefault(r["path"], []).append(pid)
            conn.commit()
        detect_duplicates(db, CONFIG)
        process_bursts(db, CONFIG)
        with sqlite3.connect(db) as conn:
            out = {p: (g, l, b) for p, g, l, b in conn.execute(
                "SELECT path, duplicate_group_id, is_duplicate_lead, is_burst_lead FROM photos")}
    return out, persons


def main():
    cases = []
    for n, seed, faces in [(60, 3, False), (400, 4, False), (401, 5, True), (1500, 6, False), (2, 7, False)]:
        rows = make_rows(n, seed)
        res, persons = run_reference(rows, faces)
        cases.append({"n": n, "seed": seed, "rows": rows, "persons": persons,
                      "result": {p: list(v) for p, v in res.items()}})
        groups = len({v[0] for v in res.values() if v[0]})
        print(n, seed, "dup groups", groups, "burst leads", sum(1 for v in res.values() if v[2]))
    with open(os.path.join(HERE, "grouping_golden.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden_grouping.py",
                   "reference": "utils/duplicate.py detect_duplicates + processing/scorer.py process_bursts (unmodified)",
                   "config": "scoring_config.json: duplicate 90 %, burst 70 % / 0.8 min / 0.4 s",
                   "cases": cases}, f)


if __name__ == "__main__":
    main()
