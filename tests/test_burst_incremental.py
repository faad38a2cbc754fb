"""`facet_b200.utils.burst.IncrementalBurstProcessor` against goldens written by the unmodified reference class
(tests/golden/make_golden_burst_incremental.py): open bursts after every photo, statistics, and the `is_burst_lead` column
`finalize` leaves in a database."""
import datetime
import json
import os
import random
import sqlite3

from conftest import GOLDEN_DIR


def _sequence_fn():
    """The seeded photo generator of the golden script (the script itself imports the reference, which the test must not need)."""
    src = open(os.path.join(GOLDEN_DIR, "make_golden_burst_incremental.py")).read()
    ns = {}
    exec(compile(src[src.index("def sequence"):src.index("def run")], "sequence", "exec"),
         {"random": random, "datetime": datetime.datetime, "timedelta": datetime.timedelta}, ns)
    return ns["sequence"]


def test_incremental_bursts_match_reference(tmp_path):
    from facet_b200.utils.burst import IncrementalBurstProcessor
    gold = json.load(open(os.path.join(GOLDEN_DIR, "burst_incremental_golden.json")))
    cfg = type("Cfg", (), {"get_burst_detection_settings": lambda self: dict(gold["settings"])})()
    sequence = _sequence_fn()
    for case in gold["cases"]:
        photos = sequence(case["seed"], case["n"])
        for mode in ("one_by_one", "batch"):
            want = case[mode]
            db = str(tmp_path / f"{case['seed']}_{mode}.db")
            conn = sqlite3.connect(db)
            conn.execute("CREATE TABLE photos (path TEXT PRIMARY KEY, date_taken TEXT, aggregate REAL, phash TEXT, is_burst_lead INTEGER)")
            conn.executemany("INSERT INTO photos VALUES (?, ?, ?, ?, 1)",
                             [(p["path"], p["date_taken"], p["aggregate"], p["phash"]) for p in photos])
            conn.commit()
            proc = IncrementalBurstProcessor(db, cfg)
            trace = []
            if mode == "batch":
                proc.add_photos_batch(photos)
            else:
                for p in photos:
                    proc.add_photo(p)
                    trace.append([len(b) for b in proc.active_bursts])
            assert [[m["path"] for m in b] for b in proc.active_bursts] == want["open_bursts"], (case["seed"], mode)
            assert trace == want["trace"], (case["seed"], mode)
            assert proc.get_stats() == want["stats"]
            assert proc.finalize(conn) == want["marked"]
            assert [r[0] for r in conn.execute("SELECT path FROM photos WHERE is_burst_lead = 1 ORDER BY path")] == want["leads"]
            conn.close()
    # finalize() without a connection opens the database itself
    proc = IncrementalBurstProcessor(db, cfg)
    proc.add_photo({"path": photos[0]["path"], "date_taken": "2024:05:01 10:00:00", "aggregate": 5.0, "phash": "00ff00ff00ff00ff"})
    assert proc.finalize() == 1
