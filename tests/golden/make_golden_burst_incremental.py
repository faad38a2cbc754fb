"""Goldens of `IncrementalBurstProcessor` from the UNMODIFIED reference class (utils/burst.py).

    python tests/golden/make_golden_burst_incremental.py        (build container: /root/reference must exist)

Seeded photo sequences (time gaps from 0.2 s to 20 min, planted near-duplicate hashes, missing dates / hashes, identified
persons on some photos) are fed to the reference class one by one and as a batch; recorded: the open bursts (paths) after the
whole sequence, `get_stats()`, and the `is_burst_lead` column `finalize` leaves in a database with the reference's schema.
"""
import json
import os
import random
import sqlite3
import sys
import tempfile
from datetime import datetime, timedelta

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from utils.burst import IncrementalBurstProcessor  # noqa: E402
from config import ScoringConfig  # noqa: E402
from db import init_database  # noqa: E402


def sequence(seed, n):
    rng = random.Random(seed)
    t = datetime(2024, 5, 1, 10, 0, 0)
    photos = []
    h = rng.getrandbits(64)
    for i in range(n):
        gap = rng.choice([0.0, 0.3, 1.0, 1.9, 2.5, 20, 60, 250, 299, 301, 400, 1200])
        t += timedelta(seconds=gap)
        if rng.random() < 0.45:
            h ^= sum(1 << rng.randrange(64) for _ in range(rng.choice([1, 3, 8, 18, 20])))
        else:
            h = rng.getrandbits(64)
        p = {"path": f"/p/{seed}_{i:04d}.jpg", "date_taken": t.strftime("%Y:%m:%d %H:%M:%S"), "aggregate": round(rng.uniform(2, 9), 2),
             "phash": "%016x" % h}
        r = rng.random()
        if r < 0.05:
            p["date_taken"] = None
        elif r < 0.08:
            p["date_taken"] = "garbage"
        elif r < 0.12:
            p["phash"] = None
        elif r < 0.15:
            p["aggregate"] = None
        if rng.random() < 0.3:
            p["face_details"] = [{"person_id": rng.choice([1, 2, 3, None])} for _ in range(rng.choice([1, 2]))]
        photos.append(p)
    return photos


def run(photos, batch):
    cfg = ScoringConfig("/root/reference/scoring_config.json")
    with tempfile.TemporaryDirectory() as d:
        db = os.path.join(d, "t.db")
        init_database(db)
        conn = sqlite3.connect(db)
        for p in photos:
            conn.execute("INSERT INTO photos (path, filename, date_taken, aggregate, phash, is_burst_lead) VALUES (?, ?, ?, ?, ?, 1)",
                         (p["path"], os.path.basename(p["path"]), p["date_taken"], p["aggregate"], p["phash"]))
        conn.commit()
        proc = IncrementalBurstProcessor(db, cfg)
        trace = []
        if batch:
            proc.add_photos_batch(photos)
        else:
            for p in photos:
                proc.add_photo(p)
                trace.append([len(b) for b in proc.active_bursts])
        open_bursts = [[m["path"] for m in b] for b in proc.active_bursts]
        stats = proc.get_stats()
        marked = proc.finalize(conn)
        leads = [r[0] for r in conn.execute("SELECT path FROM photos WHERE is_burst_lead = 1 ORDER BY path")]
        conn.close()
    return {"open_bursts": open_bursts, "stats": stats, "marked": marked, "leads": leads, "trace": trace}


def main():
    out = {"settings": ScoringConfig("/root/reference/scoring_config.json").get_burst_detection_settings(), "cases": []}
    for seed, n in ((1, 60), (2, 200), (3, 400), (4, 25)):
        photos = sequence(seed, n)
        out["cases"].append({"seed": seed, "n": n, "one_by_one": run(photos, False), "batch": run(photos, True)})
        print(seed, n, out["cases"][-1]["one_by_one"]["stats"], out["cases"][-1]["batch"]["stats"])
    with open(os.path.join(HERE, "burst_incremental_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
