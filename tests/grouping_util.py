"""Shared helpers for grouping tests: a minimal photos/faces schema (only the columns the two
reference functions touch, db/schema.py:14-90) and loaders for the golden cases."""
import json
import os
import sqlite3

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grouping_golden.json")

TEST_CONFIG = {
    "categories": [],
    "burst_detection": {"similarity_threshold_percent": 70, "time_window_minutes": 0.8, "rapid_burst_seconds": 0.4},
    "duplicate_detection": {"similarity_threshold_percent": 90},
}


def load_cases():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def make_db(path, rows, persons):
    conn = sqlite3.connect(path)
    conn.execute("CREATE TABLE photos (path TEXT PRIMARY KEY, filename TEXT, phash TEXT, aggregate REAL, "
                 "date_taken TEXT, duplicate_group_id INTEGER, is_duplicate_lead INTEGER DEFAULT 0, "
                 "is_burst_lead INTEGER DEFAULT 0)")
    # db/schema.py:144 — the reference schema carries this index; SQLite walks it backwards for
    # ORDER BY date_taken, which fixes the order of rows with equal dates (descending rowid).
    conn.execute("CREATE INDEX idx_date_taken_desc ON photos(date_taken DESC)")
    conn.execute("CREATE TABLE faces (id INTEGER PRIMARY KEY, photo_path TEXT, face_index INTEGER, person_id INTEGER)")
    for r in rows:
        conn.execute("INSERT INTO photos (path, filename, phash, aggregate, date_taken) VALUES (?,?,?,?,?)",
                     (r["path"], os.path.basename(r["path"]), r["phash"], r["aggregate"], r["date_taken"]))
    for p, pids in (persons or {}).items():
        for pid in pids:
            conn.execute("INSERT INTO faces (photo_path, face_index, person_id) VALUES (?,?,?)", (p, 0, pid))
    conn.commit()
    conn.close()


def write_config(path):
    with open(path, "w") as f:
        json.dump(TEST_CONFIG, f)


def read_result(path):
    with sqlite3.connect(path) as conn:
        return {p: [g, l, b] for p, g, l, b in conn.execute(
            "SELECT path, duplicate_group_id, is_duplicate_lead, is_burst_lead FROM photos")}


def sqlite_order(rows, key):
    """Row order of `SELECT ... ORDER BY key` on the reference schema: NULLs first; `path` is unique;
    equal `date_taken` values come back in descending rowid order (backward scan of
    idx_date_taken_desc, db/schema.py:144)."""
    idx = list(range(len(rows)))
    if key == "date_taken":
        idx.reverse()
    idx.sort(key=lambda i: (rows[i][key] is not None, rows[i][key] or ""))
    return idx
