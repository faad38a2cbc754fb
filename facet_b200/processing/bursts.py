"""Burst grouping — drop-in for ``process_bursts`` (processing/scorer.py:1880-1986).

The reference walks the photos in ``ORDER BY date_taken`` order and keeps a *contiguous* current
burst; photo i joins it when ANY member b satisfies

    (|dt| <= rapid_s  and shares_person(i,b) and hamming <= 2*thr)     scorer.py:1957-1960
 or (|dt| <= window_min*60            and hamming <= thr)             scorer.py:1962-1965

Because the burst is the contiguous run [start, i-1], "any member matches" is equivalent to
"the largest matching b (over all earlier photos in the time window) is >= start".  The GPU
kernel (csrc/hamming.cu burst_links_kernel) computes that largest b for every i in parallel;
the chain itself is then a trivial sequential scan on the host.  The person check needs
string sets, so rapid-rule candidates come back as a sparse pair list and are filtered here.
"""
from __future__ import annotations

import sqlite3
from datetime import datetime

import numpy as np

from .. import ops
from ..utils.duplicate import max_hamming_distance


def _parse_date(date_str):
    """scorer.py:1931-1937."""
    if not date_str:
        return None
    try:
        return datetime.strptime(date_str[:19], "%Y:%m:%d %H:%M:%S")
    except (ValueError, TypeError):
        return None


_EPOCH = datetime(1970, 1, 1)


def burst_leads(dates, hashes_hex, aggregates, paths=None, photo_persons=None, *, similarity_percent=88,
                time_window_minutes=60, rapid_burst_seconds=5):
    """Rows in ORDER BY date_taken order -> is_burst_lead uint8[n] (scorer.py:1970-1984)."""
    n = len(dates)
    lead = np.zeros(n, np.uint8)
    if n == 0:
        return lead
    thr = max_hamming_distance(similarity_percent)
    parsed = [_parse_date(d) for d in dates]
    flags = np.zeros(n, np.uint8)
    t = np.zeros(n, np.int64)
    h = np.zeros(n, np.uint64)
    for i in range(n):
        if parsed[i] is not None:
            flags[i] |= 1
            t[i] = int((parsed[i] - _EPOCH).total_seconds())
        if hashes_hex[i]:
            flags[i] |= 2
            h[i] = int(hashes_hex[i], 16)
    window_s = time_window_minutes * 60
    # lower bound of the window that is valid for ANY row order: prefix maxima are monotone
    valid_t = np.where(flags & 1, t, np.iinfo(np.int64).min)
    pmax = np.maximum.accumulate(valid_t)
    span = max(float(window_s), float(rapid_burst_seconds))
    lo = np.searchsorted(pmax, t - int(np.ceil(span)), side="left").astype(np.int32)
    lo = np.minimum(lo, np.arange(n, dtype=np.int32))
    last_slow, rapid = ops.burst_links(h, t, flags, lo, thr, int(np.floor(window_s)), float(rapid_burst_seconds))
    last = last_slow.astype(np.int64)
    persons = photo_persons or {}
    for i, b in rapid.tolist():
        ok = True
        if persons and paths is not None:
            p1, p2 = persons.get(paths[i], set()), persons.get(paths[b], set())
            ok = (not p1) or (not p2) or bool(p1 & p2)
        if ok and b > last[i]:
            last[i] = b

    def finalize(start, end):
        best = start
        best_v = aggregates[start] or 0
        for k in range(start + 1, end):
            v = aggregates[k] or 0
            if v > best_v:
                best, best_v = k, v
        lead[best] = 1

    start = 0
    for i in range(1, n):
        if last[i] >= start:
            continue
        finalize(start, i)
        start = i
    finalize(start, n)
    return lead


def process_bursts(db_path, config_path="scoring_config.json"):
    """Same contract as the reference: flags the highest-scoring photo of each burst in SQLite."""
    from ..config import ScoringConfig
    cfg = ScoringConfig(config_path).get_burst_detection_settings()
    pct = cfg.get("similarity_threshold_percent", 88)
    window_min = cfg.get("time_window_minutes", 60)
    rapid_s = cfg.get("rapid_burst_seconds", 5)
    print(f"Processing burst groups (rapid<={rapid_s}s, similarity>={pct}%, window={window_min}min)...")
    with sqlite3.connect(db_path) as conn:
        rows = conn.execute("SELECT path, date_taken, aggregate, phash FROM photos "
                            "WHERE phash IS NOT NULL ORDER BY date_taken").fetchall()
        if not rows:
            return
        persons = {}
        try:
            if conn.execute("SELECT 1 FROM faces LIMIT 1").fetchone():
                for path, pid in conn.execute("SELECT photo_path, person_id FROM faces WHERE person_id IS NOT NULL"):
                    persons.setdefault(path, set()).add(pid)
        except sqlite3.OperationalError:
            persons = {}
        paths = [r[0] for r in rows]
        lead = burst_leads([r[1] for r in rows], [r[3] for r in rows], [r[2] for r in rows], paths, persons,
                           similarity_percent=pct, time_window_minutes=window_min, rapid_burst_seconds=rapid_s)
        conn.execute("UPDATE photos SET is_burst_lead = 0 WHERE phash IS NOT NULL")
        conn.execute("UPDATE photos SET is_burst_lead = 1 WHERE phash IS NULL")
        conn.executemany("UPDATE photos SET is_burst_lead = 1 WHERE path = ?",
                         [(paths[k],) for k in np.flatnonzero(lead).tolist()])
        conn.commit()
