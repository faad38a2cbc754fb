"""`IncrementalBurstProcessor` with the reference's interface (utils/burst.py:8-233).

The reference never instantiates this class (it is only re-exported, utils/__init__.py:17; the live burst rule is
`process_bursts`, processing/scorer.py:1880, mirrored in facet_b200/processing/bursts.py with its Hamming test on the GPU).  It is
kept for callers that group photos as they are saved: pure host bookkeeping over a handful of open bursts, so there is nothing to
put on the device.  Same behaviour as the reference class (checked against goldens written by it, tests/test_burst_incremental.py):
  * a photo joins the FIRST open burst that has a member with |dt| <= rapid_burst_seconds and no conflicting identified persons,
    or with |dt| <= time_window_minutes * 60 and pHash distance <= int(64 * (1 - similarity / 100));
  * photos without a pHash are ignored, photos without a parsable date never match and open their own burst;
  * after every photo, bursts whose newest member is more than the window + 60 s older than that photo are dropped from the
    open list (the reference drops them without remembering them, so `finalize` only marks leaders of the bursts still open);
  * `finalize` clears `is_burst_lead` and marks the highest aggregate (first on ties) of every open burst.
"""
from __future__ import annotations

from datetime import datetime


class IncrementalBurstProcessor:
    def __init__(self, db_path, config):
        self.db_path = db_path
        self.datetime = datetime
        settings = config.get_burst_detection_settings()
        self.time_window_minutes = settings.get("time_window_minutes", 5)
        self.rapid_burst_seconds = settings.get("rapid_burst_seconds", 2)
        self.max_hamming_distance = int(64 * (1 - settings.get("similarity_threshold_percent", 70) / 100))
        self.active_bursts = []
        self.photo_persons = {}

    # -- helpers with the reference's names ----------------------------------------------------------------------------
    def _parse_date(self, date_str):
        if not date_str:
            return None
        try:
            return self.datetime.strptime(date_str[:19], "%Y:%m:%d %H:%M:%S")
        except (ValueError, TypeError):
            return None

    def _phash_distance(self, hash1, hash2):
        if not hash1 or not hash2:
            return 999
        try:
            return (int(hash1, 16) ^ int(hash2, 16)).bit_count()
        except (ValueError, TypeError):
            return 999

    def _shares_person(self, path1, path2):
        a = self.photo_persons.get(path1, set())
        b = self.photo_persons.get(path2, set())
        return True if (not a or not b) else bool(a & b)

    def _is_similar(self, photo, burst_photo):
        t0 = self._parse_date(photo.get("date_taken"))
        t1 = self._parse_date(burst_photo.get("date_taken"))
        if t0 is None or t1 is None:
            return False
        dt = abs((t0 - t1).total_seconds())
        if dt <= self.rapid_burst_seconds and self._shares_person(photo.get("path", ""), burst_photo.get("path", "")):
            return True
        return (dt <= self.time_window_minutes * 60 and
                self._phash_distance(photo.get("phash"), burst_photo.get("phash")) <= self.max_hamming_distance)

    def _find_matching_burst(self, photo):
        for burst in self.active_bursts:
            if any(self._is_similar(photo, member) for member in burst):
                return burst
        return None

    # -- interface -------------------------------------------------------------------------------------------------------
    def add_photo(self, photo_data):
        if not photo_data.get("phash"):
            return
        persons = {f["person_id"] for f in (photo_data.get("face_details") or []) if f.get("person_id")}
        if persons:
            self.photo_persons[photo_data["path"]] = persons
        burst = self._find_matching_burst(photo_data)
        if burst is None:
            self.active_bursts.append([photo_data])
        else:
            burst.append(photo_data)
        self._prune_old_bursts(photo_data.get("date_taken"))

    def _prune_old_bursts(self, current_date_str):
        now = self._parse_date(current_date_str) if current_date_str else None
        if now is None:
            return
        max_age = self.time_window_minutes * 60 + 60
        kept = []
        for burst in self.active_bursts:
            dates = [d for d in (self._parse_date(p.get("date_taken")) for p in burst) if d is not None]
            if dates and (now - max(dates)).total_seconds() <= max_age:
                kept.append(burst)
        self.active_bursts = kept

    def add_photos_batch(self, photos_data):
        for photo in sorted(photos_data, key=lambda p: p.get("date_taken") or ""):
            self.add_photo(photo)

    def finalize(self, conn=None):
        if conn is not None:
            return self._finalize_with_conn(conn)
        import sqlite3
        own = sqlite3.connect(self.db_path)
        try:
            return self._finalize_with_conn(own)
        finally:
            own.close()

    def _finalize_with_conn(self, conn):
        conn.execute("UPDATE photos SET is_burst_lead = 0")
        marked = 0
        for burst in self.active_bursts:
            if burst:
                winner = max(burst, key=lambda x: x.get("aggregate") or 0)
                conn.execute("UPDATE photos SET is_burst_lead = 1 WHERE path = ?", (winner["path"],))
                marked += 1
        conn.commit()
        return marked

    def get_stats(self):
        total = sum(len(b) for b in self.active_bursts)
        return {"active_bursts": len(self.active_bursts), "total_photos": total,
                "avg_burst_size": total / len(self.active_bursts) if self.active_bursts else 0}
