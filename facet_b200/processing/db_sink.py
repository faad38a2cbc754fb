"""Batched DB sink for the result dicts of the pass — the step right after it.

Mirrors `Facet.save_photos_batch` (processing/scorer.py:1670-1749): per result a 640-px LANCZOS JPEG thumbnail
(quality 80), the face rows of `face_details`, and ONE SQLite transaction with `INSERT OR REPLACE` into `photos`
(the 55 columns of scorer.py:1708-1722, same order) and `faces` (scorer.py:1740-1744).  The schema belongs to the
reference (`db/schema.py: init_database`); this module only writes rows into an existing database.

Differences, all deliberate:
  * rows go through `executemany` (one statement compile per batch) instead of one `execute` per photo, and the
    transaction is explicit (`BEGIN IMMEDIATE` ... `COMMIT`), so a batch is atomic;
  * the image of a (result, image) pair may be a PIL image (the reference's contract), a BGR uint8 array or CUDA
    tensor (thumbnail pixels then come from the GPU pass, `utils/image_transforms.generate_photo_thumbnail`: identical
    JPEG bytes), ready-made JPEG `bytes`, or None (thumbnail column NULL);
  * a result that lacks one of the bound columns gets NULL for it.  The reference binds `:topiq_score`, which the
    dicts of `_process_batch` (batch_processor.py:298-355) do not carry — its own call raises
    sqlite3.ProgrammingError there; the multi-pass path adds the key (scorer.py:1132).
"""
from __future__ import annotations

import sqlite3
from io import BytesIO

PHOTO_COLUMNS = (
    "path", "filename", "category", "image_width", "image_height",
    "date_taken", "camera_model", "lens_model", "iso", "f_stop",
    "shutter_speed", "focal_length", "focal_length_35mm", "aesthetic", "face_count", "face_quality",
    "eye_sharpness", "face_sharpness", "face_ratio", "tech_sharpness", "color_score",
    "exposure_score", "comp_score", "isolation_bonus", "is_blink", "phash", "aggregate", "thumbnail",
    "clip_embedding", "raw_sharpness_variance", "histogram_data", "histogram_spread",
    "mean_luminance", "histogram_bimodality", "power_point_score", "raw_color_entropy",
    "raw_eye_sharpness", "config_version",
    "shadow_clipped", "highlight_clipped", "is_silhouette", "is_group_portrait", "leading_lines_score",
    "face_confidence", "is_monochrome", "mean_saturation",
    "dynamic_range_stops", "noise_sigma", "contrast_score", "tags",
    "quality_score", "topiq_score", "composition_explanation", "scoring_model", "composition_pattern",
)
_PHOTO_SQL = (f"INSERT OR REPLACE INTO photos ({', '.join(PHOTO_COLUMNS)}) "
              f"VALUES ({', '.join('?' for _ in PHOTO_COLUMNS)})")
_FACE_SQL = ("INSERT OR REPLACE INTO faces (photo_path, face_index, embedding, bbox_x1, bbox_y1, bbox_x2, bbox_y2, "
             "confidence, face_thumbnail, landmark_2d_106) VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)")


def _thumbnail_bytes(image, size=640, quality=80):
    """scorer.py:1681-1686 for whatever the caller holds."""
    if image is None:
        return None
    if isinstance(image, (bytes, bytearray, memoryview)):
        return bytes(image)
    if hasattr(image, "thumbnail") and hasattr(image, "save"):          # PIL image: the reference's own calls
        from PIL import Image
        thumb = image.copy()
        thumb.thumbnail((size, size), Image.Resampling.LANCZOS)
        buf = BytesIO()
        thumb.save(buf, format="JPEG", quality=quality)
        return buf.getvalue()
    from ..utils.image_transforms import generate_photo_thumbnails     # BGR frame: GPU pixels, same JPEG bytes
    return generate_photo_thumbnails(image[None], size=size, quality=quality, rgb_order=False)[0]


def face_records(result):
    """scorer.py:1689-1704."""
    rows = []
    for face in result.get("face_details", []) or []:
        if face.get("embedding"):
            bbox = face.get("bbox", [0, 0, 0, 0])
            rows.append((result["path"], face["index"], face["embedding"], bbox[0], bbox[1], bbox[2], bbox[3],
                         face.get("confidence", 0), face.get("thumbnail"), face.get("landmark_2d_106")))
    return rows


def apply_pragmas(conn, mmap_size_mb=256, cache_size_mb=64):
    """db/connection.py:34-55 (WAL, busy timeout, NORMAL sync, memory temp store)."""
    conn.execute("PRAGMA journal_mode = WAL")
    conn.execute("PRAGMA busy_timeout = 5000")
    conn.execute("PRAGMA foreign_keys = ON")
    conn.execute("PRAGMA synchronous = NORMAL")
    conn.execute(f"PRAGMA cache_size = -{int(cache_size_mb) * 1000}")
    conn.execute("PRAGMA temp_store = MEMORY")
    conn.execute(f"PRAGMA mmap_size = {int(mmap_size_mb) * 1024 * 1024}")


def save_photos_batch(db_path, results_with_images, conn=None):
    """results_with_images: list of (result_dict, image) pairs (see the module docstring for `image`).  Error items
    ({'path', 'error'}) are skipped like the reference's caller does (batch_processor.py:410-413).  Returns the number
    of photo rows written.  With `conn` the caller's connection is used (and left open)."""
    pairs = [(r, img) for r, img in results_with_images if isinstance(r, dict) and "error" not in r]
    if not pairs:
        return 0
    # Phase 1: thumbnails, no DB lock held
    for res, img in pairs:
        if img is not None or "thumbnail" not in res:       # a result of the streamed pass may already carry its JPEG
            res["thumbnail"] = _thumbnail_bytes(img)
    # Phase 2: rows
    photo_rows = [tuple(res.get(c) for c in PHOTO_COLUMNS) for res, _ in pairs]
    faces = [row for res, _ in pairs for row in face_records(res)]
    # Phase 3: one short transaction
    own = conn is None
    if own:
        conn = sqlite3.connect(db_path, isolation_level=None)
        apply_pragmas(conn)
    try:
        conn.execute("BEGIN IMMEDIATE")
        try:
            conn.executemany(_PHOTO_SQL, photo_rows)
            if faces:
                conn.executemany(_FACE_SQL, faces)
            conn.execute("COMMIT")
        except Exception:
            conn.execute("ROLLBACK")
            raise
    finally:
        if own:
            conn.close()
    return len(photo_rows)


class PhotoSink:
    """Accumulates results and flushes every `batch_save_size` rows (the reference's main-thread loop,
    batch_processor.py:416-455: `pending` list, `save_photos_batch` every 50) on one open connection."""

    def __init__(self, db_path, batch_save_size=50):
        self.db_path = db_path
        self.batch_save_size = int(batch_save_size)
        self.pending = []
        self.saved = 0
        self.conn = sqlite3.connect(db_path, isolation_level=None, check_same_thread=False)
        apply_pragmas(self.conn)

    def add(self, result, image=None):
        if isinstance(result, dict) and "error" not in result:
            self.pending.append((result, image))
            if len(self.pending) >= self.batch_save_size:
                self.flush()

    def flush(self):
        if self.pending:
            self.saved += save_photos_batch(self.db_path, self.pending, conn=self.conn)
            self.pending = []

    def close(self):
        self.flush()
        self.conn.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
