"""Shared detection helpers (utils/detection.py:8-29 of the reference)."""


def detect_silhouette(histogram_data, tags, face_count):
    """1 iff (histogram silhouette or a 'silhouette' tag) and a human is present (a face, or a 'portrait' /
    'group' tag).  `tags` is the comma-joined tag string, matched by substring as the reference does."""
    histogram_silhouette = histogram_data.get("is_silhouette", 0)
    clip_silhouette = "silhouette" in tags if tags else False
    has_human = face_count > 0 or (any(t in tags for t in ("portrait", "group")) if tags else False)
    return 1 if ((histogram_silhouette or clip_silhouette) and has_human) else 0
