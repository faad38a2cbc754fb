"""Aggregate score + category for a batch of scored photos (the consumer of the per-image pass).

Restates `Facet.calculate_aggregate_logic` (processing/scorer.py:769-950) with its helpers
`_safe_float` (:345), `_calculate_scoring_penalties` (:363), `_parse_shutter_speed` (:710),
`_determine_photo_category` (:726) and the config side it leans on: `ScoringConfig.get_weights`
(config/scoring_config.py:301), `determine_category` (:792) and `CategoryFilter.matches`
(config/category_filter.py:55).

Shape of this implementation: the config is compiled ONCE (categories sorted by priority with
lower-cased tag lists, one weight row of 16 metric columns + flags per category), the per-photo
dicts are decoded into float64 columns, and the arithmetic runs over the whole batch as NumPy
float64 column operations in the reference's order of operations, so every result is bit-identical
to the scalar reference (+, -, *, / and comparisons are IEEE-exact element-wise; Python's
`min` / `max` — which are not NaN-symmetric — are restated as `where` selects).
`calculate_aggregate_logic(m)` is the batch of one.

SURVEY.md §8 rows a18 (aggregate) and (f) rank 4 (vectorised over the batch).  Pinned by
tests/golden/aggregate_golden.json (outputs of the unmodified reference class).
"""
from __future__ import annotations

import numpy as np

# metric columns in the order the reference sums them (scorer.py:882-904): Python adds the weighted terms
# in dict order, and float addition is not associative
METRICS = ("aesthetic", "quality", "face_quality", "face_sharpness", "eye_sharpness", "tech_sharpness", "composition",
           "power_point", "leading_lines", "exposure", "color", "contrast", "dynamic_range", "saturation", "noise",
           "isolation")
_M = {name: k for k, name in enumerate(METRICS)}

_NUMERIC_FILTERS = (("face_ratio", "face_ratio"), ("face_count", "face_count"), ("iso", "iso"),
                    ("shutter_speed", "shutter_speed"), ("luminance", "mean_luminance"),
                    ("focal_length", "focal_length"), ("f_stop", "f_stop"))
_BOOL_FILTERS = ("has_face", "is_monochrome", "is_silhouette", "is_group_portrait")


def safe_float(val, default=5.0):
    """scorer.py:345-360: numbers (and numeric strings) inside [-100, 100], else the default."""
    if val is None or isinstance(val, bytes):
        return default
    if isinstance(val, str):
        try:
            val = float(val)
        except ValueError:
            return default
    if isinstance(val, (int, float)):
        if val < -100 or val > 100:
            return default
        return float(val)
    return default


def _category_float(val, default=0.0):
    """The narrower helper inside `_determine_photo_category` (scorer.py:738-744): no string parsing."""
    if val is None or isinstance(val, bytes):
        return default
    if isinstance(val, (int, float)):
        return float(val) if -100 <= val <= 100 else default
    return default


def parse_shutter_speed(val):
    """scorer.py:710-724: seconds from a number or a string like '1/500'."""
    if val is None:
        return None
    if isinstance(val, (int, float)):
        return float(val)
    if isinstance(val, str):
        try:
            if "/" in val:
                num, denom = val.split("/")
                return float(num) / float(denom)
            return float(val)
        except (ValueError, ZeroDivisionError):
            return None
    return None


def _pmin(a, b):
    """Python's min(a, b) element-wise: a unless b < a (NaN in b never wins, NaN in a always stays)."""
    return np.where(np.less(b, a), b, a)


def _pmax(a, b):
    return np.where(np.greater(b, a), b, a)


class _Category:
    __slots__ = ("name", "filters", "numeric", "bools", "required", "excluded", "match_all", "weights", "row", "blink",
                 "skip_clip", "noise_tol", "clip_mult", "skip_oversat", "bonus", "aes_extra", "w_aes", "listed")

    def __init__(self, cat: dict, listed: bool = True):
        self.name = cat.get("name")
        self.listed = listed          # False: a name without an entry in the config (empty weights)
        f = cat.get("filters", {}) or {}
        self.filters = f
        self.numeric = [(key, f.get(field + "_min"), f.get(field + "_max")) for field, key in _NUMERIC_FILTERS
                        if f.get(field + "_min") is not None or f.get(field + "_max") is not None]
        self.bools = [(field, f[field]) for field in _BOOL_FILTERS if f.get(field) is not None]
        self.required = [t.lower() for t in f.get("required_tags", [])]
        self.excluded = [t.lower() for t in f.get("excluded_tags", [])]
        self.match_all = f.get("tag_match_mode", "any") != "any"
        self.weights = _convert_weights(cat)
        w = self.weights
        self.row = np.array([w.get(name, 0.0) for name in METRICS], dtype=np.float64)
        n = self.name
        self.blink = w.get("_apply_blink_penalty", n in ("portrait", "portrait_bw", "group_portrait"))
        self.skip_clip = w.get("_skip_clipping_penalty", n == "silhouette")
        self.noise_tol = w.get("noise_tolerance_multiplier", 1.0)
        self.clip_mult = w.get("_clipping_multiplier", 1.5 if n == "default" else 1.0)
        self.skip_oversat = w.get("_skip_oversaturation_penalty", n in ("night", "astro", "concert"))
        self.bonus = w.get("bonus", 0.0)
        self.aes_extra = w.get("quality", 0.0)
        self.w_aes = w.get("aesthetic", 0)

    def matches(self, pd: dict) -> bool:
        if not self.filters:
            return True
        for key, lo, hi in self.numeric:
            actual = pd.get(key)
            if actual is None:
                return False          # a constraint on a value we do not have cannot be verified
            if lo is not None and actual < lo:
                return False
            if hi is not None and actual > hi:
                return False
        for field, required in self.bools:
            actual = (pd.get("face_count") or 0) > 0 if field == "has_face" else bool(pd.get(field, 0))
            if actual != required:
                return False
        if self.required or self.excluded:
            tags = [t.strip().lower() for t in (pd.get("tags") or "").split(",") if t.strip()]
            if self.required:
                hit = (all if self.match_all else any)(t in tags for t in self.required)
                if not hit:
                    return False
            if self.excluded and any(t in tags for t in self.excluded):
                return False
        return True


def _convert_weights(cat: dict) -> dict:
    """get_weights (scoring_config.py:301-338): '<x>_percent' -> fraction, renormalised when the fractions do not
    sum to 1 within 0.001, modifiers merged on top."""
    out, keys = {}, []
    for key, value in (cat.get("weights", {}) or {}).items():
        if key.endswith("_percent"):
            out[key[:-8]] = value / 100
            keys.append(key[:-8])
        else:
            out[key] = value
    if keys:
        total = sum(out[k] for k in keys)
        if total > 0 and abs(total - 1.0) > 0.001:
            for k in keys:
                out[k] = out[k] / total
    out.update(cat.get("modifiers", {}) or {})
    return out


class AggregateScorer:
    """Compiled form of the scoring sections of a ScoringConfig (ours or the reference's: only `.config` is read)."""

    def __init__(self, config):
        cfg = config.config if hasattr(config, "config") else (config or {})
        self.version_hash = getattr(config, "version_hash", None)
        cats = cfg.get("categories", [])
        self._by_name = {}
        for cat in cats:                                   # get_weights takes the FIRST entry of a name
            self._by_name.setdefault(cat.get("name"), _Category(cat))
        # determine_category walks a stable sort by priority, every entry with its own filters
        self._ordered = [_Category(c) for c in sorted(cats, key=lambda c: c.get("priority", 100))]
        self.default_category = cfg.get("viewer", {}).get("default_category", "default")
        scoring = cfg.get("scoring", {})
        self.score_min = scoring.get("score_min", 0.0)
        self.score_max = scoring.get("score_max", 10.0)
        thr = cfg.get("thresholds", {})
        self.blink_penalty = (thr.get("blink_penalty_percent", 0) or 50) / 100
        exif = cfg.get("exif_adjustments", {"iso_sharpness_compensation": True, "aperture_isolation_boost": True})
        self.iso_comp = exif.get("iso_sharpness_compensation", True)
        self.aperture_boost = exif.get("aperture_isolation_boost", True)
        self.silhouette_detection = cfg.get("exposure", {}).get("silhouette_detection", True)
        pen = cfg.get("penalties", {"noise_sigma_threshold": 4.0, "noise_max_penalty_points": 1.5,
                                    "noise_penalty_per_sigma": 0.3, "bimodality_threshold": 2.5,
                                    "bimodality_penalty_points": 0.5, "leading_lines_blend_percent": 30})
        self.noise_thr = pen.get("noise_sigma_threshold", 4.0)
        self.noise_max = pen.get("noise_max_penalty_points", 1.5)
        self.noise_rate = pen.get("noise_penalty_per_sigma", 0.3)
        self.bimod_thr = pen.get("bimodality_threshold", 2.5)
        self.bimod_pts = pen.get("bimodality_penalty_points", 0.5)
        self.oversat_thr = pen.get("oversaturation_threshold", 0.9)
        self.oversat_pts = pen.get("oversaturation_penalty_points", 0.5)
        self.ll_blend = pen.get("leading_lines_blend_percent", 30) / 100

    # -- config side --------------------------------------------------------------------------------
    def weights_of(self, category) -> dict:
        cat = self._by_name.get(category)
        return cat.weights if cat is not None and cat.listed else {}

    def match_category(self, photo_data: dict) -> str:
        for cat in self._ordered:
            if cat.matches(photo_data):
                return cat.name
        return self.default_category

    def category_of(self, m: dict) -> str:
        """_determine_photo_category (scorer.py:726-767)."""
        return self.match_category({
            "tags": m.get("tags", "") or "",
            "face_count": int(_category_float(m.get("face_count"), 0)),
            "face_ratio": _category_float(m.get("face_ratio"), 0),
            "is_silhouette": m.get("is_silhouette", 0),
            "is_group_portrait": m.get("is_group_portrait", 0),
            "is_monochrome": m.get("is_monochrome", 0),
            "mean_luminance": _category_float(m.get("mean_luminance"), 0.5),
            "iso": m.get("iso"),
            "shutter_speed": parse_shutter_speed(m.get("shutter_speed")),
            "focal_length": m.get("focal_length"),
            "f_stop": m.get("f_stop"),
        })

    # -- the batch ----------------------------------------------------------------------------------
    def score_batch(self, rows):
        """rows: sequence of metric dicts -> (float64 scores [n], category names [n])."""
        n = len(rows)
        if n == 0:
            return np.zeros(0, np.float64), []
        names = [self.category_of(m) for m in rows]
        # a category without an entry (e.g. the viewer's default name) scores with empty weights but keeps the
        # name-based defaults of the flags
        cats = [self._by_name.get(c) or self._by_name.setdefault(c, _Category({"name": c}, listed=False)) for c in names]

        def col(key, default):
            return np.array([safe_float(m.get(key), default) for m in rows], dtype=np.float64)

        def flag(values):
            return np.array([bool(v) for v in values], dtype=bool)

        # EXIF-aware adjustments: the logarithm is taken with the scalar call the reference makes, row by row
        sharp = col("tech_sharpness", 5.0)
        if self.iso_comp:
            for i, m in enumerate(rows):
                iso = safe_float(m.get("iso"), None)
                if iso and iso > 800:
                    sharp[i] = min(10.0, sharp[i] + 0.5 * np.log2(iso / 800))
        isolation = np.array([m.get("isolation_bonus", 1.0) for m in rows], dtype=np.float64)
        if self.aperture_boost:
            f_stop = np.array([safe_float(m.get("f_stop"), None) or np.nan for m in rows], dtype=np.float64)
            mult = np.where(f_stop <= 2.0, 1.5, 1.3)
            isolation = np.where(f_stop <= 2.8, _pmin(3.0, isolation * mult), isolation)
        isolation_score = _pmin(10.0, (isolation - 1.0) * 5.0)

        # clipping penalty (skipped for silhouettes when silhouette detection is on)
        clip_pen = np.zeros(n, np.float64)
        for i, m in enumerate(rows):
            if not (m.get("is_silhouette", 0) if self.silhouette_detection else False):
                sc, hc = m.get("shadow_clipped", 0), m.get("highlight_clipped", 0)
                if sc or hc:
                    clip_pen[i] = (sc * 0.5) + (hc * 1.0)
        dynamic_range = _pmin(10.0, col("histogram_spread", 0) / 6.0)

        noise_sigma = col("noise_sigma", 0)
        noise_pen = np.where(noise_sigma > self.noise_thr, _pmin(self.noise_max, (noise_sigma - self.noise_thr) * self.noise_rate), 0.0)
        bimod_pen = np.where(col("histogram_bimodality", 0) > self.bimod_thr, float(self.bimod_pts), 0.0)
        mean_sat = col("mean_saturation", 0)
        oversat_pen = np.where(mean_sat > self.oversat_thr, float(self.oversat_pts), 0.0)
        leading = _pmin(10.0, col("leading_lines_score", 0) * 1.77)

        W = np.stack([c.row for c in cats])
        aes = col("aesthetic", 5.0)
        color = np.where(flag(m.get("is_monochrome", 0) for m in rows), 5.0, col("color_score", 5.0))
        comp_raw = col("comp_score", 5.0)
        portraitish = np.array([c in ("portrait", "group_portrait") for c in names], dtype=bool)
        comp = np.where(~portraitish & (leading > 0), _pmin(10.0, comp_raw + (leading * self.ll_blend)), comp_raw)
        w_aes = np.array([c.w_aes for c in cats], dtype=np.float64)
        aes_extra = np.array([c.aes_extra for c in cats], dtype=np.float64)
        with np.errstate(all="ignore"):
            aes_col = np.where(w_aes > 0, aes + aes_extra / _pmax(w_aes, 0.01), aes)

        values = np.empty((len(METRICS), n), np.float64)
        values[_M["aesthetic"]] = aes_col
        values[_M["quality"]] = 0.0
        values[_M["face_quality"]] = col("face_quality", 5.0)
        values[_M["face_sharpness"]] = col("face_sharpness", 5.0)
        values[_M["eye_sharpness"]] = col("eye_sharpness", 5.0)
        values[_M["tech_sharpness"]] = sharp
        values[_M["composition"]] = comp
        values[_M["power_point"]] = col("power_point_score", 5.0)
        values[_M["leading_lines"]] = leading
        values[_M["exposure"]] = col("exposure_score", 5.0)
        values[_M["color"]] = color
        values[_M["contrast"]] = col("contrast_score", 5.0)
        values[_M["dynamic_range"]] = dynamic_range
        values[_M["saturation"]] = _pmin(10.0, col("mean_saturation", 0.5) * 10.0)
        values[_M["noise"]] = _pmax(0.0, _pmin(10.0, 10.0 - noise_sigma * 0.7))
        values[_M["isolation"]] = isolation_score

        score = np.zeros(n, np.float64)
        for k in range(len(METRICS)):
            w = W[:, k]
            clamped = _pmax(0.0, _pmin(10.0, values[k]))
            score = np.where(w > 0, score + clamped * w, score)

        blink = flag(c.blink for c in cats) & flag(m.get("is_blink") for m in rows)
        score = np.where(blink, score * self.blink_penalty, score)
        score = score + np.array([c.bonus for c in cats], dtype=np.float64)
        clip_mult = np.array([c.clip_mult for c in cats], dtype=np.float64)
        score = np.where(flag(c.skip_clip for c in cats), score, score - clip_pen * clip_mult)
        score = score - noise_pen * np.array([c.noise_tol for c in cats], dtype=np.float64)
        score = score - bimod_pen
        score = np.where(flag(c.skip_oversat for c in cats), score, score - oversat_pen)
        return _pmin(self.score_max, _pmax(self.score_min, score)), names

    def score(self, m: dict):
        s, c = self.score_batch([m])
        return float(s[0]), c[0]


def calculate_aggregate_logic(m: dict, config):
    """Drop-in for `Facet.calculate_aggregate_logic(m, config)` -> (aggregate, category)."""
    scorer = config._scoring() if hasattr(config, "_scoring") else AggregateScorer(config)
    return scorer.score(m)
