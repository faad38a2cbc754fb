"""Minimal reader of the reference's ``scoring_config.json`` for the settings the scoring pass uses.

Mirrors the getters of config/scoring_config.py (:455 monochrome, :461 tagging, :473 clip,
:482 burst, :490 duplicate, :730 tag vocabulary; :301 weights, :340 limits, :349 threshold,
:357 composition, :407 exif, :414 exposure, :422 penalties, :782 categories, :792
determine_category) with the same defaults.  Validation, VRAM profiles and the model registry of
the reference's ``ScoringConfig`` are outside the scoring pass; when the reference class is
importable a caller can pass it instead — only these getters and ``.config`` are used.
"""
from __future__ import annotations

import hashlib
import json
import os


# weight columns the validator accepts (config/category_filter.py:29-33) and its tolerance (scoring_config.py:39)
VALID_WEIGHT_COLUMNS = ("aesthetic", "face_quality", "eye_sharpness", "tech_sharpness", "exposure", "composition",
                        "color", "quality", "contrast", "dynamic_range", "isolation", "leading_lines")
NORMALIZATION_TOLERANCE = 5


class ScoringConfig:
    def __init__(self, config_path=None, validate=True):
        self.config_path = config_path or "scoring_config.json"
        if not os.path.exists(self.config_path):
            raise FileNotFoundError(f"Config file not found: {self.config_path}")
        try:
            with open(self.config_path, "r") as f:
                self.config = json.load(f)
        except Exception as exc:
            raise ValueError(f"Could not load config from {self.config_path}: {exc}")
        if "categories" not in self.config:
            raise ValueError(f"Config file {self.config_path} is not v4.0 format (missing 'categories' array).")
        self.version_hash = self._compute_version_hash()
        if validate:
            self.validate_weights(verbose=False)

    @classmethod
    def from_dict(cls, cfg: dict, validate=False) -> "ScoringConfig":
        self = cls.__new__(cls)
        self.config_path = None
        self.config = dict(cfg)
        self.version_hash = self._compute_version_hash()
        if validate:
            self.validate_weights(verbose=False)
        return self

    def _compute_version_hash(self):
        return hashlib.md5(json.dumps(self.config, sort_keys=True).encode()).hexdigest()[:12]

    def save_config(self):
        with open(self.config_path, "w") as f:
            json.dump(self.config, f, indent=2)
            f.write("\n")

    # -- weight validation (scoring_config.py:130-293): the reference auto-corrects the '<x>_percent' weights of
    # every category when the config is loaded, writes the corrected file back and re-hashes it; scores depend
    # on the corrected values, so the same corrections are applied here
    @staticmethod
    def normalize_weights_to_100(weights_dict, skip_within_tolerance=True):
        if not weights_dict:
            return None
        total = sum(weights_dict.values())
        if total == 0 or abs(total - 100) <= 0.01:
            return None
        if skip_within_tolerance and abs(total - 100) <= NORMALIZATION_TOLERANCE:
            return None
        scale = 100.0 / total
        out, running = {}, 0
        keys = sorted(weights_dict.keys(), key=lambda k: weights_dict[k], reverse=True)
        for i, key in enumerate(keys):     # the smallest weight takes the remainder so the sum is exactly 100
            val = max(0, 100 - running) if i == len(keys) - 1 else round(weights_dict[key] * scale)
            running += val
            out[key] = val
        return out

    def validate_weights(self, verbose=True):
        corrected = []
        for cat in self.config.get("categories", []):
            weights = cat.get("weights", {})
            if not isinstance(weights, dict):
                continue
            items, invalid = {}, []
            for key, value in weights.items():
                if key.endswith("_percent") and isinstance(value, (int, float)):
                    (items.__setitem__(key, value) if key[:-8] in VALID_WEIGHT_COLUMNS else invalid.append(key))
            if not items:
                continue
            changed = bool(invalid)
            for key in invalid:
                del weights[key]
            for col in VALID_WEIGHT_COLUMNS:
                key = col + "_percent"
                if key not in weights:
                    weights[key] = items[key] = 0
                    changed = True
            if all(v <= 1 for v in items.values()) and sum(items.values()) <= 1.01 and len(items) > 1:
                for key, value in items.items():          # fractions -> percentages
                    new = round(value * 100)
                    if new != value:
                        weights[key] = items[key] = new
                        changed = True
            for key, value in items.items():
                if value < 0:
                    weights[key] = items[key] = 0
                    changed = True
            for key, value in items.items():
                if isinstance(value, float) and value != int(value):
                    weights[key] = items[key] = round(value)
                    changed = True
            new_weights = self.normalize_weights_to_100(items)
            if new_weights:
                for key in items:
                    changed = changed or new_weights[key] != items[key]
                    weights[key] = new_weights[key]
            if changed:
                corrected.append(cat.get("name", "unnamed"))
        if corrected:
            if self.config_path is not None:
                self.save_config()
            self.refresh()
            if verbose:
                print(f"Corrected weights of {corrected}; saved to {self.config_path}")
        return len(corrected) == 0, corrected

    def get_monochrome_settings(self):
        return self.config.get("monochrome_detection", {"saturation_threshold_percent": 10})

    def get_tagging_settings(self):
        return self.config.get("tagging", {"enabled": True, "max_tags": 5})

    def get_clip_settings(self):
        return self.config.get("models", {}).get(
            "clip", {"model_name": "ViT-L-14", "pretrained": "laion2b_s32b_b82k", "similarity_threshold_percent": 22})

    def get_burst_detection_settings(self):
        return self.config.get("burst_detection", {"similarity_threshold_percent": 88, "time_window_minutes": 60,
                                                   "rapid_burst_seconds": 5})

    def get_duplicate_detection_settings(self):
        return self.config.get("duplicate_detection", {"similarity_threshold_percent": 90})

    def get_exposure_settings(self):
        return self.config.get("exposure", {"shadow_clip_threshold_percent": 15, "highlight_clip_threshold_percent": 10,
                                            "silhouette_detection": True})

    # -- aggregate scoring (consumed by processing/aggregate.py) ---------------------------------------
    def get_scoring_limits(self):
        scoring = self.config.get("scoring", {})
        return {"score_min": scoring.get("score_min", 0.0), "score_max": scoring.get("score_max", 10.0),
                "score_precision": scoring.get("score_precision", 2)}

    def get_threshold(self, name):
        return self.config.get("thresholds", {}).get(name, 0)

    def get_composition_weights(self):
        return self.config.get("composition", {})

    def get_exif_adjustments(self):
        return self.config.get("exif_adjustments", {"iso_sharpness_compensation": True, "aperture_isolation_boost": True})

    def get_penalty_settings(self):
        return self.config.get("penalties", {
            "noise_sigma_threshold": 4.0, "noise_max_penalty_points": 1.5, "noise_penalty_per_sigma": 0.3,
            "bimodality_threshold": 2.5, "bimodality_penalty_points": 0.5, "leading_lines_blend_percent": 30})

    def get_categories(self):
        return sorted(self.config.get("categories", []), key=lambda c: c.get("priority", 100))

    def _scoring(self):
        """The compiled scoring sections (processing/aggregate.py), built on first use.  Edits made to `self.config`
        afterwards are picked up by `refresh()` (validate_weights calls it)."""
        from .processing.aggregate import AggregateScorer
        cached = getattr(self, "_aggregate_scorer", None)
        if cached is None or cached.version_hash != self.version_hash:
            cached = self._aggregate_scorer = AggregateScorer(self)
        return cached

    def refresh(self):
        """Re-hash and recompile after in-place edits of `self.config` (e.g. a weight optimiser)."""
        self.version_hash = self._compute_version_hash()
        self._aggregate_scorer = None

    def get_weights(self, category):
        """scoring_config.py:301-338, always read from the live config."""
        from .processing.aggregate import _convert_weights
        for cat in self.config.get("categories", []):
            if cat.get("name") == category:
                return _convert_weights(cat)
        return {}

    def determine_category(self, photo_data: dict) -> str:
        return self._scoring().match_category(photo_data)

    def get_tag_vocabulary(self):
        vocab = {}
        for cat in self.config.get("categories", []):
            tags = cat.get("tags", {})
            if isinstance(tags, dict):
                vocab.update(tags)
        standalone = self.config.get("standalone_tags", {})
        if isinstance(standalone, dict):
            vocab.update(standalone)
        return vocab

    def get_category_tags(self, category):
        """scoring_config.py:752-766: the tag names (keys of the category's `tags` dict) that trigger a category."""
        for cat in self.config.get("categories", []):
            if cat.get("name") == category:
                tags = cat.get("tags", {})
                if isinstance(tags, dict):
                    return list(tags.keys())
        return []

    def get_art_tags(self):
        """scoring_config.py:748-750."""
        return set(self.get_category_tags("art"))
