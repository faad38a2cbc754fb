"""Minimal reader of the reference's ``scoring_config.json`` for the settings the scoring pass uses.

Mirrors the getters of config/scoring_config.py (:455 monochrome, :461 tagging, :473 clip,
:482 burst, :490 duplicate, :730 tag vocabulary) with the same defaults.  The reference's full
``ScoringConfig`` (categories, weights, validation) is outside the hot path; when it is
importable a caller can pass it instead — only these getters are used.
"""
from __future__ import annotations

import hashlib
import json
import os


class ScoringConfig:
    def __init__(self, config_path=None, validate=False):
        self.config_path = config_path or "scoring_config.json"
        if not os.path.exists(self.config_path):
            raise FileNotFoundError(f"Config file not found: {self.config_path}")
        with open(self.config_path, "r") as f:
            self.config = json.load(f)
        self.version_hash = hashlib.md5(json.dumps(self.config, sort_keys=True).encode()).hexdigest()[:12]

    @classmethod
    def from_dict(cls, cfg: dict) -> "ScoringConfig":
        self = cls.__new__(cls)
        self.config_path = None
        self.config = dict(cfg)
        self.version_hash = hashlib.md5(json.dumps(self.config, sort_keys=True).encode()).hexdigest()[:12]
        return self

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
        return self.config.get("exposure", {"shadow_clip_threshold_percent": 15, "highlight_clip_threshold_percent": 10})

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
