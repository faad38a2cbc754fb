"""`BatchProcessor` for the in-scope part of the per-image pass (mirrors the role of
processing/batch_processor.py:27-658).

The reference collects `batch_size` loaded items (dicts with 'path', 'pil_img', 'img_cv',
'clip_input'), runs CLIP on the batch, then loops over the images on the CPU
(`_process_batch`, batch_processor.py:169-360).  Here a batch is grouped by frame shape and every
group goes through ONE device pass (`Facet.score_images`: technical metrics, perceptual hash, CLIP
preprocess, ViT-L/14 + aesthetic head + tags).  Results come back in input order as dicts with the
reference's column names (batch_processor.py:298-355); columns owned by analyzers that are out of
scope (faces, composition, EXIF, aggregate) are filled by the optional `finish` callback, which is
where the reference's own `calculate_aggregate_logic` plugs in.

Errors follow the reference's convention: a bad item never poisons the batch, it becomes
`{'path': ..., 'error': ...}` (batch_processor.py:109,359).
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np


class BatchProcessor:
    def __init__(self, scorer, batch_size=16, num_workers=4, finish=None, mono_threshold=None):
        self.scorer = scorer
        self.batch_size = int(batch_size)
        self.num_workers = int(num_workers)
        self.finish = finish
        cfg = getattr(scorer, "config", None)
        if mono_threshold is None and cfg is not None:
            mono_threshold = cfg.get_monochrome_settings().get("saturation_threshold_percent", 10) / 100
        self.mono_threshold = 0.10 if mono_threshold is None else mono_threshold
        self.metrics = {"images_processed": 0, "images_failed": 0, "batches": 0}

    def _tag_params(self):
        cfg = getattr(self.scorer, "config", None)
        if cfg is None:
            return 0.22, 5
        from ..utils.tags import get_tag_params
        return get_tag_params(cfg)

    def _process_batch(self, batch):
        """batch: list of dicts with 'path' and 'img_cv' ([H,W,3] uint8 BGR).  Returns one result per item."""
        results = [None] * len(batch)
        groups = defaultdict(list)
        for i, item in enumerate(batch):
            img = item.get("img_cv") if isinstance(item, dict) else None
            if "error" in item:
                results[i] = {"path": item.get("path"), "error": item["error"]}
            elif not isinstance(img, np.ndarray) or img.ndim != 3 or img.shape[2] != 3 or img.dtype != np.uint8:
                results[i] = {"path": item.get("path"), "error": "Failed to load image"}
            elif img.shape[0] < 2 or img.shape[1] < 2:
                results[i] = {"path": item.get("path"), "error": "image smaller than 2x2"}
            else:
                groups[img.shape[:2]].append(i)
        thr, max_tags = self._tag_params()
        for _shape, idxs in groups.items():
            try:
                frames = np.stack([batch[i]["img_cv"] for i in idxs])
                scored = self.scorer.score_images(frames, mono_threshold=self.mono_threshold, tag_threshold=thr, max_tags=max_tags)
                for i, res in zip(idxs, scored):
                    res["path"] = batch[i].get("path")
                    res["filename"] = str(batch[i].get("path", "")).rsplit("/", 1)[-1]
                    if self.finish is not None:
                        res = self.finish(batch[i], res)
                    results[i] = res
            except Exception as exc:  # a CUDA failure is loud, but it is reported per item like the reference does
                for i in idxs:
                    results[i] = {"path": batch[i].get("path"), "error": str(exc)}
        self.metrics["batches"] += 1
        self.metrics["images_processed"] += sum(1 for r in results if r and "error" not in r)
        self.metrics["images_failed"] += sum(1 for r in results if r and "error" in r)
        return results

    def process_items(self, items):
        """Stream items through `_process_batch` in chunks of batch_size; yields results in order."""
        chunk = []
        for item in items:
            chunk.append(item)
            if len(chunk) == self.batch_size:
                yield from self._process_batch(chunk)
                chunk = []
        if chunk:
            yield from self._process_batch(chunk)
