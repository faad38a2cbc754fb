"""`Facet`-shaped scoring engine for the legacy-profile pass (mirrors processing/scorer.py).

Only the members the per-image pass touches are provided:
  .device, .config, .preprocess(pil_img), .get_aesthetic_and_quality_batch(pil_images, clip_inputs),
  .get_aesthetic_with_embedding / .get_aesthetic_and_quality (single-image twins, scorer.py:587-638),
  .score_from_embedding, .tagger, .tech_analyzer
plus `score_images`, the batched entry the data-parallel driver uses: one call runs the technical
pass, the CLIP preprocess, the ViT tower + heads + tag similarities for a same-shaped batch that is
already on the device.  Aggregate scoring (`calculate_aggregate_logic`, scorer.py:769) is the
host-side consumer of the dicts produced here (processing/aggregate.py); SQLite writes stay with
the reference.
"""
from __future__ import annotations

import numpy as np

from .. import _lib, ops
from ..analyzers import _closed_form as cf
from ..analyzers.technical import TechnicalAnalyzer
from ..models.clip_vit import ClipVitL14
from ..models.tagger import CLIPTagger
from ..utils import resample


def _aesthetic_from_raw(raw: float) -> float:
    """scorer.py:669: max(0, min(10, (raw + 1) * 5))."""
    return max(0.0, min(10.0, (float(raw) + 1) * 5))


class Facet:
    def __init__(self, state_dict, config=None, text_embeddings=None, tag_names=None, mean=resample.LAION_MEAN,
                 std=resample.LAION_STD, device=None):
        torch = _lib.require_cuda()
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.config = config
        self.mean, self.std = tuple(mean), tuple(std)
        self.tech_analyzer = TechnicalAnalyzer()
        self.tagger = None
        if text_embeddings is not None:
            self.tagger = CLIPTagger(None, self.device, config=config, text_embeddings=text_embeddings, tag_names=tag_names)
        self.model = ClipVitL14(state_dict, tag_embeddings=text_embeddings, device=self.device)
        self._head = {k: state_dict[k].detach().to(self.device, torch.float32)
                      for k in ("aesthetic_head.0.weight", "aesthetic_head.0.bias", "aesthetic_head.2.weight", "aesthetic_head.2.bias")}
        self._head_tuple = (self._head["aesthetic_head.0.weight"].contiguous(), self._head["aesthetic_head.0.bias"].contiguous(),
                            self._head["aesthetic_head.2.weight"].reshape(-1).contiguous(), self._head["aesthetic_head.2.bias"].contiguous())

    # -- scorer.preprocess: PIL RGB image -> float32 [3,224,224] (the transform open_clip returns) ----------
    def preprocess(self, pil_img):
        rgb = np.asarray(pil_img.convert("RGB"))
        return ops.clip_preprocess(rgb, mean=self.mean, std=self.std, rgb_order=True)[0]

    def preprocess_batch(self, images, rgb_order=False):
        return ops.clip_preprocess(images, mean=self.mean, std=self.std, rgb_order=rgb_order)

    # -- scorer.py:640-673 -------------------------------------------------------------------------------
    def get_aesthetic_and_quality_batch(self, pil_images, clip_inputs=None):
        import torch
        if clip_inputs is not None:
            inputs = clip_inputs.to(self.device)
        else:
            inputs = torch.stack([self.preprocess(img) for img in pil_images])
        out = self.model.encode(inputs)
        raw = out["aesthetic_raw"].cpu().numpy()
        emb = out["embedding"].cpu().numpy()
        return [(_aesthetic_from_raw(raw[i]), emb[i].astype(np.float32).tobytes(), None, "clip-mlp")
                for i in range(len(raw))]

    def get_aesthetic_with_embedding(self, image_pil):
        a, e, _, _ = self.get_aesthetic_and_quality_batch([image_pil])[0]
        return a, e

    def get_aesthetic_score(self, image_pil):
        return self.get_aesthetic_with_embedding(image_pil)[0]

    def get_aesthetic_and_quality(self, pil_img):
        a, e = self.get_aesthetic_with_embedding(pil_img)
        return a, e, None, "clip-mlp"

    def score_from_embedding(self, embedding_bytes):
        """scorer.py:620-629: the head applied to a stored (normalised) embedding."""
        import torch
        f = torch.from_numpy(np.frombuffer(embedding_bytes, dtype=np.float32).copy()).to(self.device).unsqueeze(0)
        raw, _ = ops.embedding_heads(f, head=self._head_tuple)
        return _aesthetic_from_raw(float(raw.cpu()[0]))

    # -- scorer.py:726-950 ---------------------------------------------------------------------------------
    def calculate_aggregate_logic(self, m, config=None):
        from .aggregate import calculate_aggregate_logic
        cfg = config or self.config
        if cfg is None:
            raise ValueError("calculate_aggregate_logic needs a ScoringConfig (Facet(config=...))")
        return calculate_aggregate_logic(m, cfg)

    def _determine_photo_category(self, m, cfg=None):
        from .aggregate import AggregateScorer
        cfg = cfg or self.config
        return (cfg._scoring() if hasattr(cfg, "_scoring") else AggregateScorer(cfg)).category_of(m)

    # -- batched device-resident entry -------------------------------------------------------------------
    def pixel_passes_device(self, images, rgb_order=False, with_phash=True, with_thumbnails=False):
        """The per-frame pixel work for a same-shaped CUDA uint8 batch [n,H,W,3]: technical pass (+ luma plane), perceptual
        hash, CLIP preprocess.  Returns device tensors {hist256, sums, derived, phash, clip_in}; nothing is synchronised.
        with_thumbnails adds `thumbnails` (uint8 [n,h,w,3] RGB, Pillow's 640-px LANCZOS thumbnail, scorer.py:1681-1686):
        when Pillow's plan starts with a (4, 4) box reduction (24 MP frames) the technical pass emits it from its own
        read of the frame, so the thumbnail costs no further read."""
        n, h, w, _ = images.shape
        luma = None
        box = None
        if with_thumbnails and ops.thumbnail_reduces_by_4(h, w):
            import torch
            box = torch.empty((n, (h + 3) // 4, (w + 3) // 4, 3), dtype=torch.uint8, device=images.device)
        if with_phash and ops.phash_uses_luma_plane(h, w):
            # the technical pass also emits Pillow's luma plane, so the hash never re-reads the frame
            import torch
            luma = torch.empty((n, h, w), dtype=torch.uint8, device=images.device)
        hist, hs, sums, derived = ops.tech_stats_raw(images, rgb_order=rgb_order, luma_out=luma, box_out=box)
        hashes = ops.phash(images, rgb_order=rgb_order, device_only=True, luma=luma) if with_phash else None
        clip_in = ops.clip_preprocess(images, mean=self.mean, std=self.std, rgb_order=rgb_order)
        out = {"hist256": hist, "sums": sums, "derived": derived, "phash": hashes, "clip_in": clip_in}
        if with_thumbnails:
            out["thumbnails"] = ops.thumbnails(images, rgb_order=rgb_order, to_rgb=True, reduced=box)
        return out

    def score_images_device(self, images, rgb_order=False, with_phash=True):
        """images: CUDA uint8 [n,H,W,3].  Enqueues technical pass, perceptual hash, preprocess and ViT;
        returns device tensors (nothing is copied to the host)."""
        px = self.pixel_passes_device(images, rgb_order=rgb_order, with_phash=with_phash)
        vit = self.model.encode(px.pop("clip_in"))
        return {**px, **vit}

    def results_from_host(self, h, w, hist, sums, der, raw, emb, sims, hashes, mono_threshold=0.10, tag_threshold=0.22,
                          max_tags=5, shadow_threshold=0.15, highlight_threshold=0.10):
        """Host half of the pass: the small per-image device results (as host arrays: hist [n,256] int64, sums [n,4] int64,
        derived [n,4] float64, aesthetic_raw [n], embedding [n,768] float32, tag_sims [n,T] or None, hashes = list of hex
        strings or Nones) -> result dicts with the reference's metric keys (processing/batch_processor.py:298-355, the
        analyzer-derived subset)."""
        n = len(raw)
        metrics = cf.all_metrics_batch(h, w, hist, sums[:, 0], sums[:, 1], sums[:, 2], der[:, 0], der[:, 1],
                                       mono_threshold=mono_threshold, shadow_threshold=shadow_threshold,
                                       highlight_threshold=highlight_threshold)      # the 7 analyzer dicts per image, one pass
        results = []
        for i in range(n):
            m = metrics[i]
            sharp, color, hd = m["sharpness"], m["color"], m["histogram"]
            mono, dr, nz, ct = m["monochrome"], m["dynamic_range"], m["noise"], m["contrast"]
            tags = None
            if self.tagger is not None and sims is not None:
                tl = self.tagger.get_tags_from_similarities(sims[i], tag_threshold, max_tags)
                tags = ",".join(tl) if tl else None
            aest = _aesthetic_from_raw(raw[i])
            results.append({
                "image_width": w, "image_height": h,
                # the aggregate is computed from the unrounded analyzer values (batch_processor.py:271-294);
                # BatchProcessor pops these four
                "aesthetic_unrounded": aest, "tech_sharpness_unrounded": sharp["normalized"],
                "color_score_unrounded": color["normalized"], "exposure_score_unrounded": hd["exposure_score"],
                "aesthetic": round(aest, 2),
                "tech_sharpness": round(sharp["normalized"], 2),
                "color_score": round(color["normalized"], 2),
                "exposure_score": round(hd["exposure_score"], 2),
                "clip_embedding": emb[i].astype(np.float32).tobytes(),
                "raw_sharpness_variance": float(sharp["raw_variance"]),
                "histogram_data": hd["histogram_bytes"],
                "histogram_spread": float(hd["spread"]),
                "mean_luminance": float(hd["mean_luminance"]),
                "histogram_bimodality": float(hd["bimodality"]),
                "raw_color_entropy": float(color["raw_entropy"]),
                "shadow_clipped": hd["shadow_clipped"], "highlight_clipped": hd["highlight_clipped"],
                "is_silhouette": hd["is_silhouette"],
                "is_monochrome": mono["is_monochrome"], "mean_saturation": mono["mean_saturation"],
                "dynamic_range_stops": dr["dynamic_range_stops"], "noise_sigma": nz["noise_sigma"],
                "contrast_score": ct["contrast_score"],
                "tags": tags, "quality_score": None, "scoring_model": "clip-mlp", "phash": hashes[i],
            })
        return results

    def score_images(self, images, rgb_order=False, mono_threshold=0.10, tag_threshold=0.22, max_tags=5, with_phash=True,
                     shadow_threshold=0.15, highlight_threshold=0.10):
        """Full per-image pass for a same-shaped batch -> list of result dicts (see results_from_host)."""
        t = ops.to_device_u8(images)
        n, h, w, _ = t.shape
        dev = self.score_images_device(t, rgb_order=rgb_order, with_phash=with_phash)
        hashes = (["%016x" % int(v) for v in dev["phash"].cpu().numpy().view(np.uint64)]   # batch_processor.py:216
                  if with_phash else [None] * n)
        hist = dev["hist256"].cpu().numpy().view(np.uint32).astype(np.int64)
        sums = dev["sums"].cpu().numpy()
        der = dev["derived"].cpu().numpy()
        raw = dev["aesthetic_raw"].cpu().numpy()
        emb = dev["embedding"].cpu().numpy()
        sims = dev["tag_sims"].cpu().numpy() if dev["tag_sims"] is not None else None
        return self.results_from_host(h, w, hist, sums, der, raw, emb, sims, hashes, mono_threshold=mono_threshold,
                                      tag_threshold=tag_threshold, max_tags=max_tags, shadow_threshold=shadow_threshold,
                                      highlight_threshold=highlight_threshold)

    # -- scorer.py:952-1147 ----------------------------------------------------------------------------------
    def score_photo_from_pil(self, pil_img, img_cv, original_path, cache=None):
        """The reference's single-image engine (`score_photo_from_pil`, used by `process_single_photo`, scorer.py:1989): one
        frame already in memory -> the complete database row.  Differences from the batch path it keeps: the clipping thresholds
        come from the config, the subject search runs when no face box exists (`get_placement_data(..., img_cv)`), and
        `is_monochrome` / `contrast_score` are among the aggregate's inputs.  `pil_img` is not needed (the hash, the CLIP input and
        the thumbnail all come from `img_cv` on the device); `cache` is accepted for signature compatibility.  Faces / EXIF come
        from `self.face_analyzer` / `self.get_exif_data` when the scorer has them, else the reference's "nothing found" values.
        Returns None (after printing the error) when scoring fails, like the reference."""
        from pathlib import Path
        from ..analyzers.composition import CompositionAnalyzer
        from ..utils.detection import detect_silhouette
        from ..utils.tags import get_tag_params
        try:
            cfg = self.config
            if cfg is None:
                raise ValueError("score_photo_from_pil needs a ScoringConfig (Facet(config=...))")
            frame = np.asarray(img_cv)
            img_h, img_w = frame.shape[:2]
            exposure = cfg.get_exposure_settings()
            mono = cfg.get_monochrome_settings()
            thr, max_tags = get_tag_params(cfg)
            res = self.score_images(frame[None], mono_threshold=mono.get("saturation_threshold_percent", 10) / 100,
                                    tag_threshold=thr, max_tags=max_tags,
                                    shadow_threshold=exposure.get("shadow_clip_threshold_percent", 15) / 100,
                                    highlight_threshold=exposure.get("highlight_clip_threshold_percent", 10) / 100)[0]
            face_analyzer = getattr(self, "face_analyzer", None)
            if face_analyzer is not None:
                face_res = face_analyzer.analyze_faces(frame)
            else:
                from .batch_processor import NO_FACES
                face_res = dict(NO_FACES)
            face_ratio = face_res.get("face_area", 0) / (img_h * img_w)
            comp = CompositionAnalyzer.get_placement_data(face_res.get("bbox"), img_w, img_h, cfg, frame)
            lines = CompositionAnalyzer.detect_leading_lines(frame)
            isolation_bonus, is_blink = 1.0, 0
            if face_res["face_count"] > 0:
                isolation_bonus = max(1.0, face_res["face_sharpness"] / (res["raw_sharpness_variance"] + 1))
                is_blink = face_res.get("is_blink", 0)
            get_exif = getattr(self, "get_exif_data", None)
            exif = (get_exif(original_path) if get_exif is not None else None) or {}
            is_silhouette = detect_silhouette({"is_silhouette": res["is_silhouette"]}, res["tags"], face_res.get("face_count", 0))
            aggregate, category = self.calculate_aggregate_logic({
                "aesthetic": res["aesthetic_unrounded"], "face_count": face_res["face_count"], "face_quality": face_res["face_quality"],
                "eye_sharpness": face_res["eye_sharpness"], "tech_sharpness": res["tech_sharpness_unrounded"],
                "color_score": res["color_score_unrounded"], "exposure_score": res["exposure_score_unrounded"],
                "face_ratio": face_ratio, "comp_score": comp["score"], "isolation_bonus": isolation_bonus, "is_blink": is_blink,
                "shadow_clipped": res["shadow_clipped"], "highlight_clipped": res["highlight_clipped"], "is_silhouette": is_silhouette,
                "histogram_spread": res["histogram_spread"], "is_monochrome": res["is_monochrome"],
                "contrast_score": res["contrast_score"], "iso": exif.get("iso"), "f_stop": exif.get("f_stop"),
            })
            for k in ("aesthetic_unrounded", "tech_sharpness_unrounded", "color_score_unrounded", "exposure_score_unrounded"):
                res.pop(k)
            resolved = Path(original_path).resolve()
            res.update({
                "path": str(resolved), "filename": Path(original_path).name, "category": category,
                "face_count": face_res["face_count"], "face_quality": face_res["face_quality"],
                "eye_sharpness": face_res["eye_sharpness"], "face_sharpness": face_res["face_sharpness"], "face_ratio": face_ratio,
                "comp_score": round(comp["score"], 2), "isolation_bonus": round(isolation_bonus, 2), "is_blink": is_blink,
                "aggregate": round(aggregate, 2), "power_point_score": float(comp["power_point_score"]),
                "raw_eye_sharpness": float(face_res.get("raw_eye_sharpness", 0)), "config_version": getattr(cfg, "version_hash", None),
                "is_silhouette": is_silhouette, "is_group_portrait": face_res.get("is_group_portrait", 0),
                "leading_lines_score": lines.get("leading_lines_score", 0), "face_confidence": face_res.get("max_face_confidence", 0),
                "topiq_score": None, "composition_explanation": comp.get("vlm_explanation"), "composition_pattern": None,
                "face_details": face_res.get("face_details", []),
            })
            res.update(exif)
            return res
        except Exception as e:                  # scorer.py:1144-1146
            print(f"Error scoring {original_path}: {e}")
            return None
