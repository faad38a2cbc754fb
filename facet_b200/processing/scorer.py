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
                          max_tags=5):
        """Host half of the pass: the small per-image device results (as host arrays: hist [n,256] int64, sums [n,4] int64,
        derived [n,4] float64, aesthetic_raw [n], embedding [n,768] float32, tag_sims [n,T] or None, hashes = list of hex
        strings or Nones) -> result dicts with the reference's metric keys (processing/batch_processor.py:298-355, the
        analyzer-derived subset)."""
        n = len(raw)
        metrics = cf.all_metrics_batch(h, w, hist, sums[:, 0], sums[:, 1], sums[:, 2], der[:, 0], der[:, 1],
                                       mono_threshold=mono_threshold)      # the 7 analyzer dicts per image, one pass
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

    def score_images(self, images, rgb_order=False, mono_threshold=0.10, tag_threshold=0.22, max_tags=5, with_phash=True):
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
                                      tag_threshold=tag_threshold, max_tags=max_tags)
