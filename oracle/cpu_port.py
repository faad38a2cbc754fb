"""ORACLE / CPU baseline (test infrastructure): the reference's CPU path for one image, written
against the same library calls the reference makes, so it can be TIMED on the GPU box where
/root/reference does not exist (bench.py cpu_baseline, `--impl reference`; kind = "port").

Call order and library calls follow processing/batch_processor.py:198-233:
  ImageCache (analyzers/image_cache.py:30-32)  cv2.cvtColor x2, cv2.Laplacian(...).var()
  get_sharpness_data / get_color_harmony_data / get_histogram_data / detect_monochrome /
  get_dynamic_range / get_noise_estimate / get_contrast_score  (analyzers/technical.py:39-342)
  scorer.preprocess (torchvision Resize/CenterCrop/ToTensor/Normalize on a PIL image)
  encode_image + normalize + aesthetic head (processing/scorer.py:661-664), fp32 on CPU
Never imported by facet_b200.
"""
from __future__ import annotations

import struct

import numpy as np


def technical_metrics_cv(img_bgr: np.ndarray, mono_threshold: float = 0.10) -> dict:
    import cv2
    from scipy.stats import kurtosis
    gray = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2GRAY)
    hsv = cv2.cvtColor(img_bgr, cv2.COLOR_BGR2HSV)
    lap_var = cv2.Laplacian(gray, cv2.CV_64F).var()
    out = {"sharpness": {"raw_variance": lap_var, "normalized": float(min(10.0, lap_var / 50.0))}}
    # colour entropy
    hs = cv2.calcHist([hsv], [0, 1], None, [180, 256], [0, 180, 0, 256])
    tot = hs.sum()
    ent = 0
    if tot > 0:
        p = hs / tot
        nz = p > 0
        ent = -np.sum(p[nz] * np.log2(p[nz]))
    out["color"] = {"raw_entropy": ent, "normalized": float(min(10.0, ent * 10.0 / 15.5))}
    # luminance histogram block
    h = cv2.calcHist([gray], [0], None, [256], [0, 256]).flatten()
    t = h.sum()
    hn = h / t if t > 0 else h
    bins = np.arange(256)
    mu = np.sum(bins * hn)
    spread = np.sqrt(np.sum(((bins - mu) ** 2) * hn))
    lum = mu / 255.0
    shadow, highlight = np.sum(hn[:30]), np.sum(hn[225:])
    sil = 1 if (np.sum(hn[:85]) > 0.35 and np.sum(hn[170:]) > 0.25) else 0
    bim = -kurtosis(hn * 256, fisher=True)
    score = 7.0 - abs(lum - 0.5) * 8 + min(4.0, spread / 20.0) - max(0, bim - 1.0) * 0.6
    if not sil:
        score -= shadow * 4.0 + highlight * 5.0
    out["histogram"] = {"histogram_bytes": struct.pack("256f", *hn), "spread": round(spread, 4),
                        "mean_luminance": round(lum, 4), "bimodality": round(bim, 4),
                        "exposure_score": round(max(0, min(10.0, score)), 2),
                        "shadow_clipped": 1 if shadow > 0.15 else 0, "highlight_clipped": 1 if highlight > 0.10 else 0,
                        "is_silhouette": sil}
    ms = np.mean(hsv[:, :, 1]) / 255.0
    out["monochrome"] = {"is_monochrome": 1 if ms < mono_threshold else 0, "mean_saturation": round(ms, 4)}
    p2, p98 = np.percentile(gray, 2), np.percentile(gray, 98)
    p2 = 1 if p2 < 1 else p2
    out["dynamic_range"] = {"dynamic_range_stops": round(np.log2(max(p98, 1) / p2), 2)}
    g64 = gray.astype(np.float64)
    hh, ww = g64.shape
    M = np.array([[1, -2, 1], [-2, 4, -2], [1, -2, 1]])
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma = np.sum(np.abs(cv2.filter2D(g64, -1, M))) * np.sqrt(0.5 * np.pi) / (6 * (ww - 2) * (hh - 2))
    out["noise"] = {"noise_sigma": round(sigma, 2)}
    g64 = gray.astype(np.float64)
    p5, p95 = np.percentile(g64, [5, 95])
    pc = (p95 - p5) / 255.0
    rms = np.std(g64) / 255.0
    out["contrast"] = {"contrast_score": round(min(10.0, pc * 5.0 + rms * 20.0), 2),
                       "percentile_contrast": round(pc, 4), "rms_contrast": round(rms, 4)}
    return out


def technical_metrics_ref(img_bgr: np.ndarray, image_cache_cls, analyzer_cls, mono_threshold: float = 0.10) -> dict:
    """The same seven dicts through the reference's OWN classes (oracle/_ref, placed by oracle/build_ref.py), in the
    call order of processing/batch_processor.py:198-233."""
    cache = image_cache_cls(img_bgr)
    ta = analyzer_cls
    return {"sharpness": ta.get_sharpness_data(img_bgr, cache=cache),
            "color": ta.get_color_harmony_data(img_bgr, cache=cache),
            "histogram": ta.get_histogram_data(img_bgr, cache=cache),
            "monochrome": ta.detect_monochrome(img_bgr, threshold=mono_threshold, cache=cache),
            "dynamic_range": ta.get_dynamic_range(img_bgr, cache=cache),
            "noise": ta.get_noise_estimate(img_bgr, cache=cache),
            "contrast": ta.get_contrast_score(img_bgr, cache=cache)}


def clip_preprocess_pil(img_bgr: np.ndarray, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    import torchvision.transforms as T
    from PIL import Image
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), T.ToTensor(),
                    T.Normalize(mean, std)])
    return tf(Image.fromarray(np.ascontiguousarray(img_bgr[..., ::-1])))


def score_images_cpu(images_bgr, state_dict, tag_embeddings=None, ref_analyzers=None):
    """The reference's per-image pass on the CPU for a list of BGR frames (fp32 tower).  ref_analyzers =
    (ImageCache, TechnicalAnalyzer) of the reference (oracle.build_ref.load()) routes the technical metrics through the
    reference's own code instead of the port."""
    import torch
    from . import vit_torch
    from . import phash as _ph
    if ref_analyzers is not None:
        tech = [technical_metrics_ref(im, *ref_analyzers) for im in images_bgr]
    else:
        tech = [technical_metrics_cv(im) for im in images_bgr]
    for t, im in zip(tech, images_bgr):
        t["phash"] = _ph.phash_hex(im)                      # imagehash.phash, batch_processor.py:216
    clip_in = torch.stack([clip_preprocess_pil(im) for im in images_bgr])
    vit = vit_torch.score_batch(state_dict, clip_in, tag_embeddings)
    return tech, vit
