"""Crop sharpness for the isolation bonus (analyzers/face.py:272-279, batch_processor.py:254-260).

Face detection itself (InsightFace) is outside this path; what the scoring pass needs from it is the
Laplacian variance of the gray crop of every detected box, which `fb_roi_laplacian` computes from
exact integer sums (reflect-101 borders of the CROP, like cv2.Laplacian on the cropped array).
"""
from __future__ import annotations

import numpy as np


def _clamp_box(shape, bbox):
    h, w = shape[:2]
    return max(0, int(bbox[0])), max(0, int(bbox[1])), min(w, int(bbox[2])), min(h, int(bbox[3]))


def crop_sharpness_batch(img_cv, bboxes, rgb_order: bool = False):
    """`FaceAnalyzer._get_crop_sharpness(img, bbox)` for every box [x1, y1, x2, y2] of one frame:
    Laplacian(gray(crop)).var(), 0 for an empty crop.  One kernel launch for all boxes."""
    from .. import ops
    shape = img_cv.shape[-3:-1] if img_cv.ndim == 4 else img_cv.shape[:2]
    boxes = [_clamp_box(shape, b) for b in bboxes]
    out = [0] * len(boxes)
    live = [k for k, (x1, y1, x2, y2) in enumerate(boxes) if x2 > x1 and y2 > y1]
    if live:
        sums = ops.roi_laplacian(img_cv, [boxes[k] for k in live], rgb_order=rgb_order)
        for k, (n, s1, s2) in zip(live, sums.tolist()):
            mean = s1 / n
            out[k] = s2 / n - mean * mean
    return out


def crop_sharpness(img_cv, bbox, rgb_order: bool = False):
    return crop_sharpness_batch(img_cv, [bbox], rgb_order)[0]


def isolation_bonus(face_sharpness, full_variance, face_count=1):
    """batch_processor.py:254-260: max(1, face_sharpness / (laplacian_variance + 1)), 1.0 without faces."""
    if face_count <= 0:
        return 1.0
    return max(1.0, face_sharpness / (full_variance + 1))


def mean_face_sharpness(img_cv, bboxes, rgb_order: bool = False):
    """analyzers/face.py:180,236: the per-face crop sharpness averaged over the detected faces (0 without faces)."""
    vals = crop_sharpness_batch(img_cv, bboxes, rgb_order)
    return float(np.mean(vals)) if vals else 0
