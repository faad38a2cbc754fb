"""Rule-based composition analyzer (mirrors analyzers/composition.py:12-289 of the reference).

Per frame the reference runs, on the CPU, gray conversion, a 5x5 Gaussian blur and Canny (leading lines,
:210-218), or the median of the gray plane and Canny (subject search, :30-36), and then OpenCV's sequential
geometry (`cv2.HoughLinesP`, `cv2.findContours` + moments).  Here the per-pixel work — gray, blur, Sobel,
non-maximum suppression, hysteresis — runs on the GPU (`csrc/canny.cu`, bit-exact with OpenCV) and only the
edge map goes to the host, where the same OpenCV calls as in the reference do the sequential part (probabilistic
Hough visits points in the order of OpenCV's own RNG and erases votes as it accepts lines; border following walks
one contour at a time: neither has a parallel formulation that reproduces its output).
The closed-form placement scores (:95-188) and `integrate_leading_lines` (:263-289) are plain host arithmetic.
"""
import math

import numpy as np

from .. import ops

_THIRDS = (1 / 3, 2 / 3)


def _median_from_hist(hist256) -> float:
    """np.median of the uint8 plane whose 256-bin histogram is hist256 (mean of the two middle order statistics)."""
    h = np.asarray(hist256, dtype=np.int64)
    n = int(h.sum())
    cum = np.cumsum(h)
    lo = int(np.searchsorted(cum, (n - 1) // 2 + 1))      # value of order statistic (n - 1) // 2
    hi = int(np.searchsorted(cum, n // 2 + 1))            # value of order statistic n // 2
    return (lo + hi) / 2.0


def _device_gray(img_cv, cache, want_hist=False):
    """(CUDA gray plane, hist256 or None) of the frame; the frame may be a numpy array or a CUDA tensor."""
    rgb = bool(getattr(cache, "_rgb", False)) if cache is not None else False
    src = img_cv if img_cv is not None else getattr(cache, "_img", None)
    if want_hist:
        gray, hist = ops.gray_plane(src, rgb_order=rgb, want_hist=True)
        return gray, hist.cpu().numpy().view(np.uint32)
    return ops.gray_plane(src, rgb_order=rgb), None


class CompositionAnalyzer:
    @staticmethod
    def edge_map(img_cv, low, high, blur, cache=None):
        """uint8 [H,W] numpy edge map: cv2.Canny(cv2.GaussianBlur(gray, (5, 5), 0) if blur else gray, low, high)."""
        gray, _ = _device_gray(img_cv, cache)
        return ops.canny_edges(gray, low, high, blur=blur).cpu().numpy()

    @staticmethod
    def detect_subject_region(img_cv):
        """composition.py:16-93: box [x1, y1, x2, y2] of the main subject, or None."""
        if img_cv is None:
            return None
        import cv2
        h, w = int(img_cv.shape[0]), int(img_cv.shape[1])
        gray, hist = _device_gray(img_cv, None, want_hist=True)
        median_val = _median_from_hist(hist)
        lower = int(max(0, 0.5 * median_val))
        upper = int(min(255, 1.5 * median_val))
        edges = ops.canny_edges(gray, lower, upper, blur=False).cpu().numpy()
        contours, _ = cv2.findContours(edges, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        min_area = (h * w) * 0.0001
        valid = [c for c in contours if cv2.contourArea(c) > min_area]
        if valid:
            thirds_x = [w / 3, 2 * w / 3]
            thirds_y = [h / 3, 2 * h / 3]
            best, best_score = None, 0
            for contour in valid:
                m = cv2.moments(contour)
                if m["m00"] == 0:
                    continue
                cx = m["m10"] / m["m00"]
                cy = m["m01"] / m["m00"]
                area_score = cv2.contourArea(contour) / (h * w)
                dist_x = min(abs(cx - t) for t in thirds_x) / w
                dist_y = min(abs(cy - t) for t in thirds_y) / h
                score = area_score * (1 + max(0, 1 - (dist_x + dist_y)))
                if score > best_score:
                    best_score, best = score, contour
            if best is not None:
                x, y, bw, bh = cv2.boundingRect(best)
                return [x, y, x + bw, y + bh]
        # second strategy of the reference: spectral-residual saliency, when this OpenCV build has the module
        try:
            saliency = cv2.saliency.StaticSaliencySpectralResidual_create()
            frame = img_cv if isinstance(img_cv, np.ndarray) else img_cv.cpu().numpy()
            success, saliency_map = saliency.computeSaliency(frame)
            if success:
                saliency_map = (saliency_map * 255).astype(np.uint8)
                _, thresh = cv2.threshold(saliency_map, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
                sal_contours, _ = cv2.findContours(thresh, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
                if sal_contours:
                    x, y, bw, bh = cv2.boundingRect(max(sal_contours, key=cv2.contourArea))
                    return [x, y, x + bw, y + bh]
        except (cv2.error, AttributeError):
            pass
        return None

    @staticmethod
    def get_placement_score(bbox, img_w, img_h, config=None):
        if bbox is None:
            return 5.0
        cx = (bbox[0] + bbox[2]) / 2 / img_w
        cy = (bbox[1] + bbox[3]) / 2 / img_h
        dx = min(abs(cx - t) for t in _THIRDS)
        dy = min(abs(cy - t) for t in _THIRDS)
        thirds_score = max(0, 10 - (dx + dy) * 20)
        center_score = max(0, 10 - (abs(cx - 0.5) * 20))
        return max(thirds_score, center_score)

    @staticmethod
    def get_placement_data(bbox, img_w, img_h, config=None, img_cv=None):
        if bbox is None and img_cv is not None:
            bbox = CompositionAnalyzer.detect_subject_region(img_cv)
        if bbox is None:        # no subject: assume a centred one
            return {"score": 7.0, "power_point_score": 5.0, "line_score": 5.0, "center_score": 7.0}
        power_weight, line_weight = 2.0, 1.0
        if config:
            comp = config.get_composition_weights()
            power_weight = comp.get("power_point_weight", 2.0)
            line_weight = comp.get("line_weight", 1.0)
        cx = (bbox[0] + bbox[2]) / 2 / img_w
        cy = (bbox[1] + bbox[3]) / 2 / img_h
        # nearest of the four thirds intersections, in the reference's enumeration order (x outer, y inner)
        min_power = min(math.sqrt((cx - px) ** 2 + (cy - py) ** 2) for px in _THIRDS for py in _THIRDS)
        power_point_score = max(0, 10 - min_power * 25)
        dx = min(abs(cx - t) for t in _THIRDS)
        dy = min(abs(cy - t) for t in _THIRDS)
        line_score = max(0, 10 - (dx + dy) * 15)
        center_score = max(0, 10 - (abs(cx - 0.5) + abs(cy - 0.5)) * 10)
        weighted = (power_point_score * power_weight + line_score * line_weight) / (power_weight + line_weight)
        return {"score": round(max(weighted, center_score), 2), "power_point_score": round(power_point_score, 2),
                "line_score": round(line_score, 2), "center_score": round(center_score, 2)}

    @staticmethod
    def score_lines(lines, h, w):
        """composition.py:227-261: score of the segments cv2.HoughLinesP returned."""
        if lines is None:
            return {"leading_lines_score": 0, "line_count": 0}
        total_score = 0
        valid_lines = 0
        diagonal = np.sqrt(h ** 2 + w ** 2)
        for line in lines:
            x1, y1, x2, y2 = line[0]
            length = np.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
            if x2 - x1 != 0:
                angle = abs(np.degrees(np.arctan((y2 - y1) / (x2 - x1))))
            else:
                angle = 90
            angle_bonus = 1.5 if 15 <= angle <= 75 else 1.0
            total_score += (length / diagonal) * 10 * angle_bonus
            valid_lines += 1
        leading_lines_score = min(10.0, total_score / max(1, valid_lines) * 2)
        return {"leading_lines_score": round(leading_lines_score, 2), "line_count": len(lines)}

    @staticmethod
    def lines_from_edges(edges, h, w):
        """composition.py:220-225: the probabilistic Hough transform on an edge map (OpenCV, host)."""
        import cv2
        return cv2.HoughLinesP(edges, 1, np.pi / 180, 80, minLineLength=int(min(h, w) * 0.15), maxLineGap=20)

    @staticmethod
    def detect_leading_lines(img_cv, cache=None):
        """composition.py:190-261.  img_cv: BGR uint8 [H,W,3] (numpy or CUDA tensor); cache: optional ImageCache."""
        if img_cv is None:
            return {"leading_lines_score": 0, "line_count": 0}
        h, w = int(img_cv.shape[0]), int(img_cv.shape[1])
        edges = CompositionAnalyzer.edge_map(img_cv, 50, 150, True, cache=cache)
        return CompositionAnalyzer.score_lines(CompositionAnalyzer.lines_from_edges(edges, h, w), h, w)

    @staticmethod
    def integrate_leading_lines(base_comp_score, leading_lines_score, has_faces):
        if has_faces:
            return base_comp_score
        return min(10.0, base_comp_score + min(2.0, leading_lines_score / 5.0))
