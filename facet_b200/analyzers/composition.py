"""Rule-of-thirds placement of a subject box (analyzers/composition.py:95-188 of the reference).

Only the closed-form part the per-image pass needs to fill `comp_score` / `power_point_score`
(batch_processor.py:239-242 calls `get_placement_data(face_bbox, w, h, config)` without the frame).
The edge / saliency subject search (:16-93) and the Hough leading-lines score (:190-261) are
SURVEY.md §8(f) rank 3 and not part of this path.
"""
import math

_THIRDS = (1 / 3, 2 / 3)


class CompositionAnalyzer:
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
            raise NotImplementedError("subject search on the frame (composition.py:16-93) is outside the scoring pass")
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
