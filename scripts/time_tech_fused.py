"""Technical pass with its side products (16 x 24 MP frames per launch, CUDA events, best of 3 x 3 launches):
plain / + luma plane / + (4, 4) box reduction / both, beside the separate reduction pass and the two thumbnail routes."""
import json
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from facet_b200 import ops
from time_tech import make_frames


def best_ms(fn, reps=3, inner=3):
    torch.cuda.synchronize()
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


res = {}
n = 16
for kind in ("noise", "photo"):
    fr = make_frames(kind, n)
    h, w = fr.shape[1:3]
    luma = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    box = torch.empty((n, h // 4, w // 4, 3), dtype=torch.uint8, device="cuda")
    t = {"plain": best_ms(lambda: ops.tech_stats_raw(fr)),
         "luma": best_ms(lambda: ops.tech_stats_raw(fr, luma_out=luma)),
         "box": best_ms(lambda: ops.tech_stats_raw(fr, box_out=box)),
         "luma_box": best_ms(lambda: ops.tech_stats_raw(fr, luma_out=luma, box_out=box)),
         "thumbnails_separate": best_ms(lambda: ops.thumbnails(fr)),
         "thumbnails_from_box": best_ms(lambda: ops.thumbnails(fr, reduced=box))}
    res[kind] = {k: {"us_per_frame": round(v * 1e3 / n, 2), "GB_s": round(n * 72e6 / (v * 1e-3) / 1e9, 1)} for k, v in t.items()}
    del fr, luma, box
print(json.dumps(res, indent=1))
