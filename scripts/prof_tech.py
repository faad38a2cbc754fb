"""One technical-pass launch on 8 'photo' 24 MP frames (target of the ncu capture)."""
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from facet_b200 import ops  # noqa: E402
from time_tech import make_frames  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "photo"
with_luma = len(sys.argv) > 2 and sys.argv[2] == "luma"
fr = make_frames(kind, 8)
luma = torch.empty(fr.shape[:3], dtype=torch.uint8, device=fr.device) if with_luma else None
for _ in range(3):
    ops.tech_stats_raw(fr, luma_out=luma)
torch.cuda.synchronize()
print("ok")
