"""One fb_orient launch per EXIF code on 8 x 24 MP frames (target of the ncu capture)."""
import sys

import torch

sys.path.insert(0, ".")
from facet_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
fr = torch.randint(0, 256, (8, 4000, 6000, 3), dtype=torch.uint8, device="cuda", generator=g)
for code in (1, 6, 3):
    ops.orient(fr, code, swap_rb=True)
torch.cuda.synchronize()
print("ok")
