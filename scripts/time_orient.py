"""Device timing of fb_orient on 24 MP frames, every EXIF orientation (not the bench)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from facet_b200 import ops  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
n = 16
g = torch.Generator(device="cuda").manual_seed(0)
fr = torch.randint(0, 256, (n, 4000, 6000, 3), dtype=torch.uint8, device="cuda", generator=g)
res = {}
for code in range(1, 9):
    for _ in range(2):
        ops.orient(fr, code, swap_rb=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.orient(fr, code, swap_rb=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[code] = {"us_per_frame": ms * 1e3 / n, "GB_s": n * 144e6 / (ms * 1e-3) / 1e9}
    print(code, res[code], flush=True)
json.dump(res, open("gpurun_out/time_orient.json", "w"))
