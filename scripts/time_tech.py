"""Quick device timing of the technical pass on 24 MP frames of several kinds (not the bench)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from facet_b200 import ops, _lib  # noqa: E402


def make_frames(kind, n, h=4000, w=6000, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if kind == "noise":
        return torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    yy = torch.linspace(0, 1, h, device="cuda")[None, :, None, None]
    xx = torch.linspace(0, 1, w, device="cuda")[None, None, :, None]
    ph = torch.rand((n, 1, 1, 3), device="cuda", generator=g) * 6.28
    fr = 1 + 4 * torch.rand((n, 1, 1, 3), device="cuda", generator=g)
    base = 0.5 + 0.25 * torch.sin(fr * 6.28 * xx + ph) + 0.25 * torch.cos(fr * 4.1 * yy + ph)
    sigma = {"smooth": 1.0, "photo": 4.0, "flat": 0.0}[kind]
    if kind == "flat":
        base = torch.round(base * 4) / 4
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device="cuda")
    for i in range(n):
        noise = torch.randn((h, w, 3), device="cuda", generator=g) * sigma
        out[i] = (base[i] * 255 + noise).clamp(0, 255).to(torch.uint8)
    return out


def main():
    res = {}
    with_luma = "luma" in sys.argv[1:]          # also emit the pHash luma plane, as the full pass does
    for kind in ("noise", "photo", "smooth", "flat"):
        n = 16
        fr = make_frames(kind, n)
        luma = torch.empty(fr.shape[:3], dtype=torch.uint8, device=fr.device) if with_luma else None
        torch.cuda.synchronize()
        for _ in range(2):
            ops.tech_stats_raw(fr, luma_out=luma)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            ops.tech_stats_raw(fr, luma_out=luma)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = n * 72e6 / (ms * 1e-3) / 1e9
        res[kind] = {"ms_per_16_frames": ms, "GB_s": gbs, "img_s": n / (ms * 1e-3)}
        print(kind, res[kind], flush=True)
        del fr
    json.dump(res, open("gpurun_out/time_tech.json", "w"))


if __name__ == "__main__":
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    main()
