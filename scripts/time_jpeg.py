"""Device timing of the JPEG decoder on 24 MP frames (not the bench): per-stage CUDA-event times for a batch of 16."""
import io
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, ".")
from facet_b200 import _lib, ops  # noqa: E402
from facet_b200.synth import synth_frame_int, synth_image_bgr  # noqa: E402
from facet_b200.utils import jpeg as fj  # noqa: E402

H, W = 4000, 6000
res = {}
for name, gen in (("int", lambda i: synth_frame_int(i, H, W)), ("photo", lambda i: synth_image_bgr(2000 + i, H, W))):
    for ri in (8, 25, 375, 0):
        kw = {"quality": 90}
        if ri:
            kw["restart_marker_blocks"] = ri
        if ri == 375 and name == "photo":
            continue
        datas = []
        t0 = time.perf_counter()
        for i in range(2):
            buf = io.BytesIO()
            Image.fromarray(gen(i)[:, :, ::-1].copy()).save(buf, "JPEG", **kw)
            datas.append(np.frombuffer(buf.getvalue(), np.uint8))
        enc_s = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter()
        Image.open(io.BytesIO(datas[0].tobytes())).convert("RGB").load()
        pil_ms = (time.perf_counter() - t0) * 1e3
        n = 32 if ri in (0, 8) else 16
        streams = [datas[i % 2] for i in range(n)]
        infos = [fj.parse(s) for s in streams]
        slot = (max(len(s) for s in streams) + 255) & ~255
        buf = torch.empty(n * slot + 256, dtype=torch.uint8, device="cuda")
        for k, s in enumerate(streams):
            buf[k * slot:k * slot + s.size].copy_(torch.from_numpy(s))
        torch.cuda.synchronize()
        ops.jpeg_decode_device(buf, slot, infos)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frames, status = ops.jpeg_decode_device(buf, slot, infos)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[f"{name}_ri{ri}"] = {"n": n, "MB_per_image": round(len(datas[0]) / 1e6, 2), "ms": round(ms, 3), "images_per_s": round(n / (ms * 1e-3), 1),
                                 "pil_decode_ms": round(pil_ms, 1), "pil_encode_s": round(enc_s, 2), "status": status.cpu().tolist()[:2]}
        print(name, ri, res[f"{name}_ri{ri}"], flush=True)
        del buf, frames
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/time_jpeg.json", "w"), indent=1)
