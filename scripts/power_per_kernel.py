"""Board power, SM clock and energy per launch of the step's kernels, each looped alone for ~1.5 s (NVML samples every 20 ms;
the first 0.4 s are discarded).  The step as a whole runs at the 1 000 W cap: this shows which kernels sit at the cap."""
import json
import sys
import threading
import time

import pynvml
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from facet_b200 import ops
from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
from time_tech import make_frames

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def measure(name, fn, seconds=1.5):
    fn(); torch.cuda.synchronize()
    samples = []
    stop = threading.Event()

    def poll():
        while not stop.is_set():
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.02)
    th = threading.Thread(target=poll); th.start()
    t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(8):
            fn()
        n += 8
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    late = [s for s in samples if s[0] - t0 > 0.4]
    w = sum(s[1] for s in late) / max(1, len(late))
    mhz = sorted(s[2] for s in late)[len(late) // 2] if late else 0
    out = {"ms_per_launch": round(ms, 4), "watts": round(w, 1), "sm_mhz": mhz, "joules_per_launch": round(w * ms * 1e-3, 4)}
    print(name, out, flush=True)
    return out


res = {}
fr = make_frames("photo", 64)
luma = torch.empty(fr.shape[:3], dtype=torch.uint8, device="cuda")
res["technical_64_frames"] = measure("technical (64 frames, + luma)", lambda: ops.tech_stats_raw(fr, luma_out=luma))
res["preprocess_64_frames"] = measure("clip_preprocess (64 frames)", lambda: ops.clip_preprocess(fr))
res["phash_64_frames"] = measure("phash (64 frames, luma ready)", lambda: ops.phash(fr, device_only=True, luma=luma))
del fr, luma
m, k = 32896, 1024
for name, n, kk, mode in (("gemm_qkv", 3072, 1024, ops.GEMM_BIAS_BF16), ("gemm_fc_gelu", 4096, 1024, ops.GEMM_BIAS_GELU_BF16)):
    a = torch.randn(m, kk, device="cuda").to(torch.float16)
    b = (torch.randn(n, kk, device="cuda") * kk ** -0.5).to(torch.float16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.float16)
    res[name] = measure(name, lambda: ops.gemm_bf16(a, b, mode, bias=bias, out=out))
    del a, b, out
qkv = (torch.randn(128 * 257, 3072, device="cuda") * 2).to(torch.float16)
res["attention_b128"] = measure("attention (batch 128, one layer)", lambda: ops.vit_attention(qkv, 128))
del qkv
model = ClipVitL14(random_state_dict(0))
x = torch.randn(128, 3, 224, 224, device="cuda")
res["vit_tower_b128"] = measure("ViT-L/14 tower (batch 128)", lambda: model.encode(x), seconds=2.5)
json.dump(res, open("gpurun_out/r2_power_per_kernel.json", "w"), indent=1)
