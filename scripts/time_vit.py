"""Device timing of the ViT-L/14 tower at several batch sizes (not the bench)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from facet_b200.models.clip_vit import ClipVitL14, random_state_dict  # noqa: E402
from facet_b200 import ops  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
model = ClipVitL14(random_state_dict(0))
res = {}
for b in [int(a) for a in sys.argv[1:]] or [64, 256, 512]:
    x = torch.randn(b, 3, 224, 224, device="cuda")
    for _ in range(2):
        model.encode(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        model.encode(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[b] = {"ms": ms, "img_s": b / ms * 1e3, "tflops": 162.0e9 * b / (ms * 1e-3) / 1e12}
    print(b, res[b], flush=True)
# single GEMM shapes
for (m, n, k) in [(131584, 3072, 1024), (131584, 1024, 1024), (131584, 4096, 1024), (131584, 1024, 4096)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    bias = torch.zeros(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm_bf16(a, w, ops.GEMM_BIAS_BF16, bias=bias, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.gemm_bf16(a, w, ops.GEMM_BIAS_BF16, bias=bias, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tf = 2.0 * m * n * k / (ms * 1e-3) / 1e12
    e0.record()
    for _ in range(5):
        torch.matmul(a, w.T, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 5
    res[f"gemm_{m}x{n}x{k}"] = {"ms": ms, "tflops": tf, "cublas_ms": ms2, "cublas_tflops": 2.0 * m * n * k / (ms2 * 1e-3) / 1e12}
    print((m, n, k), res[f"gemm_{m}x{n}x{k}"], flush=True)
    del a, w, out
json.dump(res, open("gpurun_out/time_vit.json", "w"))
