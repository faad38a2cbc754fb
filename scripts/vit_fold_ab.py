"""A/B of the LayerNorm fold: two towers (fold on / off) in one process, measurements interleaved (each 1.2 s of back-to-back passes)."""
import sys, time
import torch
sys.path.insert(0, ".")
from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
sd = random_state_dict(0)
models = {"fold": ClipVitL14(sd, fold_layernorm=True), "no fold": ClipVitL14(sd, fold_layernorm=False)}
for b in (128, 512):
    x = torch.randn(b, 3, 224, 224, device="cuda")
    res = {k: [] for k in models}
    for rep in range(4):
        for name, m in models.items():
            m.encode(x); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); n = 0
            e0.record()
            while time.perf_counter() - t0 < 1.2:
                m.encode(x); n += 1
                torch.cuda.synchronize()
            e1.record(); torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / n)
    for name, v in res.items():
        print(f"batch {b} {name:8s}: " + " ".join(f"{t:.2f}" for t in v) + f"  ms  (median {sorted(v)[len(v)//2]:.2f}, {b/sorted(v)[len(v)//2]*1e3:.0f} images/s)", flush=True)
