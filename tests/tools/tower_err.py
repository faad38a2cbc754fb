import sys, torch
sys.path.insert(0, ".")
from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
from oracle import vit_torch
sd = random_state_dict(0)
x = torch.randn(16, 3, 224, 224, generator=torch.Generator().manual_seed(3))
sd_gpu = {k: v.cuda() for k, v in sd.items()}
ref = vit_torch.score_batch(sd_gpu, x.cuda(), None)
model = ClipVitL14(sd, dtype=sys.argv[1] if len(sys.argv) > 1 else "fp16")
out = model.encode(x.cuda())
aest = ((out["aesthetic_raw"] + 1) * 5).clamp(0, 10)
err = (aest - ref["aesthetic"]).abs()
cos = torch.nn.functional.cosine_similarity(out["embedding"], ref["embedding"], dim=-1)
rel = (out["features"] - ref["features"]).norm(dim=-1) / ref["features"].norm(dim=-1)
print("aest err max %.4f mean %.4f | min cos %.6f | feature rel err mean %.5f" % (float(err.max()), float(err.mean()), float(cos.min()), float(rel.mean())))
