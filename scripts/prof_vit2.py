"""Three forward passes of a 2-layer tower at batch 128 (per-GEMM ncu comparisons)."""
import sys
import torch
sys.path.insert(0, ".")
from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
model = ClipVitL14(random_state_dict(0, layers=2))
x = torch.randn(128, 3, 224, 224, device="cuda")
for _ in range(3):
    model.encode(x)
torch.cuda.synchronize()
print("ok")
