"""Two ViT forward passes at a given batch (target of ncu captures)."""
import sys
import torch
sys.path.insert(0, ".")
from facet_b200.models.clip_vit import ClipVitL14, random_state_dict
b = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = ClipVitL14(random_state_dict(0))
x = torch.randn(b, 3, 224, 224, device="cuda")
for _ in range(2):
    model.encode(x)
torch.cuda.synchronize()
print("ok")
