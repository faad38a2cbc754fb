"""Device timing of the thumbnail path on 24 MP frames."""
import sys, torch
sys.path.insert(0, ".")
from facet_b200 import ops
fr = torch.randint(0, 256, (32, 4000, 6000, 3), dtype=torch.uint8, device="cuda")
for _ in range(2):
    ops.thumbnails(fr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.thumbnails(fr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("thumbnails: %.3f ms per 32 frames = %.1f us per frame" % (ms, ms * 1e3 / 32))
