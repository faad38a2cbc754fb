import sys, torch
sys.path.insert(0, ".")
from facet_b200 import ops
b = int(sys.argv[1]) if len(sys.argv) > 1 else 128
qkv = (torch.randn(b * 257, 3072, device="cuda") * 1.5).to(torch.bfloat16)
for legacy in (False, True):
    for _ in range(3):
        ops.vit_attention(qkv, b, legacy_mma=legacy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.vit_attention(qkv, b, legacy_mma=legacy)
    e1.record()
    torch.cuda.synchronize()
    print("legacy" if legacy else "tcgen05", e0.elapsed_time(e1) / 10 * 1e3, "us")
