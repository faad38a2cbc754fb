"""A/B of library variants on the attention kernel alone (batch 128 and 512, fp16), same box, back to back."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from facet_b200 import ops
for bsz in (128, 512):
    qkv = (torch.randn(bsz * 257, 3072, device="cuda") * 2).to(torch.float16)
    for _ in range(3):
        ops.vit_attention(qkv, bsz)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.vit_attention(qkv, bsz)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"  batch {bsz}: {best*1e3:.1f} us per layer, {bsz*16*(257*257*64*4)/best/1e9:.0f} TFLOP/s", flush=True)
'''
for name in sys.argv[1:]:
    env = dict(os.environ)
    if name != "main":
        env["FACET_B200_LIB"] = os.path.abspath(f"facet_b200/variants/lib_{name}.so")
    print(name, flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
