"""A/B of library variants on the fc GEMM (erf-GELU epilogue) at batch 128 and 512, same box, back to back."""
import os, subprocess, sys
CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from facet_b200 import ops
for m in (32896, 131584):
    k, n = 1024, 4096
    a = (torch.randn(m, k, device="cuda")).to(torch.float16)
    b = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.float16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.float16)
    for mode, name in ((ops.GEMM_BIAS_BF16, "bias"), (ops.GEMM_BIAS_GELU_BF16, "gelu")):
        for _ in range(3):
            ops.gemm_bf16(a, b, mode, bias=bias, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.gemm_bf16(a, b, mode, bias=bias, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(f"  M={m} {name}: {best*1e3:.1f} us, {2*m*n*k/best/1e9:.0f} TFLOP/s", flush=True)
'''
for name in sys.argv[1:]:
    env = dict(os.environ)
    if name != "main":
        env["FACET_B200_LIB"] = os.path.abspath(f"facet_b200/variants/lib_{name}.so")
    print(name, flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
