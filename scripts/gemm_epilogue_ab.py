"""Back-to-back timing of the bias and erf-GELU epilogues on the QKV / fc shapes of ViT-L/14 at batch 128."""
import sys, torch
sys.path.insert(0, ".")
from facet_b200 import ops
m, k = 32896, 1024
for n in (3072, 4096):
    a = (torch.randn(m, k, device="cuda")).to(torch.float16)
    b = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.float16)
    bias = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.float16)
    for mode, name in ((ops.GEMM_BIAS_BF16, "bias"), (ops.GEMM_BIAS_GELU_BF16, "gelu")):
        for _ in range(3):
            ops.gemm_bf16(a, b, mode, bias=bias, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm_bf16(a, b, mode, bias=bias, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"N={n} {name}: {ms*1e3:.1f} us, {2*m*n*k/ms/1e9:.0f} TFLOP/s")
