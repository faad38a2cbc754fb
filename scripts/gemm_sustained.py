"""Sustained (2 s loops) TFLOP/s, board power and SM clock of this library's tcgen05 GEMM against torch.matmul (cuBLAS), bf16."""
import sys, threading, time
import pynvml, torch
sys.path.insert(0, ".")
from facet_b200 import ops
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)

def run(name, fn, flops, seconds=2.0):
    fn(); torch.cuda.synchronize()
    samples = []; stop = threading.Event()
    def poll():
        while not stop.is_set():
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.02)
    th = threading.Thread(target=poll); th.start()
    t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(16): fn()
        n += 16; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    late = [s for s in samples if s[0] - t0 > 0.6]
    w = sum(s[1] for s in late) / len(late); mhz = sorted(s[2] for s in late)[len(late) // 2]
    print(f"{name:44s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.0f} TFLOP/s  {w:6.0f} W  {mhz} MHz  {w*ms*1e-3/flops*1e12:.3f} pJ/flop", flush=True)

for (m, n, k) in [(32896, 3072, 1024), (32896, 1024, 4096), (8192, 8192, 8192)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    wt = w.t()
    fl = 2.0 * m * n * k
    run(f"facet_b200 gemm  {m}x{n}x{k}", lambda: ops.gemm_bf16(a, w, ops.GEMM_BIAS_BF16, bias=bias, out=out), fl)
    run(f"torch.matmul     {m}x{n}x{k}", lambda: torch.matmul(a, wt), fl)
    run(f"F.linear (+bias) {m}x{n}x{k}", lambda: torch.nn.functional.linear(a, w, bias.to(torch.bfloat16)), fl)
    del a, w, out
