"""Time the technical kernel of several library variants (facet_b200/variants/lib_<name>.so), one subprocess each.
Prints GB/s per frame kind; `check` also compares the fast kernel with the generic one (bit-exact variants only)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import json, sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from facet_b200 import ops
from time_tech import make_frames
check = sys.argv[1] == "1"
res = {}
for kind in ("noise", "photo"):
    n = 16
    fr = make_frames(kind, n)
    torch.cuda.synchronize()
    for _ in range(2):
        ops.tech_stats_raw(fr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(3):
            ops.tech_stats_raw(fr)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 3)
    res[kind] = round(n * 72e6 / (best * 1e-3) / 1e9, 1)
    luma = torch.empty(fr.shape[:3], dtype=torch.uint8, device=fr.device)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(3):
            ops.tech_stats_raw(fr, luma_out=luma)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 3)
    res[kind + "_luma"] = round(n * 72e6 / (best * 1e-3) / 1e9, 1)
    del luma
    if check:
        a = ops.tech_stats_raw(fr[:2])
        b = ops.tech_stats_raw(fr[:2], force_generic=True)
        res[kind + "_exact"] = bool(all(torch.equal(x, y) for x, y in zip(a[:3], b[:3])))
    del fr
print("RESULT " + json.dumps(res))
'''


def main():
    names = sys.argv[1:]
    out = {}
    for name in names:
        env = dict(os.environ)
        if name != "main":
            env["FACET_B200_LIB"] = os.path.abspath(f"facet_b200/variants/lib_{name}.so")
        check = "0" if name.startswith("ko_") else "1"
        p = subprocess.run([sys.executable, "-c", CHILD, check], env=env, capture_output=True, text=True, timeout=600)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        out[name] = json.loads(line[0][7:]) if line else {"error": (p.stderr or p.stdout)[-400:]}
        print(name, out[name], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/tech_variants.json", "w"), indent=1)


if __name__ == "__main__":
    main()
