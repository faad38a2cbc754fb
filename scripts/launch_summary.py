"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_summary.py <csv> [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = OrderedDict()
for r in rows[1 + skip:]:
    name = re.sub(r"<.*", "", r[ik].split("(")[0]).split("::")[-1].strip()
    v = float(r[iv].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}.get(r[iu], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:34s} n={n:4d} total={us / 1e3:8.3f} ms  avg={us / n:9.1f} us  {100 * us / tot:5.1f}%")
print(f"total {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
