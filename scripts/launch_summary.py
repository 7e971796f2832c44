"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_summary.py file.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    full = r[ki]
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", full)
    name = ((m.group(1) + (m.group(2) or "")) if m else full)[:60].replace("mp::", "")
    t = float(r[vi].replace(",", ""))
    t = t / 1000 if r[ui] == "ns" else t * 1000 if r[ui] == "ms" else t
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
print(f"| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f} % |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | | |")
