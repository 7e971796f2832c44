"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_summary.py file.csv [top]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = list(csv.reader(lines))
    hdr = r[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for row in r[1:]:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v = v / 1000 if row[ui] == "ns" else (v * 1000 if row[ui] == "ms" else v)
        short = re.sub(r"\(.*", "", row[ki])
        short = re.sub(r"void |mp::|<unnamed>::|\(anonymous namespace\)::", "", short)[:72]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(t for _, t in agg.values())
    print(f"launches {n}, total {tot:.1f} us")
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
