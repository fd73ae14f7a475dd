"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name.
usage: python scripts/summarize_ncu_launches.py launches.csv [title] > summary.md"""
import csv, sys, re
from collections import defaultdict
path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu].strip(), 1e-3)
    name = re.sub(r"\(.*", "", r[ik]).strip()
    tot[name] += v; cnt[name] += 1
allt = sum(tot.values())
print(f"# {title}\n")
print(f"{sum(cnt.values())} launches, {allt:.0f} us of kernel time in the capture (per-launch times are cold-cache and serialised under ncu;"
      " the SHARE per kernel is what carries over to the un-profiled step).\n")
print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {100 * v / allt:.1f}% | {v:.1f} | {cnt[k]} | {v / cnt[k]:.1f} | `{k}` |")
