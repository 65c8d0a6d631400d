"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total/avg ns, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/xxx_launches.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("sfm::", "").replace("void ", "")
    t = float(r[14])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot / 1e6:.3f} ms total (ncu per-launch times: cold-cache, serialised — compare shares)")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {t / 1e3:12.1f} {t / 1e3 / n:10.2f} {100 * t / tot:6.1f}%")
