"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the last full step
(delimited by env_round_kernel launches).  usage: python profiles/launch_summary.py launches.csv [--list]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
L = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
idx = [i for i, (k, _) in enumerate(L) if "env_round" in k]
a, b = idx[-2], idx[-1]
step = L[a + 1:b + 1]
if "--list" in sys.argv:
    for k, v in step:
        print(f"{v / 1e3:9.1f} us  {k[:90]}")
agg = collections.OrderedDict()
for k, v in step:
    k = k[:70]
    agg.setdefault(k, [0, 0])
    agg[k][0] += v
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
for k, (v, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{v / 1e6:8.3f} ms {n:3d} {100 * v / tot:5.1f}%  {k}")
print(f"{tot / 1e6:8.3f} ms total ({len(step)} launches)")
