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
# the last window between two env_round launches that holds a whole round (bench.py ends with env-only and fill launches)
pairs = [(idx[i], idx[i + 1]) for i in range(len(idx) - 1) if any("conv2_attn" in k or "edge_bf16" in k for k, _ in L[idx[i] + 1:idx[i + 1]])]
a, b = pairs[-1] if pairs else (idx[-2], idx[-1])
step = [(k, v) for k, v in L[a + 1:b + 1] if "FillFunctor" not in k]     # bench.py's L2 flush (256 MiB fill) sits outside its timed events
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
