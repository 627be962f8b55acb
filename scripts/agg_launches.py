"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) by kernel.

    python scripts/agg_launches.py gpurun_out/r02_launches_token200.csv [max name length]
"""
import collections
import csv
import sys

f = sys.argv[1]
nlen = int(sys.argv[2]) if len(sys.argv) > 2 else 70
rows = list(csv.reader(open(f)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.defaultdict(dict)
name = {}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
    name[r[ii]] = r[ki]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    k = name[i]
    k = k[:k.index("(")] if "(" in k and not k.startswith("void") else k
    k = k.replace("void ", "")[:nlen]
    a = agg[k]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
totb = sum(v[2] for v in agg.values())
print(f"{f}: {len(per)} launches, {tot / 1e3:.1f} us of kernel time (serialised, cold), {totb / 1e6:.1f} MB of DRAM traffic")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"  {k:{nlen}s} n={v[0]:4d}  {v[1] / 1e3:9.1f} us {100 * v[1] / tot:5.1f} %   {v[1] / v[0] / 1e3:7.2f} us each   "
          f"{v[2] / max(v[1], 1e-9):7.2f} GB/s... {v[2] / v[0] / 1e6:8.2f} MB each")
