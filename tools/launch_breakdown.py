#!/usr/bin/env python3
"""Per-kernel breakdown of the LAST fused drg_movegen call in an ncu launch list
(`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv --log-file X.csv ...`).
usage: launch_breakdown.py X.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hi]
ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
data = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    d = data.setdefault(r[ii], {"k": r[ki].split("(")[0].replace("void ", "").replace("drg::", "")})
    d[r[mi]] = float(r[vi].replace(",", ""))
last = list(data.values())
idx = [i for i, d in enumerate(last) if "k_emit<" in d["k"] and d["k"].replace("(bool)", "").endswith("1, 0, 1, 0>")]
for d in last[idx[-1]:]:
    inst = d.get("smsp__inst_executed.sum", 0)
    print(f"{d['k'][:58]:58s} {d.get('gpu__time_duration.sum', 0) / 1000:9.1f} us  {inst / 1e6:8.2f} M warp inst  "
          f"{d.get('smsp__thread_inst_executed.sum', 0) / max(inst, 1):5.1f} thr/inst"
          + (f"  dram rd {d['dram__bytes_read.sum'] / 1e6:8.1f} MB wr {d['dram__bytes_write.sum'] / 1e6:8.1f} MB"
             if 'dram__bytes_read.sum' in d else ""))
