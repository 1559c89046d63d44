#!/usr/bin/env python3
"""Aggregate an `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu_lines.py dump.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ci: continue
    line, src = r[0], r[1]
    key = (fname, line)
    a = agg.setdefault(key, [0, 0, 0, src])
    if src and not a[3]: a[3] = src
    try:
        a[0] += int(r[ci] or 0); a[1] += int(r[ti] or 0); a[2] += int(r[si] or 0)
    except ValueError:
        pass
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[2] for a in agg.values()) or 1
print(f"total warp-inst {tot}  samples {tots}  avg active threads {sum(a[1] for a in agg.values())/tot:.1f}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]/tot*100:5.1f}% inst {a[2]/tots*100:5.1f}% smp thr={a[1]/max(a[0],1):4.1f} {f}:{l:>4} {a[3].strip()[:100]}")
