#!/usr/bin/env python3
"""Per-launch summary of an `ncu --page raw --csv` export: the counters the roofline arguments of DESIGN.md use.

    python tools/ncu_summary.py gpurun_out/x_raw.csv [--json out.json]
"""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "dur_us", 1e-3),
        ("launch__grid_size", "grid", 1),
        ("smsp__inst_executed.sum", "warp_inst", 1),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%act", 1),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%el", 1),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
        ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
        ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("launch__registers_per_thread", "regs", 1)]


def load(path):
    rows = list(csv.reader(open(path)))
    H, U = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) < len(H):
            continue
        rec = {"kernel": r[H.index("Kernel Name")].split("(")[0].replace("void ", "").replace("drg::", "")}
        for col, name, scale in COLS:
            if col in H:
                v = r[H.index(col)].replace(",", "")
                unit = U[H.index(col)]
                try:
                    x = float(v)
                    if name == "dur_us" and unit in ("us", "usecond"):
                        x *= 1e3
                    if name == "dur_us" and unit in ("ms", "msecond"):
                        x *= 1e6
                    if name.endswith("_MB") and unit.startswith("K"):
                        x *= 1e3
                    if name.endswith("_MB") and unit.startswith("M"):
                        x *= 1e6
                    if name.endswith("_MB") and unit.startswith("G"):
                        x *= 1e9
                    rec[name] = x * scale
                except ValueError:
                    rec[name] = v
        out.append(rec)
    return out


if __name__ == "__main__":
    recs = load(sys.argv[1])
    names = ["kernel"] + [c[1] for c in COLS]
    print(" ".join(f"{n:>12s}" if n != "kernel" else f"{n:48s}" for n in names))
    for r in recs:
        print(" ".join((f"{r.get(n, ''):>12.2f}" if isinstance(r.get(n), float) else f"{str(r.get(n, '')):>12s}") if n != "kernel"
                       else f"{r[n][:48]:48s}" for n in names))
    if "--json" in sys.argv:
        json.dump(recs, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
