#!/usr/bin/env python3
"""ms per fused drg_movegen step (1 Mi random boards, L2 flushed between steps) for the three shapes of the call --
records only, + mask, + mask + tensor -- for sweeping the library's environment knobs (one process per setting):

    DRG_CHAIN_BLOCKS=4 python tools/fused_step_time.py [variant] [shapes, e.g. r,m,mt] [boards]
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))
import numpy as np
import torch
import draughts_b200 as d

variant = sys.argv[1] if len(sys.argv) > 1 else "standard"
shapes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["m"]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
b = d.random_positions(variant, n, seed=0)
S = b.S
db = d.DeviceBoardBatch.from_host(b, "cuda")
total = int(db.counts().sum().item())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
knobs = {k: v for k, v in os.environ.items() if k.startswith("DRG_")}
for shape in shapes:
    out = {"counts": torch.empty(n, dtype=torch.int32, device="cuda"), "offsets": torch.empty(n + 1, dtype=torch.int64, device="cuda"),
           "moves": torch.empty((total, 32), dtype=torch.uint8, device="cuda"), "overflow": torch.zeros(1, dtype=torch.int64, device="cuda")}
    if "m" in shape:
        out["mask"] = torch.empty((n, S * S), dtype=torch.bool, device="cuda")
    if "t" in shape:
        out["tensor"] = torch.empty((n, 4, S), dtype=torch.float32, device="cuda")
    for _ in range(5):
        flush.zero_(); db.movegen(out)
    ts = []
    for _ in range(30):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); db.movegen(out); z.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(z))
    print(f"{variant} {shape:>2}: {np.median(ts):.4f} ms median, {min(ts):.4f} min  {knobs}")
