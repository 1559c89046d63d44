#!/usr/bin/env python3
"""ms per fused drg_movegen step (1 Mi random Standard boards, mask + records, L2 flushed between steps) -- the bench
step alone, for sweeping the library's environment knobs (one process per setting):

    DRG_DEEP_BLOCKS_BESIDE=2 DRG_CHAIN_REC_BLOCKS=8 DRG_CHAIN_BLOCKS=6 python tools/fused_step_time.py
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

n, S = 1 << 20, 50
b = d.random_positions("standard", n, seed=0)
db = d.DeviceBoardBatch.from_host(b, "cuda")
total = int(db.counts().sum().item())
out = {"counts": torch.empty(n, dtype=torch.int32, device="cuda"), "offsets": torch.empty(n + 1, dtype=torch.int64, device="cuda"),
       "moves": torch.empty((total, 32), dtype=torch.uint8, device="cuda"), "mask": torch.empty((n, S * S), dtype=torch.bool, device="cuda"),
       "overflow": torch.zeros(1, dtype=torch.int64, device="cuda")}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    flush.zero_(); db.movegen(out)
ts = []
for _ in range(30):
    flush.zero_()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); db.movegen(out); z.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(z))
knobs = {k: v for k, v in os.environ.items() if k.startswith("DRG_")}
print(f"{np.median(ts):.4f} ms median, {min(ts):.4f} min  {knobs}")
