#!/usr/bin/env python3
"""descents-per-board statistics of the device DFS, replayed on the host (tools/devsim)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
from oracle import oracle as O
from fuzz_parity import synthetic

L = C.CDLL(str(Path(__file__).resolve().parent / "libdevsim.so"))
FAM = {"standard": 0, "american": 1, "frisian": 2, "russian": 3, "brazilian": 4, "antidraughts": 0, "breakthrough": 0, "frysk": 2}

def sim(variant, wm, wk, bm, bk, turn):
    n = len(wm)
    counts = np.zeros(n, np.uint32); desc = np.zeros(n, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L.sim_count(FAM[variant], C.c_int64(n), p(wm), p(wk), p(bm), p(bk), p(turn), p(counts), p(desc))
    return counts, desc

def report(name, counts, desc):
    d = desc[desc > 0]
    print(f"{name}: n={len(desc)} deep={len(d)} ({len(d)/len(desc):.3%}) total_desc={d.sum()} mean={d.mean():.2f} "
          f"p50={np.percentile(d,50):.0f} p90={np.percentile(d,90):.0f} p99={np.percentile(d,99):.0f} p999={np.percentile(d,99.9):.0f} max={d.max()} maxmoves={counts.max()}")
    for b in (8, 12, 16, 24, 32, 64, 256, 1024, 10000, 100000):
        over = d[d > b]
        print(f"   >{b}: boards={len(over)} descents_in_them={over.sum()} ({over.sum()/max(d.sum(),1):.1%})")

if __name__ == "__main__":
    which = sys.argv[1]
    variant = sys.argv[2]
    n = int(sys.argv[3])
    if which == "synthetic":
        rng = np.random.default_rng(100)
        wm, wk, bm, bk, turn = synthetic(variant, O.SQUARES[variant], n, rng)
    else:
        sys.path.insert(0, str(ROOT / "py-draughts_b200"))
        from draughts_b200.batch import splitmix64_np
        g = np.arange(0, n, dtype=np.uint64)
        stop = ((splitmix64_np(np.uint64(0 ^ 0xD1B54A32D192ED03) + g) >> np.uint64(32)) * np.uint64(81) >> np.uint64(32)).astype(np.uint16)
        r = O.rollout(variant, n, 0, game_id0=0, max_plies=80, stop_ply=stop, flags=0, threads=2)
        wm, wk, bm, bk, turn = r["wm"], r["wk"], r["bm"], r["bk"], r["turn"]
    t0 = time.time()
    counts, desc = sim(variant, wm, wk, bm, bk, turn)
    print("sim seconds", time.time() - t0)
    want = O.batch_count(variant, wm, wk, bm, bk, turn, threads=2)
    print("counts equal oracle:", np.array_equal(counts, want))
    report(f"{which}/{variant}", counts, desc)
