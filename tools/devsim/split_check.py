#!/usr/bin/env python3
"""host replay of the split walk vs the oracle: counts + records, and per-round statistics"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools")); sys.path.insert(0, str(ROOT / "py-draughts_b200"))
from oracle import oracle as O
from fuzz_parity import synthetic

L = C.CDLL(str(Path(__file__).resolve().parent / "libdevsim.so"))
FAM = {"standard": 0, "american": 1, "frisian": 2, "russian": 3, "brazilian": 4, "antidraughts": 0, "breakthrough": 0, "frysk": 2}
p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)


def split(variant, arrs, budgets, pool_cap=1 << 24, emit=True):
    wm, wk, bm, bk, turn = arrs
    n = len(wm)
    counts = np.zeros(n, np.uint32)
    stats = np.zeros(98, np.uint64)
    b = np.asarray(budgets, np.int32)
    rc = L.sim_split(FAM[variant], C.c_int64(n), p(wm), p(wk), p(bm), p(bk), p(turn), p(b), len(b), C.c_int64(pool_cap),
                     p(counts), None, None, p(stats))
    assert rc >= 0, rc
    moves = None
    if emit:
        offsets = np.zeros(n + 1, np.uint64)
        np.cumsum(counts, out=offsets[1:])
        moves = np.zeros(int(offsets[-1]) + 1, O.MOVE_DTYPE)
        c2 = np.zeros(n, np.uint32)
        rc = L.sim_split(FAM[variant], C.c_int64(n), p(wm), p(wk), p(bm), p(bk), p(turn), p(b), len(b), C.c_int64(pool_cap),
                         p(c2), p(offsets), p(moves), p(stats))
        assert rc >= 0, rc
        assert np.array_equal(c2, counts)
        moves = moves[:-1]
    return counts, moves, stats


def positions(which, variant, n, seed=100):
    if which == "synthetic":
        return synthetic(variant, O.SQUARES[variant], n, np.random.default_rng(seed))
    from draughts_b200.batch import splitmix64_np
    g = np.arange(0, n, dtype=np.uint64)
    stop = ((splitmix64_np(np.uint64(seed ^ 0xD1B54A32D192ED03) + g) >> np.uint64(32)) * np.uint64(161) >> np.uint64(32)).astype(np.uint16)
    r = O.rollout(variant, n, seed, game_id0=0, max_plies=160, stop_ply=stop, flags=0, threads=4)
    return r["wm"], r["wk"], r["bm"], r["bk"], r["turn"]


if __name__ == "__main__":
    which, variant, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
    budgets = [int(x) for x in sys.argv[4].split(",")]
    cap = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 24
    if len(sys.argv) > 6:
        L.sim_set_explode(int(sys.argv[6]))
    arrs = positions(which, variant, n)
    t0 = time.time()
    counts, moves, stats = split(variant, arrs, budgets, cap, emit=True)
    print("sim seconds", round(time.time() - t0, 2))
    want = O.batch_emit(variant, *arrs, want_mask=False, want_tensor=False, threads=6)
    print("counts equal:", np.array_equal(counts, want["counts"]), " records equal:", moves.tobytes() == want["moves"].tobytes(),
          " total moves", int(want["offsets"][-1]))
    print("deep boards", int(stats[96]), "records", int(stats[97]))
    for r in range(len(budgets)):
        print(f"  round {r}: budget {budgets[r]} walks {int(stats[3*r])} descents {int(stats[3*r+1])} max {int(stats[3*r+2])}")
