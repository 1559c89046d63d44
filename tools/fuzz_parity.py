#!/usr/bin/env python3
"""Differential fuzz of the CUDA path against the CPU oracle, far beyond what the test-suite runs.

    python tools/fuzz_parity.py [--millions 2] [--seed 100] [--variants standard,frisian,...]

Per variant and round: (1) random-play positions (counter RNG, snapshots up to ply 160 so that king endgames are
reached) and (2) synthetic king endgames (1-3 kings of the side to move against 3-15 random pieces) and dense random
boards; for every batch the full result of `BoardBatch.legal_moves(want_mask=True, want_tensor=True)` -- counts,
offsets, every 32-byte record in canonical order, mask rows, tensor planes -- must equal the oracle's bit for bit, and
so must the same batch through the fused device call `DeviceBoardBatch.movegen` (the two-stream pipeline);
then every move is pushed on the device and the children compared with the oracle's; finally complete random games
(`drg_rollout`) are compared with the oracle's replay.  Prints one JSON line per variant; exit code 1 on any mismatch."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))

import numpy as np  # noqa: E402


def synthetic(variant, S, n, rng):
    """king endgames and dense boards; turn random"""
    full = (1 << S) - 1
    wm = np.zeros(n, np.uint64); wk = np.zeros(n, np.uint64); bm = np.zeros(n, np.uint64); bk = np.zeros(n, np.uint64)
    for i in range(n):
        if i % 4 == 3:   # dense random board
            occ = int(rng.integers(0, 1 << 62)) & int(rng.integers(0, 1 << 62)) & full
            side = int(rng.integers(0, 1 << 62)) & full
            king = int(rng.integers(0, 1 << 62)) & int(rng.integers(0, 1 << 62)) & full
            wm[i] = occ & side & ~king; wk[i] = occ & side & king
            bm[i] = occ & ~side & ~king; bk[i] = occ & ~side & king
            continue
        k = int(rng.integers(4, 17))
        sq = rng.choice(S, size=k, replace=False)
        nk = int(rng.integers(1, 4))
        a = b = c = d = 0
        for t, s in enumerate(sq):
            bit = 1 << int(s)
            if t < nk: b |= bit
            elif rng.random() < 0.75: c |= bit
            else: d |= bit
        wm[i], wk[i], bm[i], bk[i] = a, b, c, d
    turn = rng.integers(0, 2, n).astype(np.uint8)
    swap = turn == 1   # the kings belong to the side to move
    wm2, wk2, bm2, bk2 = wm.copy(), wk.copy(), bm.copy(), bk.copy()
    wm2[swap], bm2[swap] = bm[swap], wm[swap]
    wk2[swap], bk2[swap] = bk[swap], wk[swap]
    return wm2, wk2, bm2, bk2, turn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--millions", type=float, default=1.0, help="random-play positions per variant (millions)")
    ap.add_argument("--synthetic", type=int, default=200_000)
    ap.add_argument("--games", type=int, default=100_000)
    ap.add_argument("--seed", type=int, default=100)
    ap.add_argument("--variants", default="standard,american,frisian,russian,brazilian,antidraughts,breakthrough,frysk")
    args = ap.parse_args()
    import draughts_b200 as d
    from oracle import oracle as O
    th = O.host_threads()
    bad = 0
    for variant in args.variants.split(","):
        t0 = time.perf_counter()
        S = d.BoardBatch.start(variant, 1).S
        stats = {"variant": variant, "positions": 0, "moves": 0, "children": 0, "max_moves": 0, "mismatch": 0}
        rng = np.random.default_rng(args.seed)
        batches = []
        left = int(args.millions * 1e6)
        rnd = 0
        while left > 0:
            m = min(left, 500_000)
            batches.append(d.random_positions(variant, m, seed=args.seed + rnd, max_ply=(40, 80, 120, 160)[rnd % 4]))
            left -= m
            rnd += 1
        if args.synthetic:
            batches.append(d.BoardBatch(variant, *synthetic(variant, S, args.synthetic, rng)))
        for b in batches:
            got = b.legal_moves(want_mask=True, want_tensor=True)
            want = O.batch_emit(variant, b.wm, b.wk, b.bm, b.bk, b.turn, threads=th)
            ok = (np.array_equal(got["counts"], want["counts"]) and np.array_equal(got["offsets"], want["offsets"])
                  and got["moves"].tobytes() == want["moves"].tobytes() and np.array_equal(got["mask"], want["mask"])
                  and got["tensor"].tobytes() == want["tensor"].tobytes())
            if ok and len(b):
                # the same batch through the fused device call (mask-row kernel beside the count -> records chain)
                import torch
                dev = d.DeviceBoardBatch.from_host(b, "cuda")
                total = int(want["offsets"][-1])
                out = {"counts": torch.empty(len(b), dtype=torch.int32, device="cuda"),
                       "offsets": torch.empty(len(b) + 1, dtype=torch.int64, device="cuda"),
                       "moves": torch.empty((max(total, 1), 32), dtype=torch.uint8, device="cuda"),
                       "mask": torch.empty((len(b), S * S), dtype=torch.bool, device="cuda"),
                       "tensor": torch.empty((len(b), 4, S), dtype=torch.float32, device="cuda"),
                       "overflow": torch.zeros(1, dtype=torch.int64, device="cuda")}
                dev.movegen(out)
                ok = (np.array_equal(out["counts"].cpu().numpy().astype(np.uint32), want["counts"])
                      and np.array_equal(out["offsets"].cpu().numpy().astype(np.uint64), want["offsets"])
                      and out["moves"].cpu().numpy()[:total].tobytes() == want["moves"].tobytes()
                      and np.array_equal(out["mask"].cpu().numpy().view(np.uint8), want["mask"])
                      and out["tensor"].cpu().numpy().tobytes() == want["tensor"].tobytes()
                      and int(out["overflow"].item()) == 0)
                del out, dev
                torch.cuda.empty_cache()
            stats["positions"] += len(b)
            stats["moves"] += int(want["offsets"][-1])
            stats["max_moves"] = max(stats["max_moves"], int(want["counts"].max(initial=0)))
            if not ok:
                stats["mismatch"] += 1
                diff = np.flatnonzero(got["counts"] != want["counts"])
                i = int(diff[0]) if len(diff) else -1
                print(f"MISMATCH {variant}: first differing count at board {i}: "
                      f"{b.fens()[i] if i >= 0 else 'counts equal, records/mask/tensor differ'}", file=sys.stderr)
                continue
            # push every move (children in move-list order) on a sample that the oracle finishes quickly
            m = min(len(b), 100_000)
            sub = d.BoardBatch(variant, b.wm[:m], b.wk[:m], b.bm[:m], b.bk[:m], b.turn[:m])
            sub.clock[:] = (np.arange(m) % 50).astype(np.uint8)
            o = O.batch_emit(variant, sub.wm, sub.wk, sub.bm, sub.bk, sub.turn, want_mask=False, want_tensor=False, threads=th)
            par = np.repeat(np.arange(m), o["counts"].astype(np.int64))
            kids = d.BoardBatch(variant, sub.wm[par], sub.wk[par], sub.bm[par], sub.bk[par], sub.turn[par], sub.clock[par])
            if len(kids):
                st = kids.push(o["moves"])
                ref = O.batch_apply(variant, sub.wm[par], sub.wk[par], sub.bm[par], sub.bk[par], sub.turn[par], sub.clock[par], o["moves"])
                same = np.array_equal(st, ref[6]) and all(np.array_equal(getattr(kids, k), r) for k, r in
                                                          zip(("wm", "wk", "bm", "bk", "turn", "clock"), ref))
                stats["children"] += len(kids)
                if not same:
                    stats["mismatch"] += 1
                    print(f"MISMATCH {variant}: push", file=sys.stderr)
        if args.games:
            g = d.BoardBatch.start(variant, args.games)
            res = g.rollout(seed=args.seed + 7)
            w = O.rollout(variant, args.games, args.seed + 7, threads=th)
            same = (np.array_equal(res["result"], w["result"]) and np.array_equal(res["plies"], w["plies"])
                    and all(np.array_equal(getattr(g, k), w[k]) for k in ("wm", "wk", "bm", "bk", "turn", "clock")))
            stats["games"] = args.games
            stats["plies"] = int(w["plies"].sum())
            if not same:
                stats["mismatch"] += 1
                print(f"MISMATCH {variant}: rollout", file=sys.stderr)
        stats["seconds"] = round(time.perf_counter() - t0, 1)
        bad += stats["mismatch"]
        print(json.dumps(stats), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
