#!/usr/bin/env python3
"""Small fixed workloads for ncu (launch lists / --set full captures); prints CUDA-event times when run plain.

    python tools/profile_case.py king  [--variant frisian] [--n 150000]   # synthetic king endgames: count, fused movegen
    python tools/profile_case.py bench [--n 1048576]                      # the bench step: fused drg_movegen with mask
    python tools/profile_case.py perft [--variant standard] [--depth 10]
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200", ROOT / "tools"):
    sys.path.insert(0, str(p))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=["king", "bench", "perft"])
    ap.add_argument("--variant", default=None)
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import draughts_b200 as d
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn):
        a, z = ev(), ev()
        a.record()
        r = fn()
        z.record()
        torch.cuda.synchronize()
        return a.elapsed_time(z), r

    if args.case == "perft":
        v = args.variant or "standard"
        for _ in range(args.reps):
            import time
            t0 = time.perf_counter()
            nodes = d.perft(None, args.depth, variant=v)
            print(f"perft {v} d{args.depth} = {nodes} in {time.perf_counter() - t0:.4f} s")
        return
    if args.case == "king":
        from fuzz_parity import synthetic
        v = args.variant or "frisian"
        n = args.n or 150_000
        S = d.BoardBatch.start(v, 1).S
        b = d.BoardBatch(v, *synthetic(v, S, n, np.random.default_rng(100)))
    else:
        v = args.variant or "standard"
        n = args.n or (1 << 20)
        b = d.random_positions(v, n, seed=0)
        S = b.S
    dev = d.DeviceBoardBatch.from_host(b, "cuda")
    for _ in range(args.reps):
        ms, counts = timed(dev.counts)
        print(f"{args.case} {v} n={n}: drg_movegen_count {ms:.3f} ms")
    total = int(counts.sum().item())
    out = {"counts": torch.empty(n, dtype=torch.int32, device="cuda"),
           "offsets": torch.empty(n + 1, dtype=torch.int64, device="cuda"),
           "moves": torch.empty((max(total, 1), 32), dtype=torch.uint8, device="cuda"),
           "mask": torch.empty((n, S * S), dtype=torch.bool, device="cuda"),
           "overflow": torch.zeros(1, dtype=torch.int64, device="cuda")}
    for _ in range(args.reps + 1):
        ms, _ = timed(lambda: dev.movegen(out))
        print(f"{args.case} {v} n={n}: drg_movegen (records + mask) {ms:.3f} ms, total moves {total}")


if __name__ == "__main__":
    main()
