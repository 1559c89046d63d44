#!/usr/bin/env python3
"""BASELINE configs[4]: capture-heavy king-endgame batches (Frisian, Russian, Frysk!, ...) -- throughput and parity.

    python tools/king_endgames.py [--n 150000] [--variants frisian,russian,frysk] [--no-check] [--cpu]

Input: tools/fuzz_parity.py's synthetic generator (1-3 kings of the side to move against 3-15 random pieces, every
fourth board a dense random board): the batches on which one board can have a capture tree of 500 000 nodes and
43 000 legal captures.  Per variant one JSON line: drg_movegen_count and the fused drg_movegen (records + mask) timed
on the device with CUDA events, the largest move list, and -- unless --no-check -- whether counts, offsets and every
record equal the CPU oracle's (bit-exact).  --cpu also times the oracle port on all host threads on the same batch."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200", ROOT / "tools"):
    sys.path.insert(0, str(p))

import numpy as np  # noqa: E402


def run(variant, n, seed, check, cpu, reps=5):
    import torch
    import draughts_b200 as d
    from fuzz_parity import synthetic
    S = d.BoardBatch.start(variant, 1).S
    cols = synthetic(variant, S, n, np.random.default_rng(seed))
    b = d.BoardBatch(variant, *cols)
    dev = d.DeviceBoardBatch.from_host(b, "cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    counts = dev.counts()
    torch.cuda.synchronize()
    t_count = []
    for _ in range(reps):
        a, z = ev(), ev()
        a.record()
        counts = dev.counts()
        z.record()
        torch.cuda.synchronize()
        t_count.append(a.elapsed_time(z))
    total = int(counts.sum().item())
    out = {"counts": torch.empty(n, dtype=torch.int32, device="cuda"),
           "offsets": torch.empty(n + 1, dtype=torch.int64, device="cuda"),
           "moves": torch.empty((max(total, 1), 32), dtype=torch.uint8, device="cuda"),
           "mask": torch.empty((n, S * S), dtype=torch.bool, device="cuda"),
           "overflow": torch.zeros(1, dtype=torch.int64, device="cuda")}
    dev.movegen(out)
    torch.cuda.synchronize()
    t_full = []
    for _ in range(reps):
        a, z = ev(), ev()
        a.record()
        dev.movegen(out)
        z.record()
        torch.cuda.synchronize()
        t_full.append(a.elapsed_time(z))
    line = {"variant": variant, "boards": n, "total_moves": total, "max_moves": int(counts.max().item()),
            "count_ms": round(min(t_count), 3), "movegen_mask_ms": round(min(t_full), 3),
            "count_positions_per_s": n / (min(t_count) / 1e3), "movegen_positions_per_s": n / (min(t_full) / 1e3),
            "overflow": int(out["overflow"].item())}
    if check or cpu:
        from oracle import oracle as O
        th = O.host_threads()
        t0 = time.perf_counter()
        want = O.batch_emit(variant, b.wm, b.wk, b.bm, b.bk, b.turn, want_mask=True, want_tensor=False, threads=th)
        dt = time.perf_counter() - t0
        if cpu:
            line["cpu_port"] = {"positions_per_s": n / dt, "cores": th, "seconds": round(dt, 2), "what": "count pass + emit pass with mask"}
        if check:
            line["parity"] = bool(
                np.array_equal(out["counts"].cpu().numpy().astype(np.uint32), want["counts"])
                and np.array_equal(out["offsets"].cpu().numpy().astype(np.uint64), want["offsets"])
                and out["moves"].cpu().numpy()[:total].tobytes() == want["moves"].tobytes()
                and np.array_equal(out["mask"].cpu().numpy().view(np.uint8), want["mask"]))
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=150_000)
    ap.add_argument("--seed", type=int, default=100)
    ap.add_argument("--variants", default="frisian,russian,frysk")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    bad = 0
    for v in args.variants.split(","):
        line = run(v, args.n, args.seed, not args.no_check, args.cpu)
        bad += line.get("parity") is False
        print(json.dumps(line), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
