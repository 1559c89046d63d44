#!/usr/bin/env python3
"""BASELINE configs[3]: random self-play to game_over with per-ply tensor + legal-move-mask export for RL.

    python tools/selfplay_bench.py --games 1048576 [--variant standard] [--no-export]
    torchrun --nproc-per-node 8 tools/selfplay_bench.py --games 1048576      # games sharded by id range

Prints one JSON line: ply-steps/s (each ply-step = one position exported + one move played), games/s, and the
export bandwidth against the measured write-only ceiling.  Results are keyed by global game id (counter RNG), so
they do not depend on the number of GPUs."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=1 << 20)
    ap.add_argument("--variant", default="standard")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-plies", type=int, default=1000)
    ap.add_argument("--no-export", action="store_true")
    ap.add_argument("--repeat", type=int, default=3, help="timed repetitions; the fastest is reported")
    ap.add_argument("--check-every", type=int, default=4, help="plies between host checks (one 16-byte read)")
    ap.add_argument("--compact-frac", type=float, default=0.97, help="drop finished games when fewer than this share runs")
    ap.add_argument("--compact-min", type=int, default=None, help="no compaction below this many working games")
    args = ap.parse_args()
    import numpy as np
    import torch
    import draughts_b200 as d
    from draughts_b200 import drivers
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = args.games * rank // world, args.games * (rank + 1) // world
    n = hi - lo

    def fresh():
        return d.DeviceBoardBatch.from_host(d.BoardBatch.start(args.variant, n), torch.device("cuda", local))

    fresh().rollout(args.seed, game_id0=lo, max_plies=8)  # warm-up: context, kernels
    if not args.no_export:  # ... and the ring buffers (torch's caching allocator keeps them for the timed run)
        fresh().selfplay_export(args.seed, game_id0=lo, max_plies=8, draw_rules=True, ring=2)
    torch.cuda.synchronize()
    best = None
    for _ in range(max(1, args.repeat)):
        db = fresh()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.no_export:
            r = db.rollout(args.seed, game_id0=lo, max_plies=args.max_plies, draw_rules=True)
            torch.cuda.synchronize()
            steps = int(r["counters"][0].item())
            exported = 0
            rounds = None
        else:
            r = db.selfplay_export(args.seed, game_id0=lo, max_plies=args.max_plies, draw_rules=True, ring=2,
                                   check_every=args.check_every, compact_frac=args.compact_frac,
                                   **({} if args.compact_min is None else {"compact_min": args.compact_min}))
            torch.cuda.synchronize()
            steps, exported, rounds = r["ply_steps"], r["exported_bytes"], r["rounds"]
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    dt = best
    res = r["result"].cpu().numpy()
    ctr = np.zeros(8, np.uint64)
    ctr[0] = steps
    for code, slot in ((1, 2), (2, 3), (3, 4), (0, 5)):
        ctr[slot] = int((res == code).sum())
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ctr = drivers.allreduce_counters(ctr)
    dt = float(tt.item())
    if rank == 0:
        print(json.dumps({"variant": args.variant, "games": args.games, "gpus": world, "export": not args.no_export,
                          "seconds": dt, "ply_steps": int(ctr[0]), "ply_steps_per_s": int(ctr[0]) / dt,
                          "games_per_s": args.games / dt, "rounds": rounds,
                          "exported_GB_per_gpu": exported / 1e9, "export_GBps_per_gpu": exported / dt / 1e9,
                          "white_wins": int(ctr[2]), "black_wins": int(ctr[3]), "draws": int(ctr[4]),
                          "unfinished": int(ctr[5])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
