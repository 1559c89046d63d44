#!/usr/bin/env python3
"""perft driver CLI (the new caller SURVEY 8(f)-1 names; the reference has no perft).

    python tools/perft.py --variant standard --depth 11 [--fen "W:W31,...:B1,..."] [--divide]
    torchrun --nproc-per-node 8 tools/perft.py --variant american --depth 12     # root subtrees sharded, 1 all-reduce

Prints one JSON line per depth: nodes, seconds, nodes/s (wall clock around the synchronising C-ABI call)."""
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
    ap.add_argument("--variant", default="standard")
    ap.add_argument("--depth", type=int, default=9)
    ap.add_argument("--from-depth", type=int, default=1)
    ap.add_argument("--fen", default=None)
    ap.add_argument("--divide", action="store_true")
    args = ap.parse_args()
    import draughts_b200 as d
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    if args.divide:
        for mv, nodes in d.perft_divide(args.fen, args.depth, variant=args.variant).items():
            print(f"{mv}: {nodes}")
        return
    d.perft(args.fen, max(1, args.depth - 1), variant=args.variant)  # warm-up: context, (almost) all frontier buffers
    for depth in range(args.from_depth, args.depth + 1):
        t0 = time.perf_counter()
        nodes, ctr = d.perft(args.fen, depth, variant=args.variant, return_counters=True)
        dt = time.perf_counter() - t0
        if rank == 0:
            print(json.dumps({"variant": args.variant, "depth": depth, "nodes": nodes, "seconds": round(dt, 6),
                              "nodes_per_s": nodes / dt, "movegen_positions": int(ctr[1]), "gpus": world}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
