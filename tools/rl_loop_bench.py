#!/usr/bin/env python3
"""The self-play data-collection loop of examples/reinforcement_learning.py:95-137, batched on the device.

The reference plays one game at a time: to_tensor -> net -> logits[~mask] = -1e9 -> softmax(logits / temp) ->
multinomial -> index_to_move -> push, with a CPU round trip per ply.  Here every ply of N games is one fused
drg_movegen call (records + legal-move mask + tensor into device ring buffers), one forward pass of the policy net
on the exported tensors, drg_sample_action (masked softmax sample + index_to_move on the device) and
drg_selfplay_step; the host looks at a 16-byte status word every few plies.

    python tools/rl_loop_bench.py [--variant american|standard] [--games 65536] [--hidden 256] [--temp 0.5]

Prints one JSON line: plies/s, games/s, and the share of the time spent in the policy net."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="american")
    ap.add_argument("--games", type=int, default=1 << 16)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--temp", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-plies", type=int, default=300)
    args = ap.parse_args()
    import torch
    import torch.nn as nn
    import draughts_b200 as d

    S = d.BoardBatch.start(args.variant, 1).S
    torch.manual_seed(args.seed)
    # the reference example's policy net shape: flattened 4 x S planes -> hidden -> S*S logits (random weights here)
    net = nn.Sequential(nn.Flatten(), nn.Linear(4 * S, args.hidden), nn.ReLU(), nn.Linear(args.hidden, args.hidden), nn.ReLU(),
                        nn.Linear(args.hidden, S * S)).cuda().eval()
    net_ms = []

    @torch.no_grad()
    def policy_logits(o):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = net(o["tensor"]).float().contiguous()
        z.record()
        net_ms.append((a, z))
        return out

    def run():
        db = d.DeviceBoardBatch.from_host(d.BoardBatch.start(args.variant, args.games), "cuda")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = db.selfplay_export(args.seed, max_plies=args.max_plies, policy_logits=policy_logits, temperature=args.temp,
                               compact_min=4096)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, r

    run()  # warm-up (allocator, kernels, cuBLAS handles)
    net_ms.clear()
    dt, r = run()
    net_s = sum(a.elapsed_time(z) for a, z in net_ms) / 1e3
    res = r["result"]
    print(json.dumps({"variant": args.variant, "games": args.games, "policy_head": S * S, "hidden": args.hidden, "temperature": args.temp,
                      "seconds": dt, "plies": r["ply_steps"], "plies_per_s": r["ply_steps"] / dt, "games_per_s": args.games / dt,
                      "rounds": r["rounds"], "policy_net_seconds": net_s, "policy_net_share": net_s / dt,
                      "white_wins": int((res == 1).sum()), "black_wins": int((res == 2).sum()), "draws": int((res == 3).sum()),
                      "unfinished": int((res == 0).sum()),
                      "loop": "drg_movegen (records + mask + tensor) -> policy net -> drg_sample_action -> drg_selfplay_step, all on the device"}))


if __name__ == "__main__":
    main()
