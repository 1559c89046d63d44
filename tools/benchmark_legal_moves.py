#!/usr/bin/env python3
"""Legal-moves benchmark worker with the reference worker's command line and JSON.

    python tools/benchmark_legal_moves.py <positions.json> <warmup> <rounds> [iterations]
                                          [--variant standard] [--mode batch|board] [--replicate K]

Mirrors `tools/workers/benchmark_legal_moves.py` of the reference: same positional arguments, same FEN file
(`{"positions": [fen, ...]}`, e.g. the reference's tools/random_positions.json; the golden fixture
tests/golden/ref_fixtures.json, which holds those 73 FENs, is accepted too and is the default), same cleaning of
the legacy `W:W:`/`B:B:` prefix, one round = every position parsed from its FEN and its legal moves generated,
median over rounds per iteration, and the same output keys (`iteration_medians`, `median_ms`, `positions_count`,
`high_priority`, `times`).  Added keys say how the round ran:

  --mode batch (default)  one `BoardBatch.from_fens` (native FEN codec) + one `legal_moves` C-ABI call per round
                          (counts + packed move records on the host); `--replicate K` repeats the file K times
                          inside the round so the batch is large enough for a GPU to matter.
  --mode board            the reference's loop verbatim through the drop-in classes:
                          `board = get_board(variant, fen); list(board.legal_moves)` -- one batch-of-1 launch per
                          position (latency, not throughput).
"""
import argparse
import gc
import json
import sys
import time
from pathlib import Path
from statistics import median

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))


def load_positions(path: Path) -> list[str]:
    doc = json.loads(path.read_text())
    if "positions" in doc:
        fens = list(doc["positions"])
    else:  # tests/golden/ref_fixtures.json
        fens = [r["src_fen"] for r in doc["random_positions"]]
    return [f[2:] if f.startswith(("B:B:", "W:W:")) else f for f in fens]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("positions_file", nargs="?", default=str(ROOT / "tests" / "golden" / "ref_fixtures.json"))
    ap.add_argument("warmup", nargs="?", type=int, default=3)
    ap.add_argument("rounds", nargs="?", type=int, default=10)
    ap.add_argument("iterations", nargs="?", type=int, default=1)
    ap.add_argument("--variant", default="standard")
    ap.add_argument("--mode", choices=("batch", "board"), default="batch")
    ap.add_argument("--replicate", type=int, default=1)
    args = ap.parse_args()

    import draughts_b200 as d
    fens = load_positions(Path(args.positions_file))
    moves_total = 0

    if args.mode == "batch":
        work = fens * max(1, args.replicate)

        def one_round():
            nonlocal moves_total
            out = d.BoardBatch.from_fens(args.variant, work).legal_moves()
            moves_total = int(out["offsets"][-1])
    else:
        work = fens

        def one_round():
            nonlocal moves_total
            moves_total = 0
            for fen in work:
                board = d.get_board(args.variant, fen)
                moves_total += len(list(board.legal_moves))

    end = time.perf_counter() + 1.0  # the reference worker runs the workload for a second before measuring
    while time.perf_counter() < end:
        one_round()
    for _ in range(args.warmup):
        one_round()

    iteration_medians = []
    for _ in range(args.iterations):
        gc.collect()
        gc.disable()
        times = []
        for _ in range(args.rounds):
            t0 = time.perf_counter()
            one_round()
            times.append(time.perf_counter() - t0)
        gc.enable()
        iteration_medians.append(median(times) * 1000)

    med = median(iteration_medians)
    print(json.dumps({
        "iteration_medians": iteration_medians,
        "median_ms": med,
        "positions_count": len(fens),
        "high_priority": False,
        "times": [m / 1000 for m in iteration_medians],
        "mode": args.mode, "variant": args.variant, "positions_per_round": len(work), "moves_per_round": moves_total,
        "positions_per_s": len(work) / (med / 1000), "us_per_position": med * 1000 / len(work),
    }))


if __name__ == "__main__":
    main()
