#!/usr/bin/env python3
"""One-off pin of the CPU oracle to the UNMODIFIED Python reference on far more positions than tests/golden holds.

    PYTHONPATH=/root/reference python tools/validate_oracle.py [--per-variant 40000] [--seed 555]

Runs in the authoring container only (needs /root/reference; no GPU).  For every variant: random-play positions
(oracle sampler, snapshots up to ply 160) and synthetic king endgames / dense boards (tools/fuzz_parity.py's
generator).  For each position the reference's `legal_moves` (move strings, captured squares, captured entities, in
order) and the position after `push` of every move (FEN + halfmove clock, or ValueError) must equal the oracle's.
Prints one JSON line per variant; the log of the round is profiles/r1_q_oracle_vs_reference.log."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import numpy as np  # noqa: E402
from loguru import logger  # noqa: E402

logger.remove()

from draughts import (AmericanBoard, AntidraughtsBoard, BrazilianBoard, BreakthroughBoard,  # noqa: E402
                      FrisianBoard, FryskBoard, RussianBoard, StandardBoard)
from draughts.models import Color  # noqa: E402
from oracle import oracle as O  # noqa: E402
from fuzz_parity import synthetic  # noqa: E402

BOARDS = {"standard": StandardBoard, "american": AmericanBoard, "frisian": FrisianBoard, "russian": RussianBoard,
          "brazilian": BrazilianBoard, "antidraughts": AntidraughtsBoard, "breakthrough": BreakthroughBoard,
          "frysk": FryskBoard}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-variant", type=int, default=40000)
    ap.add_argument("--seed", type=int, default=555)
    args = ap.parse_args()
    bad = 0
    for variant, cls in BOARDS.items():
        t0 = time.perf_counter()
        S = cls.SQUARES_COUNT
        n_play = args.per_variant * 5 // 8
        rng = np.random.default_rng(args.seed)
        stop = rng.integers(0, 161, n_play).astype(np.uint16)
        r = O.rollout(variant, n_play, args.seed, stop_ply=stop, flags=0, threads=O.host_threads())
        syn = synthetic(variant, S, args.per_variant - n_play, rng)
        cols = [np.concatenate([r[k], s]) for k, s in zip(("wm", "wk", "bm", "bk", "turn"), syn)]
        moves = pushes = mism = 0
        for i in range(len(cols[0])):
            wm, wk, bm, bk, turn = (int(c[i]) for c in cols)
            fen = O.make_fen(wm, wk, bm, bk, turn, S)

            def make():  # bitboards set directly: Russian positions may hold two pieces on a square,
                b = cls()  # which a FEN cannot say
                b.white_men, b.white_kings, b.black_men, b.black_kings = wm, wk, bm, bk
                b.turn = Color.WHITE if turn == 0 else Color.BLACK
                return b
            board = make()
            ref = [(str(m), list(m.captured_list), [int(e) for e in m.captured_entities]) for m in board.legal_moves]
            got = [(O.move_str(m), m["cap"], m["ent"]) for m in O.legal_moves(variant, wm, wk, bm, bk, turn)]
            moves += len(ref)
            if ref != got:
                mism += 1
                print(f"MISMATCH {variant} legal_moves {fen}", file=sys.stderr)
                continue
            if i % 4 == 0 and len(ref) <= 64:   # push of every move on a quarter of the positions
                want = []
                for k in range(len(ref)):
                    b2 = make()
                    try:
                        b2.push(b2.legal_moves[k])
                        want.append((b2.white_men, b2.white_kings, b2.black_men, b2.black_kings,
                                     0 if b2.turn == Color.WHITE else 1, int(b2.halfmove_clock)))
                    except ValueError:
                        want.append(None)
                res = O.push_all(variant, wm, wk, bm, bk, turn, 0)
                have = [None if x is None else tuple(x[:6]) for x in res]
                pushes += len(want)
                if want != have:
                    mism += 1
                    print(f"MISMATCH {variant} push {fen}", file=sys.stderr)
        bad += mism
        print(json.dumps({"variant": variant, "positions": len(cols[0]), "moves": moves, "pushes": pushes,
                          "mismatch": mism, "seconds": round(time.perf_counter() - t0, 1)}), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
