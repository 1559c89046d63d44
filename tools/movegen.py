#!/usr/bin/env python3
"""Batched legal-move generation from a FEN file (FEN file in, JSON out).

    python tools/movegen.py positions.json [--variant standard] [--mask] [--limit N] > moves.ndjson

Input: `{"positions": [fen, ...]}` (the reference's tools/random_positions.json layout), a JSON list of FENs, or a
text file with one FEN per line.  Output: one JSON object per position, in input order:
`{"fen": ..., "moves": ["31-27", "26x17x8", ...], "mask_indices": [from*S+to, ...]}` -- the strings are the
reference's `str(Move)` (move.py:67-77), the order is the reference's `legal_moves` order.  All positions go
through ONE `BoardBatch.legal_moves` call (drg_legal_moves_host)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))


def read_fens(path: Path) -> list[str]:
    text = path.read_text()
    try:
        doc = json.loads(text)
    except json.JSONDecodeError:
        return [ln.strip() for ln in text.splitlines() if ln.strip()]
    if isinstance(doc, dict):
        doc = doc["positions"] if "positions" in doc else [r["src_fen"] for r in doc["random_positions"]]
    return [f[2:] if f.startswith(("B:B:", "W:W:")) else f for f in doc]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("positions_file")
    ap.add_argument("--variant", default="standard")
    ap.add_argument("--mask", action="store_true", help="also print the set indices of legal_moves_mask()")
    ap.add_argument("--limit", type=int, default=0)
    args = ap.parse_args()
    import numpy as np
    import draughts_b200 as d
    fens = read_fens(Path(args.positions_file))
    if args.limit:
        fens = fens[: args.limit]
    batch = d.BoardBatch.from_fens(args.variant, fens)
    out = batch.legal_moves(want_mask=args.mask)
    canon = batch.fens()
    off = out["offsets"].astype(np.int64)
    S = batch.S
    for i, fen in enumerate(fens):
        wm, wk, bm, bk = (int(a[i]) for a in (batch.wm, batch.wk, batch.bm, batch.bk))

        def cell(sq, wm=wm, wk=wk, bm=bm, bk=bk):  # base.py:152-163 priority
            bit = 1 << sq
            return -1 if wm & bit else -2 if wk & bit else 1 if bm & bit else 2 if bk & bit else 0

        rec = {"fen": canon[i],
               "moves": [str(d.Move.from_record(r, cell, S)) for r in out["moves"][off[i]:off[i + 1]]]}
        if args.mask:
            rec["mask_indices"] = np.flatnonzero(out["mask"][i]).tolist()
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
