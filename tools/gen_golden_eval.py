#!/usr/bin/env python3
"""Golden vectors for the leaf-evaluation kernel, written by the UNMODIFIED reference.

    PYTHONPATH=/root/reference python tools/gen_golden_eval.py        # -> tests/golden/evaluate.json

For every variant: the four piece-square tables `AlphaBetaEngine._get_pst_tables(S, rows)` (alpha_beta.py:27-57,
166-177; rows chosen as in get_best_move, alpha_beta.py:271-276) and `AlphaBetaEngine.evaluate(board)`
(alpha_beta.py:205-244) for the FENs of tests/golden/movelists_<variant>.json plus their first successors.  Doubles
are stored as `float.hex()` strings so the comparison is bit-exact."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))

from loguru import logger  # noqa: E402

logger.remove()

import draughts  # noqa: E402
from draughts import AlphaBetaEngine  # noqa: E402
from draughts.boards.american import Board as American  # noqa: E402
from draughts.boards.antidraughts import Board as Anti  # noqa: E402
from draughts.boards.brazilian import Board as Brazilian  # noqa: E402
from draughts.boards.breakthrough import Board as Breakthrough  # noqa: E402
from draughts.boards.frisian import Board as Frisian  # noqa: E402
from draughts.boards.frysk import Board as Frysk  # noqa: E402
from draughts.boards.russian import Board as Russian  # noqa: E402
from draughts.boards.standard import Board as Standard  # noqa: E402

BOARDS = {"standard": Standard, "american": American, "frisian": Frisian, "russian": Russian, "brazilian": Brazilian,
          "antidraughts": Anti, "breakthrough": Breakthrough, "frysk": Frysk}


def main():
    gold = ROOT / "tests" / "golden"
    out = {"meta": {"reference_version": getattr(draughts, "__version__", "?"), "generator": "tools/gen_golden_eval.py"},
           "variants": {}}
    for name, cls in BOARDS.items():
        S = cls.SQUARES_COUNT
        rows = 10 if S == 50 else 8
        eng = AlphaBetaEngine(depth_limit=1)
        pst = eng._get_pst_tables(S, rows)
        eng._current_pst = pst
        recs = json.loads((gold / f"movelists_{name}.json").read_text())["positions"]
        fens = []
        for r in recs:
            fens.append(r["fen"])
            for a in (r.get("after") or [])[:2]:
                fens.append(a.split("|")[0])
        seen, rows_out = set(), []
        for f in fens:
            if f in seen:
                continue
            seen.add(f)
            b = cls.from_fen(f)
            rows_out.append([f, float(eng.evaluate(b)).hex()])
        out["variants"][name] = {"S": S, "rows": rows, "pst": [[float(x).hex() for x in t] for t in pst],
                                 "positions": rows_out}
        print(name, len(rows_out))
    (gold / "evaluate.json").write_text(json.dumps(out, separators=(",", ":")))


if __name__ == "__main__":
    main()
