#!/usr/bin/env python3
"""tests/golden/pdn_games.json: real games (the reference's test/games/<variant>/random_pdns.json fixtures, lidraughts
records) replayed by the UNMODIFIED reference's from_pdn -- the strongest rule-conformance check it has
(test/test_pdn.py:25-46).  Stored: the PDN text (input), the moves as the reference parsed them, final FEN/result.

    PYTHONPATH=/root/reference python tools/gen_golden_pdn.py
"""
import json
import sys
from pathlib import Path

REF = Path("/root/reference")
sys.path.insert(0, str(REF))
import draughts  # noqa: E402
from draughts import (AmericanBoard, AntidraughtsBoard, BrazilianBoard, BreakthroughBoard, FrisianBoard,  # noqa: E402
                      FryskBoard, RussianBoard, StandardBoard)

BOARDS = {"standard": StandardBoard, "american": AmericanBoard, "frisian": FrisianBoard, "russian": RussianBoard,
          "brazilian": BrazilianBoard, "antidraughts": AntidraughtsBoard, "breakthrough": BreakthroughBoard,
          "frysk": FryskBoard}
PER_VARIANT = 10
out = {"meta": {"reference_version": draughts.__version__, "generator": "tools/gen_golden_pdn.py",
                "source": "test/games/<variant>/random_pdns.json (first %d games each)" % PER_VARIANT}, "games": {}}
for v, cls in BOARDS.items():
    data = json.loads((REF / "test" / "games" / v / "random_pdns.json").read_text())
    pdns = (data.get("games") or data.get("pdn_positions") or [])[:PER_VARIANT]
    games = []
    for pdn in pdns:
        b = cls.from_pdn(pdn)
        games.append({"pdn": pdn, "moves": [str(m) for m in b._moves_stack], "fen": b.fen, "result": b.result,
                      "clock": b.halfmove_clock, "pdn_out": b.pdn})
    out["games"][v] = games
    print(v, len(games), sum(len(g["moves"]) for g in games))
p = Path(__file__).resolve().parent.parent / "tests" / "golden" / "pdn_games.json"
p.write_text(json.dumps(out, separators=(",", ":")))
print(p.stat().st_size // 1024, "KiB")
