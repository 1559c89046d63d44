#!/usr/bin/env python3
"""Generate tests/golden/*.json by importing the UNMODIFIED Python reference.

Run in the authoring container only (the reference does not travel to the GPU box):

    PYTHONPATH=/root/reference python tools/gen_golden.py

Everything written here is produced by the reference's own public API
(Board.from_fen / legal_moves / push / pop / fen / legal_moves_mask / to_tensor /
game_over / result, draughts/boards/*.py, draughts/move.py).  The vectors pin the
CPU oracle (oracle/drg_oracle.c) in tests/test_oracle_golden.py; the CUDA path is
then compared with the oracle at scale.
"""
from __future__ import annotations

import hashlib
import json
import random
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
sys.path.insert(0, str(REF))

import draughts  # noqa: E402
from draughts import (AmericanBoard, AntidraughtsBoard, BrazilianBoard, BreakthroughBoard,  # noqa: E402
                      FrisianBoard, FryskBoard, RussianBoard, StandardBoard)
from draughts.models import Color  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

BOARDS = {  # order of test/_test_helpers.py:19-28
    "standard": StandardBoard, "american": AmericanBoard, "frisian": FrisianBoard, "russian": RussianBoard,
    "brazilian": BrazilianBoard, "antidraughts": AntidraughtsBoard, "breakthrough": BreakthroughBoard,
    "frysk": FryskBoard,
}
MASK64 = (1 << 64) - 1


def bare_fen(board) -> str:
    f = board.fen
    return f[len('[FEN "'):-2]


def move_entry(m):
    """[str(move), captured_list, captured_entities, is_promotion-as-generated]"""
    return [str(m), list(m.captured_list), [int(e) for e in m.captured_entities], bool(m.is_promotion)]


def position_record(board, full: bool):
    """Everything the reference says about one position.  `full` keeps the post-push FENs verbatim."""
    S = board.SQUARES_COUNT
    moves = board.legal_moves
    rec = {"fen": bare_fen(board), "clock": int(board.halfmove_clock), "moves": [move_entry(m) for m in moves]}
    after = []
    for m in board.legal_moves:  # fresh Move objects: push mutates is_promotion (Q8)
        try:
            board.push(m)
        except ValueError:
            after.append("ValueError")
            continue
        after.append(f"{bare_fen(board)}|{board.halfmove_clock}|{int(bool(m.is_promotion))}")
        board.pop()
    if full:
        rec["after"] = after
    rec["after_sha"] = hashlib.sha256("\n".join(after).encode()).hexdigest()[:16]
    mask = board.legal_moves_mask()
    assert mask.dtype == np.bool_ and mask.shape == (S * S,)
    rec["mask"] = np.flatnonzero(mask).tolist()
    t = board.to_tensor()
    assert t.dtype == np.float32 and t.shape == (4, S)
    rec["tensor"] = np.flatnonzero(t.ravel()).tolist()
    assert set(np.unique(t)) <= {0.0, 1.0}
    if len(moves) > 400:  # keep huge Frisian fan-outs as digests only
        rec["moves_sha"] = hashlib.sha256(json.dumps(rec["moves"], separators=(",", ":")).encode()).hexdigest()[:16]
        rec["n_moves"] = len(rec["moves"])
        del rec["moves"]
        rec.pop("after", None)
    return rec


def random_play_positions(variant, seeds, max_plies):
    """test/_test_helpers.py:44-63 scheme; one snapshot per seed after U{0..max_plies} plies."""
    out = []
    for seed in seeds:
        rng = random.Random(seed)
        board = BOARDS[variant]()
        plies = rng.randint(0, max_plies)
        for _ in range(plies):
            lm = list(board.legal_moves)
            if not lm:
                break
            board.push(rng.choice(lm))
        out.append(board)
    return out


def king_fuzz(variant, n, seed):
    """4-16 distinct squares, first 1-3 white kings, rest black (75 % men), white to move."""
    rng = random.Random(seed)
    cls = BOARDS[variant]
    S = cls.SQUARES_COUNT
    out = []
    for _ in range(n):
        k = rng.randint(4, 16)
        sqs = rng.sample(range(S), k)
        nk = rng.randint(1, 3)
        pos = np.zeros(S, dtype=np.int8)
        for i, s in enumerate(sqs):
            if i < nk:
                pos[s] = -2
            else:
                pos[s] = 1 if rng.random() < 0.75 else 2
        turn = Color.WHITE
        if rng.random() < 0.3:  # also exercise the black side: colour-swap
            pos = -pos
            turn = Color.BLACK
        out.append(cls(pos, turn))
    return out


def dense_fuzz(variant, n, seed):
    """random mixed positions with men of both colours (men captures, promotion rows, Q12 men on promo row)."""
    rng = random.Random(seed)
    cls = BOARDS[variant]
    S = cls.SQUARES_COUNT
    out = []
    for _ in range(n):
        k = rng.randint(3, min(S - 6, 30))
        sqs = rng.sample(range(S), k)
        pos = np.zeros(S, dtype=np.int8)
        for s in sqs:
            pos[s] = rng.choice([-1, -1, -1, 1, 1, 1, -2, 2])
        out.append(cls(pos, rng.choice([Color.WHITE, Color.BLACK])))
    return out


def russian_promo_fuzz(n, seed):
    """SURVEY 8(d).5 second fuzz: one white man on rows 2-3, 3-9 black pieces in rows 0-5 (mid-capture promotion, Q4)."""
    rng = random.Random(seed)
    out = []
    tries = 0
    while len(out) < n and tries < 400000:
        tries += 1
        pos = np.zeros(32, dtype=np.int8)
        man = rng.randrange(8, 16)
        pos[man] = -1
        k = rng.randint(3, 9)
        for s in rng.sample([x for x in range(24) if x != man], k):
            pos[s] = 1 if rng.random() < 0.6 else 2
        b = RussianBoard(pos, Color.WHITE)
        lm = b.legal_moves
        if any(m.is_promotion and len(m.captured_list) >= 2 for m in lm):
            out.append(b)
    return out


def perft(b, d):
    ms = b.legal_moves
    if d == 1:
        return len(ms)
    n = 0
    for m in ms:
        b.push(m)
        n += perft(b, d - 1)
        b.pop()
    return n


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK64
    return x ^ (x >> 31)


def rollout(variant, seed, game_id, max_plies, draw_rules):
    """include/drg.h drg_rollout contract, played by the reference."""
    b = BOARDS[variant]()
    ply = 0
    h = hashlib.sha256()
    while True:
        if draw_rules:
            if b.game_over:
                res = b.result
                break
        lm = b.legal_moves
        if not lm:
            res = b.result if draw_rules else ("0-1" if b.turn == Color.WHITE else "1-0")
            if variant == "antidraughts" and not draw_rules:
                res = "1-0" if b.turn == Color.WHITE else "0-1"
            break
        if ply >= max_plies:
            res = "-"
            break
        x = splitmix64((seed + game_id * 0x9E3779B97F4A7C15 + ply * 0xBF58476D1CE4E5B9) & MASK64)
        idx = ((x >> 32) * len(lm)) >> 32
        m = lm[idx]
        h.update((str(m) + " ").encode())
        b.push(m)
        ply += 1
    return {"seed": seed, "game_id": game_id, "max_plies": max_plies, "draw_rules": draw_rules, "result": res,
            "plies": ply, "fen": bare_fen(b), "clock": b.halfmove_clock, "moves_sha": h.hexdigest()[:16]}


def dump(name, obj):
    p = OUT / name
    p.write_text(json.dumps(obj, separators=(",", ":")))
    print(f"{name}: {p.stat().st_size / 1024:.0f} KiB")


def main():
    meta = {"reference_version": draughts.__version__, "generator": "tools/gen_golden.py"}

    # ---- geometry tables ---------------------------------------------------
    from draughts.boards import american as am, frisian as fr, russian as ru, standard as st
    dump("tables.json", {
        "meta": meta,
        "standard": {"MOVE_TGT": st.MOVE_TGT, "JUMP_TGT": st.JUMP_TGT, "KING_RAYS": st.KING_RAYS},
        "american": {"MOVE_TGT": am.MOVE_TGT, "JUMP_TGT": am.JUMP_TGT},
        "russian": {"MOVE_TGT": ru.MOVE_TGT, "JUMP_TGT": ru.JUMP_TGT, "KING_RAYS": ru.KING_RAYS},
        "frisian": {"MOVE_TGT": fr.DIAG_MOVE_TGT, "JUMP_TGT": fr.DIAG_JUMP_TGT, "KING_RAYS": fr.KING_DIAG_RAYS,
                    "ORTHO_RAYS": fr.ORTHO_RAYS, "ORTHO_JUMP": fr.ORTHO_JUMP},
    })

    # ---- the reference's own fixtures ---------------------------------------
    fx = {"meta": meta}
    fx["legal_moves_len"] = json.loads((REF / "test/games/standard/legal_moves_len.json").read_text())
    fens = json.loads((REF / "tools/random_positions.json").read_text())["positions"]
    lines = []
    recs = []
    for fen in fens:
        b = StandardBoard.from_fen(fen)
        lines.append(fen + "|" + ",".join(str(m) for m in b.legal_moves))
        recs.append(position_record(b, full=True))
        recs[-1]["src_fen"] = fen
    fx["random_positions_sha256"] = hashlib.sha256("\n".join(lines).encode()).hexdigest()
    fx["random_positions"] = recs
    # exact strings the reference's tests assert (test_standard_board.py:84-161, test_russian_board.py:179-247)
    named = {
        "standard": ["W:WK50:B44,33,28,31", "W:WK49:B43,27,24,38", "W:WK4:B9,22,32,43,20",
                     "W:WK43:B8,9,11,14,19,38,39,42,47,48", "W:W28:B23,13,14,24", "B:W28,38:BK5"],
        "russian": ["W:W10:B6,K9,K15,K17,K23", "W:W22:B18,10,11", "W:W18:B14,7", "B:W23,27:B18",
                    "W:WK29:B25,18,11", "W:W9:B5,6"],
        "frisian": ["W:WK1,K15,K20:BK4,7,9,10,12,17,21,26,29,34,40,41,44",
                    "W:WK27:B1,5,10,11,17,20,22,38,39", "W:W28:B23,22,18", "W:W33,K5:B28,K10"],
        "american": ["W:W22,K30:B18,17,26,K9", "B:W22,23:B18,K14"],
        "brazilian": ["W:W22:B18,10,11", "W:WK29:B25,18,11,K7"],
        "frysk": ["W:W28,K46:B23,22,33"],
    }
    fx["named"] = {v: [position_record(BOARDS[v].from_fen(f), full=True) for f in fl] for v, fl in named.items()}
    dump("ref_fixtures.json", fx)

    # ---- perft tables ---------------------------------------------------------
    depth = {"standard": 7, "american": 8, "frisian": 6, "russian": 8, "brazilian": 8, "antidraughts": 5,
             "breakthrough": 5, "frysk": 6}
    pf = {"meta": meta, "start": {}, "fen": []}
    for v, dmax in depth.items():
        pf["start"][v] = [perft(BOARDS[v](), d) for d in range(1, dmax + 1)]
        print(v, pf["start"][v], flush=True)
    # SURVEY 8(c) deep values measured with the same recursion (multi-process) in the survey session
    pf["survey_deep"] = {"standard": {"8": 6483971, "9": 41022614}, "american": {"9": 109416570, "10": 906645019},
                         "russian": {"9": 4571392}, "frisian": {"7": 550314, "8": 2907905},
                         "frysk": {"7": 2572197, "8": 20147712}}
    for v, fen, d in [("standard", "W:WK4,K5,29,31:B18,K20,22", 4), ("standard", "B:W21,K33:B5,6,14,K36", 5),
                      ("russian", "W:WK29,22:B18,10,11,K3", 5), ("frisian", "W:W28,33,K46:B23,22,18,K5", 4),
                      ("american", "W:W22,K30:B18,17,26,K9", 6), ("brazilian", "B:WK29,22,23:B18,10,11,K3", 5)]:
        b = BOARDS[v].from_fen(fen)
        pf["fen"].append({"variant": v, "fen": fen, "perft": [perft(b, k) for k in range(1, d + 1)]})
    dump("perft.json", pf)

    # ---- move lists ------------------------------------------------------------
    for v in BOARDS:
        boards = []
        boards += random_play_positions(v, range(300), 100 if v != "american" else 160)
        if v not in ("antidraughts", "breakthrough"):
            boards += king_fuzz(v, 260 if v not in ("frisian", "frysk") else 160, 7)
            boards += dense_fuzz(v, 200, 11)
        else:
            boards += king_fuzz(v, 40, 7)
        if v == "russian":
            boards += russian_promo_fuzz(150, 13)
        recs = []
        for i, b in enumerate(boards):
            recs.append(position_record(b, full=(i % 10 == 0)))
        dump(f"movelists_{v}.json", {"meta": meta, "variant": v, "positions": recs})

    # ---- rollouts --------------------------------------------------------------
    ro = {"meta": meta, "games": {}}
    for v in BOARDS:
        games = []
        for g in range(24):
            games.append(rollout(v, 12345, g, 400, True))
        for g in range(8):
            games.append(rollout(v, 777, 1000 + g, 30 + 5 * g, False))
        ro["games"][v] = games
    dump("rollouts.json", ro)


if __name__ == "__main__":
    main()
