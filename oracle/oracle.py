"""ctypes front-end of the CPU oracle (oracle/drg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It restates, on the CPU, what the reference
(miskibin/py-draughts v1.8.3) computes for the hot path; the product path
(py-draughts_b200/) never touches it.

The helpers below (FEN <-> bitboards, move strings) restate draughts/boards/base.py:408-495
and draughts/move.py:67-77 so that golden vectors written by the Python reference can be
compared without importing either the reference or the product.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

VARIANTS = ["standard", "american", "frisian", "russian", "brazilian", "antidraughts", "breakthrough", "frysk"]
VARIANT_ID = {n: i for i, n in enumerate(VARIANTS)}
SQUARES = {"standard": 50, "american": 32, "frisian": 50, "russian": 32, "brazilian": 32,
           "antidraughts": 50, "breakthrough": 50, "frysk": 50}
MAXP = 64

# include/drg.h DrgMove
MOVE_DTYPE = np.dtype([("captured", "<u8"), ("n_sq", "u1"), ("flags", "u1"), ("path", "u1", (22,))])
assert MOVE_DTYPE.itemsize == 32

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    so = _HERE / "liborc.so"
    src = _HERE / "drg_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        L.orc_init()
        L.orc_perft.restype = C.c_uint64
        L.orc_perft.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int,
                                C.POINTER(C.c_uint64)]
        L.orc_perft_batch.restype = C.c_uint64
        L.orc_expand.restype = C.c_int64
        L.orc_legal_moves.restype = C.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------
# FEN helpers (base.py:408-495)
# ---------------------------------------------------------------------------
def parse_fen(fen: str, S: int):
    """-> (wm, wk, bm, bk, turn) with turn 0 = white, 1 = black.  Restates BaseBoard.from_fen."""
    fen = fen.upper()
    fen = re.sub(r"(G[0-9]+|P[0-9]+)(,|)", "", fen)
    wrap = re.search(r'\[FEN\s*"([^"]*)"\]', fen)
    if wrap:
        fen = wrap.group(1)
    fields = fen.split(":")
    if len(fields) == 4 and fields[0] in ("W", "B"):
        del fields[0]
        fen = ":".join(fields)
    turn_m = re.search(r"[WB]:", fen)
    white_m = re.search(r":W([0-9K,]*)", fen)
    black_m = re.search(r":B([0-9K,]*)", fen)
    if not turn_m or not white_m or not black_m:
        raise ValueError(f"Invalid FEN: {fen}")
    pos = [0] * S
    for s in white_m.group(1).split(","):
        if s.isdigit():
            pos[int(s) - 1] = -1
        elif s.startswith("K"):
            pos[int(s[1:]) - 1] = -2
    for s in black_m.group(1).split(","):
        if s.isdigit():
            pos[int(s) - 1] = 1
        elif s.startswith("K"):
            pos[int(s[1:]) - 1] = 2
    wm = wk = bm = bk = 0
    for sq, v in enumerate(pos):
        if v == -1:
            wm |= 1 << sq
        elif v == -2:
            wk |= 1 << sq
        elif v == 1:
            bm |= 1 << sq
        elif v == 2:
            bk |= 1 << sq
    return wm, wk, bm, bk, (0 if turn_m.group(0)[0] == "W" else 1)


def make_fen(wm, wk, bm, bk, turn, S) -> str:
    """Restates BaseBoard.fen (base.py:408-433)."""
    w, b = [], []
    for sq in range(S):
        bit = 1 << sq
        if wm & bit:
            w.append(str(sq + 1))
        elif wk & bit:
            w.append(f"K{sq + 1}")
        if bm & bit:
            b.append(str(sq + 1))
        elif bk & bit:
            b.append(f"K{sq + 1}")
    return f'[FEN "{"W" if turn == 0 else "B"}:W{",".join(w)}:B{",".join(b)}"]'


def start_position(variant: str):
    """(wm, wk, bm, bk, turn) of the start position: standard.py:93-97, american.py:87-91, frysk.py:33-37."""
    if variant == "frysk":
        return (((1 << 5) - 1) << 45, 0, (1 << 5) - 1, 0, 0)
    if SQUARES[variant] == 50:
        return (((1 << 20) - 1) << 30, 0, (1 << 20) - 1, 0, 0)
    return (((1 << 12) - 1) << 20, 0, (1 << 12) - 1, 0, 0)


# ---------------------------------------------------------------------------
# single-board verbose API
# ---------------------------------------------------------------------------
def legal_moves(variant: str, wm, wk, bm, bk, turn):
    """list of dicts {sq, cap, ent, promo} in canonical order (Move fields of move.py:28-37)."""
    L = lib()
    cap = 64
    while True:
        n_sq = np.zeros(cap, np.int32)
        n_cap = np.zeros(cap, np.int32)
        promo = np.zeros(cap, np.int32)
        sq = np.zeros((cap, MAXP), np.uint8)
        cs = np.zeros((cap, MAXP), np.uint8)
        en = np.zeros((cap, MAXP), np.int8)
        n = L.orc_legal_moves(VARIANT_ID[variant], C.c_uint64(wm), C.c_uint64(wk), C.c_uint64(bm), C.c_uint64(bk),
                              int(turn), cap, _ptr(n_sq), _ptr(n_cap), _ptr(promo), _ptr(sq), _ptr(cs), _ptr(en))
        if n <= cap:
            break
        cap = n
    out = []
    for i in range(n):
        out.append({"sq": sq[i, : n_sq[i]].tolist(), "cap": cs[i, : n_cap[i]].tolist(),
                    "ent": en[i, : n_cap[i]].tolist(), "promo": bool(promo[i])})
    return out


def move_str(m) -> str:
    """Move.__str__ (move.py:67-77)."""
    if not m["cap"]:
        return f"{m['sq'][0] + 1}-{m['sq'][-1] + 1}"
    return "x".join(str(s + 1) for s in m["sq"])


def push_index(variant: str, wm, wk, bm, bk, turn, clock, idx):
    """board.push(board.legal_moves[idx]) -> new (wm, wk, bm, bk, turn, clock); raises ValueError like base.py:241-245."""
    L = lib()
    a = [C.c_uint64(wm), C.c_uint64(wk), C.c_uint64(bm), C.c_uint64(bk)]
    t, c = C.c_int(int(turn)), C.c_int(int(clock))
    rc = L.orc_push_index(VARIANT_ID[variant], C.byref(a[0]), C.byref(a[1]), C.byref(a[2]), C.byref(a[3]),
                          C.byref(t), C.byref(c), int(idx))
    if rc == -1:
        raise ValueError("illegal move: source square holds no piece of the side to move")
    if rc != 0:
        raise IndexError(idx)
    return a[0].value, a[1].value, a[2].value, a[3].value, t.value, c.value


def push_all(variant: str, wm, wk, bm, bk, turn, clock):
    """board.push(m) for every m in legal_moves, each on a fresh copy.
    -> list of (wm, wk, bm, bk, turn, clock, promo_after) or None where the reference raises ValueError."""
    L = lib()
    cap = 64
    while True:
        o = [np.zeros(cap, np.uint64) for _ in range(4)]
        ot, oc, rc, pr = (np.zeros(cap, np.int32) for _ in range(4))
        n = L.orc_push_all(VARIANT_ID[variant], C.c_uint64(wm), C.c_uint64(wk), C.c_uint64(bm), C.c_uint64(bk),
                           int(turn), int(clock), cap, *[_ptr(a) for a in o], _ptr(ot), _ptr(oc), _ptr(rc), _ptr(pr))
        if n <= cap:
            break
        cap = n
    return [None if rc[i] else (int(o[0][i]), int(o[1][i]), int(o[2][i]), int(o[3][i]), int(ot[i]), int(oc[i]),
                                bool(pr[i])) for i in range(n)]


def perft(variant: str, wm, wk, bm, bk, turn, depth) -> int:
    ill = C.c_uint64(0)
    return int(lib().orc_perft(VARIANT_ID[variant], wm, wk, bm, bk, int(turn), int(depth), C.byref(ill)))


# ---------------------------------------------------------------------------
# batched API on numpy SoA arrays (same layout as include/drg.h)
# ---------------------------------------------------------------------------
def _soa(wm, wk, bm, bk, turn):
    return (np.ascontiguousarray(wm, np.uint64), np.ascontiguousarray(wk, np.uint64),
            np.ascontiguousarray(bm, np.uint64), np.ascontiguousarray(bk, np.uint64),
            np.ascontiguousarray(turn, np.uint8))


def batch_count(variant: str, wm, wk, bm, bk, turn, threads: int = 1) -> np.ndarray:
    wm, wk, bm, bk, turn = _soa(wm, wk, bm, bk, turn)
    n = len(wm)
    counts = np.zeros(n, np.uint32)
    lib().orc_batch_count(VARIANT_ID[variant], C.c_int64(n), _ptr(wm), _ptr(wk), _ptr(bm), _ptr(bk), _ptr(turn),
                          _ptr(counts), int(threads))
    return counts


def batch_emit(variant: str, wm, wk, bm, bk, turn, want_mask=True, want_tensor=True, threads: int = 1,
               out=None):
    """-> dict(counts, offsets, moves, mask, tensor); two oracle passes (count, then emit)."""
    wm, wk, bm, bk, turn = _soa(wm, wk, bm, bk, turn)
    n = len(wm)
    S = SQUARES[variant]
    counts = batch_count(variant, wm, wk, bm, bk, turn, threads)
    offsets = np.zeros(n + 1, np.uint64)
    np.cumsum(counts, out=offsets[1:])
    total = int(offsets[-1])
    if out is None:
        moves = np.zeros(max(total, 1), MOVE_DTYPE)
        mask = np.zeros((n, S * S), np.uint8) if want_mask else None
        tensor = np.zeros((n, 4, S), np.float32) if want_tensor else None
    else:
        moves, mask, tensor = out
    lib().orc_batch_emit(VARIANT_ID[variant], C.c_int64(n), _ptr(wm), _ptr(wk), _ptr(bm), _ptr(bk), _ptr(turn),
                         _ptr(offsets), _ptr(moves), _ptr(mask), _ptr(tensor), None, int(threads))
    return {"counts": counts, "offsets": offsets, "moves": moves[:total], "mask": mask, "tensor": tensor}


def batch_apply(variant: str, wm, wk, bm, bk, turn, clock, chosen):
    wm, wk, bm, bk, turn = (a.copy() for a in _soa(wm, wk, bm, bk, turn))
    clock = np.ascontiguousarray(clock, np.uint8).copy()
    chosen = np.ascontiguousarray(chosen, MOVE_DTYPE)
    status = np.zeros(len(wm), np.uint8)
    lib().orc_batch_apply(VARIANT_ID[variant], C.c_int64(len(wm)), _ptr(wm), _ptr(wk), _ptr(bm), _ptr(bk),
                          _ptr(turn), _ptr(clock), _ptr(chosen), _ptr(status))
    return wm, wk, bm, bk, turn, clock, status


def perft_batch(variant: str, wm, wk, bm, bk, turn: int, depth: int, threads: int = 1) -> int:
    wm, wk, bm, bk, _ = _soa(wm, wk, bm, bk, np.zeros(len(wm), np.uint8))
    ill = C.c_uint64(0)
    return int(lib().orc_perft_batch(VARIANT_ID[variant], C.c_int64(len(wm)), _ptr(wm), _ptr(wk), _ptr(bm),
                                     _ptr(bk), int(turn), int(depth), int(threads), C.byref(ill)))


def expand(variant: str, wm, wk, bm, bk, turn: int):
    """children (canonical order) of a batch of same-turn boards -> (wm, wk, bm, bk) arrays."""
    wm, wk, bm, bk, _ = _soa(wm, wk, bm, bk, np.zeros(len(wm), np.uint8))
    cap = max(64, 48 * len(wm))
    while True:
        o = [np.zeros(cap, np.uint64) for _ in range(4)]
        k = int(lib().orc_expand(VARIANT_ID[variant], C.c_int64(len(wm)), _ptr(wm), _ptr(wk), _ptr(bm), _ptr(bk),
                                 int(turn), C.c_int64(cap), *[_ptr(a) for a in o]))
        if k <= cap:
            return tuple(a[:k] for a in o)
        cap = k


def rollout(variant: str, n: int, seed: int, game_id0: int = 0, max_plies: int = 1000, stop_ply=None,
            flags: int = 1, start=None, threads: int = 1):
    """-> dict(wm, wk, bm, bk, turn, clock, result, plies) after random self-play (see include/drg.h drg_rollout)."""
    if start is None:
        s = start_position(variant)
        wm = np.full(n, s[0], np.uint64)
        wk = np.full(n, s[1], np.uint64)
        bm = np.full(n, s[2], np.uint64)
        bk = np.full(n, s[3], np.uint64)
        turn = np.full(n, s[4], np.uint8)
        clock = np.zeros(n, np.uint8)
    else:
        wm, wk, bm, bk, turn = (a.copy() for a in _soa(*start[:5]))
        clock = np.ascontiguousarray(start[5], np.uint8).copy()
    result = np.zeros(n, np.uint8)
    plies = np.zeros(n, np.uint16)
    sp = None if stop_ply is None else np.ascontiguousarray(stop_ply, np.uint16)
    lib().orc_rollout(VARIANT_ID[variant], C.c_int64(n), C.c_uint64(seed), C.c_uint64(game_id0), int(max_plies),
                      _ptr(sp), C.c_uint32(flags), _ptr(wm), _ptr(wk), _ptr(bm), _ptr(bk), _ptr(turn), _ptr(clock),
                      _ptr(result), _ptr(plies), int(threads))
    return {"wm": wm, "wk": wk, "bm": bm, "bk": bk, "turn": turn, "clock": clock, "result": result, "plies": plies}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def pst(S: int) -> np.ndarray:
    """AlphaBetaEngine._get_pst_tables(S, rows) (alpha_beta.py:166-177) -> float64[4, S]."""
    out = np.zeros((4, S), np.float64)
    lib().orc_pst(int(S), _ptr(out))
    return out


def evaluate(variant: str, wm, wk, bm, bk, turn) -> np.ndarray:
    """AlphaBetaEngine.evaluate (alpha_beta.py:205-244) per board -> float64[n], side-to-move relative."""
    wm, wk, bm, bk, turn = _soa(wm, wk, bm, bk, turn)
    out = np.zeros(len(wm), np.float64)
    lib().orc_evaluate(VARIANT_ID[variant], C.c_int64(len(wm)), _ptr(wm), _ptr(wk), _ptr(bm), _ptr(bk), _ptr(turn),
                       _ptr(out))
    return out
