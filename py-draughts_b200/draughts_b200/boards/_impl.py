"""Board classes with the reference's public API (draughts/boards/base.py:49-891 and the eight
variant modules), backed by the CUDA library: `legal_moves`, `legal_moves_mask`, `to_tensor` and
`push` are batch-of-one calls through the C-ABI (no CPU move generator exists in this package).
FEN / PDN text handling, `pop` bookkeeping and the terminal predicates are plain host code.
"""
from __future__ import annotations

import copy as _copy
import re
from dataclasses import dataclass
from typing import Generator, Literal, Optional

import numpy as np

from .. import _lib
from .._lib import MOVE_DTYPE, VARIANT_ID
from ..batch import BoardBatch, start_bitboards
from ..models import FIGURE_REPR, Color, Figure
from ..move import Move

__all__ = ["BaseBoard", "BoardFeatures", "StandardBoard", "AmericanBoard", "FrisianBoard", "RussianBoard",
           "BrazilianBoard", "AntidraughtsBoard", "BreakthroughBoard", "FryskBoard", "parse_fen", "format_fen"]


# -------------------------------------------------------------------------------------------------
# FEN text (base.py:408-495)
# -------------------------------------------------------------------------------------------------
def parse_fen(fen: str, S: int):
    """'W:W31,K32:B1,2' (optionally wrapped in [FEN "..."], with a legacy leading turn token or
    :H0:F2 tails, empty sides allowed) -> (wm, wk, bm, bk, turn) with turn 0 = white."""
    text = re.sub(r"(G[0-9]+|P[0-9]+)(,|)", "", fen.upper())
    wrapped = re.search(r'\[FEN\s*"([^"]*)"\]', text)
    if wrapped:
        text = wrapped.group(1)
    parts = text.split(":")
    if len(parts) == 4 and parts[0] in ("W", "B"):
        text = ":".join(parts[1:])
    side = re.search(r"[WB]:", text)
    white = re.search(r":W([0-9K,]*)", text)
    black = re.search(r":B([0-9K,]*)", text)
    if not side or not white or not black:
        raise ValueError(f"Invalid FEN: {text}")
    cells = [0] * S
    for listing, man in ((white.group(1), -1), (black.group(1), 1)):
        for tok in listing.split(","):
            if tok.isdigit():
                cells[int(tok) - 1] = man
            elif tok.startswith("K"):
                cells[int(tok[1:]) - 1] = 2 * man
    bbs = {-1: 0, -2: 0, 1: 0, 2: 0}
    for sq, v in enumerate(cells):
        if v:
            bbs[v] |= 1 << sq
    return bbs[-1], bbs[-2], bbs[1], bbs[2], (0 if side.group(0)[0] == "W" else 1)


def format_fen(wm: int, wk: int, bm: int, bk: int, turn: int, S: int) -> str:
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


@dataclass(frozen=True, slots=True)
class BoardFeatures:
    white_men: int
    white_kings: int
    black_men: int
    black_kings: int
    turn: int
    mobility: int
    material_balance: float
    phase: str


def _popcount(x: int) -> int:
    return bin(x).count("1")


class _Scratch1:
    """Reusable batch-of-1 buffers of one variant with their addresses taken once: a single-board call is dominated
    by Python overhead (array construction, ctypes pointer extraction), not by the 27 us library call."""
    _per_variant: dict = {}

    def __init__(self, variant: str, S: int):
        import ctypes as C
        self.vid = VARIANT_ID[variant]
        self.state = np.zeros(4, np.uint64)           # wm, wk, bm, bk
        self.turn = np.zeros(1, np.uint8)
        self.clock = np.zeros(1, np.uint8)
        self.counts = np.zeros(1, np.uint32)
        self.offsets = np.zeros(2, np.uint64)
        self.moves = np.zeros(512, MOVE_DTYPE)
        self.mask = np.zeros((1, S * S), np.uint8)
        self.tensor = np.zeros((1, 4, S), np.float32)
        self.rec = np.zeros(1, MOVE_DTYPE)
        self.status = np.zeros(1, np.uint8)
        self.total = C.c_int64(0)
        self.p_total = C.byref(self.total)
        a = self.state.ctypes.data
        self.p_state = (a, a + 8, a + 16, a + 24)
        self.p_turn, self.p_clock = self.turn.ctypes.data, self.clock.ctypes.data
        self.p_counts, self.p_offsets = self.counts.ctypes.data, self.offsets.ctypes.data
        self.p_moves, self.p_mask, self.p_tensor = self.moves.ctypes.data, self.mask.ctypes.data, self.tensor.ctypes.data
        self.p_rec, self.p_status = self.rec.ctypes.data, self.status.ctypes.data

    @classmethod
    def of(cls, variant: str, S: int) -> "_Scratch1":
        sc = cls._per_variant.get(variant)
        if sc is None:
            sc = cls._per_variant[variant] = cls(variant, S)
        return sc


class BaseBoard:
    GAME_TYPE: int = -1
    VARIANT_NAME: str = "Abstract"
    VARIANT: str = ""
    STARTING_COLOR: Color = Color.WHITE
    SQUARES_COUNT: int = 50
    PROMO_WHITE: int = 0
    PROMO_BLACK: int = 0
    ROW_IDX: dict = {}
    COL_IDX: dict = {}
    STARTING_POSITION: np.ndarray = np.array([], dtype=np.int8)
    SQUARE_NAMES: list = []

    __slots__ = ("white_men", "white_kings", "black_men", "black_kings", "turn", "halfmove_clock", "_moves_stack",
                 "shape")

    def __init__(self, starting_position: Optional[np.ndarray] = None, turn: Optional[Color] = None) -> None:
        size = int(np.sqrt(self.SQUARES_COUNT * 2))
        self.shape = (size, size)
        self.turn = turn if turn is not None else self.STARTING_COLOR
        self.halfmove_clock = 0
        self._moves_stack: list[Move] = []
        if starting_position is not None:
            self._from_array(starting_position)
        else:
            self._init_default_position()

    def _init_default_position(self) -> None:
        self.white_men, self.white_kings, self.black_men, self.black_kings = start_bitboards(self.VARIANT)

    def _from_array(self, arr) -> None:
        bbs = {-1: 0, -2: 0, 1: 0, 2: 0}
        for sq, v in enumerate(arr):
            v = int(v)
            if v in bbs:
                bbs[v] |= 1 << sq
        self.white_men, self.white_kings, self.black_men, self.black_kings = bbs[-1], bbs[-2], bbs[1], bbs[2]

    # -- small helpers ---------------------------------------------------------------------------
    def _all(self) -> int:
        return self.white_men | self.white_kings | self.black_men | self.black_kings

    def _empty(self) -> int:
        return ~self._all() & ((1 << self.SQUARES_COUNT) - 1)

    def _enemy(self) -> int:
        return (self.black_men | self.black_kings) if self.turn == Color.WHITE else (self.white_men | self.white_kings)

    def _get(self, sq: int) -> int:
        bit = 1 << sq
        if self.white_men & bit:
            return -1
        if self.white_kings & bit:
            return -2
        if self.black_men & bit:
            return 1
        if self.black_kings & bit:
            return 2
        return 0

    def _set(self, sq: int, piece: int) -> None:
        keep = ~(1 << sq)
        self.white_men &= keep
        self.white_kings &= keep
        self.black_men &= keep
        self.black_kings &= keep
        if piece == -1:
            self.white_men |= 1 << sq
        elif piece == -2:
            self.white_kings |= 1 << sq
        elif piece == 1:
            self.black_men |= 1 << sq
        elif piece == 2:
            self.black_kings |= 1 << sq

    _popcount = staticmethod(_popcount)

    def _turn_bit(self) -> int:
        return 0 if self.turn == Color.WHITE else 1

    def _batch1(self) -> BoardBatch:
        m = (1 << 64) - 1
        return BoardBatch(self.VARIANT, [self.white_men & m], [self.white_kings & m], [self.black_men & m],
                          [self.black_kings & m], [self._turn_bit()], [min(self.halfmove_clock, 255)])

    # -- the hot path: CUDA through the C-ABI -----------------------------------------------------------
    def _movegen1(self, want_mask: bool = False, want_tensor: bool = False):
        """one drg_legal_moves_host call for this board through the per-variant scratch buffers -> (scratch, total),
        or (None, batch result) when the move list does not fit the scratch (fan-out monsters)"""
        sc = _Scratch1.of(self.VARIANT, self.SQUARES_COUNT)
        m = (1 << 64) - 1
        st = sc.state
        st[0], st[1], st[2], st[3] = self.white_men & m, self.white_kings & m, self.black_men & m, self.black_kings & m
        sc.turn[0] = 0 if self.turn == Color.WHITE else 1
        rc = _lib.lib().drg_legal_moves_host(sc.vid, 1, *sc.p_state, sc.p_turn, sc.p_counts, sc.p_offsets, sc.p_moves, 512,
                                             sc.p_total, sc.p_mask if want_mask else None,
                                             sc.p_tensor if want_tensor else None)
        if rc == _lib.E_OVERFLOW and sc.total.value > 512:
            return None, self._batch1().legal_moves(want_mask=want_mask, want_tensor=want_tensor)
        _lib.check(rc)
        return sc, sc.total.value

    @property
    def legal_moves(self) -> list[Move]:
        sc, res = self._movegen1()
        recs = sc.moves[:res] if sc is not None else res["moves"]
        return Move.list_from_records(recs, self._get, self.SQUARES_COUNT)

    def _gen_captures(self) -> list[Move]:
        """The legal capture moves (the reference's generator of the same name returns the candidates its
        `legal_moves` filters; here the variant's filter has already been applied on the device)."""
        return [m for m in self.legal_moves if m.captured_list]

    def _gen_simple(self) -> list[Move]:
        """The legal quiet moves."""
        return [m for m in self.legal_moves if not m.captured_list]

    def legal_moves_mask(self) -> np.ndarray:
        sc, res = self._movegen1(want_mask=True)
        row = sc.mask[0] if sc is not None else res["mask"][0]
        return row.view(np.bool_).copy()

    def to_tensor(self, perspective: Optional[Color] = None) -> np.ndarray:
        sc, res = self._movegen1(want_tensor=True)
        t = sc.tensor[0] if sc is not None else res["tensor"][0]
        if perspective is not None and perspective != self.turn:
            t = t[[2, 3, 0, 1]]
        return np.array(t, dtype=np.float32)

    def push(self, move: Move, is_finished: bool = True) -> None:
        src, tgt = move.square_list[0], move.square_list[-1]
        piece = self._get(src)
        if piece == 0 or (piece < 0) != (self.turn == Color.WHITE):
            raise ValueError(f"Illegal move {move}: square {src + 1} holds no "
                             f"{'white' if self.turn == Color.WHITE else 'black'} piece to move.")
        if not is_finished:
            raise NotImplementedError("push(is_finished=False) is an internal helper of the reference's generators")
        move.halfmove_clock = self.halfmove_clock
        sc = _Scratch1.of(self.VARIANT, self.SQUARES_COUNT)
        cap = 0
        for c in move.captured_list:
            cap |= 1 << c
        rec = sc.rec
        rec["captured"][0] = cap
        rec["n_sq"][0] = 2
        rec["flags"][0] = (_lib.MF_PROMO if move.is_promotion else 0) | (_lib.MF_CAPTURE if cap else 0)
        rec["path"][0, 0], rec["path"][0, 1] = src, tgt
        m = (1 << 64) - 1
        st = sc.state
        st[0], st[1], st[2], st[3] = self.white_men & m, self.white_kings & m, self.black_men & m, self.black_kings & m
        sc.turn[0] = 0 if self.turn == Color.WHITE else 1
        sc.clock[0] = min(self.halfmove_clock, 255)
        _lib.check(_lib.lib().drg_apply_host(sc.vid, 1, *sc.p_state, sc.p_turn, sc.p_clock, sc.p_rec, sc.p_status))
        if sc.status[0]:
            raise ValueError(f"Illegal move {move}")
        self.white_men, self.white_kings = int(st[0]), int(st[1])
        self.black_men, self.black_kings = int(st[2]), int(st[3])
        if abs(piece) == 1 and (self.white_kings if piece < 0 else self.black_kings) & (1 << tgt):
            move.is_promotion = True
        quiet_king = abs(piece) == 2 and not move.captured_list
        self.halfmove_clock = self.halfmove_clock + 1 if quiet_king else 0  # not capped at 255 like the device clock
        self._moves_stack.append(move)
        self.turn = Color.BLACK if self.turn == Color.WHITE else Color.WHITE

    def pop(self, is_finished: bool = True) -> Move:
        move = self._moves_stack.pop()
        src, tgt = move.square_list[0], move.square_list[-1]
        piece = self._get(tgt)
        if move.is_promotion:
            piece //= 2
        self._set(tgt, 0)
        self._set(src, piece)
        for sq, ent in zip(move.captured_list, move.captured_entities):
            self._set(sq, ent)
        self.halfmove_clock = move.halfmove_clock
        if is_finished:
            self.turn = Color.BLACK if self.turn == Color.WHITE else Color.WHITE
        return move

    def push_uci(self, str_move: str) -> None:
        self.push(Move.from_uci(str_move, self.legal_moves))

    # -- game end ----------------------------------------------------------------------------------
    @property
    def is_threefold_repetition(self) -> bool:
        s = self._moves_stack
        return len(s) >= 9 and s[-1].square_list == s[-5].square_list == s[-9].square_list

    @property
    def is_draw(self) -> bool:
        raise NotImplementedError

    @property
    def game_over(self) -> bool:
        return self.is_draw or not self.legal_moves

    @property
    def result(self) -> Literal["1/2-1/2", "1-0", "0-1", "-"]:
        if self.is_draw:
            return "1/2-1/2"
        if self.game_over:
            return "0-1" if self.turn == Color.WHITE else "1-0"
        return "-"

    @staticmethod
    def is_capture(move: Move) -> bool:
        return bool(move.captured_list)

    # -- text ----------------------------------------------------------------------------------------
    @property
    def fen(self) -> str:
        return format_fen(self.white_men, self.white_kings, self.black_men, self.black_kings, self._turn_bit(),
                          self.SQUARES_COUNT)

    @classmethod
    def from_fen(cls, fen: str) -> "BaseBoard":
        wm, wk, bm, bk, turn = parse_fen(fen, cls.SQUARES_COUNT)
        board = cls.__new__(cls)
        BaseBoard.__init__(board, np.zeros(0, dtype=np.int8), Color.WHITE if turn == 0 else Color.BLACK)
        board.white_men, board.white_kings, board.black_men, board.black_kings = wm, wk, bm, bk
        return board

    @property
    def pdn(self) -> str:
        head = f'[GameType "{self.GAME_TYPE}"]\n[Variant "{self.VARIANT_NAME}"]\n[Result "{self.result}"]\n'
        turns: list[list[str]] = []
        for i, m in enumerate(self._moves_stack):
            if i % 2 == 0:
                turns.append([str(i // 2 + 1), str(m)])
            else:
                turns[-1].append(str(m))
        body = " ".join(f"{t[0]}. {' '.join(t[1:])}" for t in turns)
        return head + body + ("" if self.result == "-" else f" {self.result}")

    @classmethod
    def from_pdn(cls, pdn: str) -> "BaseBoard":
        board = cls()
        names = {name: i for i, name in enumerate(cls.SQUARE_NAMES)} if cls.SQUARE_NAMES else {}
        algebraic = re.findall(r"\b([a-h]\d[-x][a-h]\d)\b", pdn)
        if algebraic and names:
            moves = [cls._alg_to_uci(m, names) for m in algebraic]
        else:
            results = {"2-0", "0-2", "1-1", "1-0", "0-1", "1/2-1/2"}
            moves = [m for m in re.findall(r"\b(\d+[-x]\d+(?:[-x]\d+)*)\b", pdn) if m not in results]
        i, chain_start = 0, None
        while i < len(moves):
            text = moves[i]
            is_cap = "x" in text
            fields = text.split("x" if is_cap else "-")
            start, end = int(fields[0]), int(fields[-1])
            if not is_cap:
                board.push_uci(text)
                chain_start = None
            else:
                src = chain_start or start
                cap = next((m for m in board.legal_moves
                            if m.captured_list and m.square_list[0] == src - 1 and (end - 1) in m.square_list), None)
                if not cap:
                    raise ValueError(f"No legal capture for {text}")
                if i + 1 < len(moves) and "x" in moves[i + 1]:
                    if int(moves[i + 1].split("x")[0]) == end and (end - 1) in cap.square_list[1:-1]:
                        chain_start = src
                        i += 1
                        continue
                board.push(cap)
                chain_start = None
            i += 1
        return board

    @staticmethod
    def _alg_to_uci(move: str, mapping: dict) -> str:
        sep = "x" if "x" in move else "-"
        a, b = move.lower().split(sep)
        return f"{mapping[a] + 1}{sep}{mapping[b] + 1}"

    # -- array views -----------------------------------------------------------------------------------
    @property
    def position(self) -> np.ndarray:
        return np.array([self._get(sq) for sq in range(self.SQUARES_COUNT)], dtype=np.int8)

    @property
    def _pos(self) -> np.ndarray:
        return self.position

    @property
    def friendly_form(self) -> np.ndarray:
        pos, half = self.position, self.shape[0] // 2
        cells = [0]
        for idx, v in enumerate(pos):
            cells.extend([0] * (idx % half != 0))
            cells.extend([0, 0] * (idx % self.shape[0] == 0 and idx != 0))
            cells.append(v)
        cells.append(0)
        return np.array(cells)

    def __repr__(self) -> str:
        pos, n = self.friendly_form, self.shape[0]
        return "".join(f" {FIGURE_REPR[pos[i * n + j]]}" + ("\n" if j == n - 1 else "")
                       for i in range(n) for j in range(n))

    def __str__(self) -> str:
        n = self.shape[0]
        lines = []
        for i, line in enumerate(repr(self).strip().split("\n")):
            sq = iter(range(i * n // 2 + 1, (i + 1) * n // 2 + 1))
            nums = " ".join(f"{next(sq):2d}" if (i + j) % 2 else "." for j in range(n))
            lines.append(f"{line}     {nums}")
        return "\n".join(lines)

    def __iter__(self) -> Generator[int, None, None]:
        for sq in range(self.SQUARES_COUNT):
            yield self._get(sq)

    def __getitem__(self, key: int) -> int:
        return self._get(key)

    # -- copies / ML helpers -------------------------------------------------------------------------
    def copy(self) -> "BaseBoard":
        new = object.__new__(self.__class__)
        new.white_men, new.white_kings = self.white_men, self.white_kings
        new.black_men, new.black_kings = self.black_men, self.black_kings
        new.turn, new.halfmove_clock, new.shape = self.turn, self.halfmove_clock, self.shape
        new._moves_stack = []
        return new

    def __copy__(self) -> "BaseBoard":
        return self.copy()

    def __deepcopy__(self, memo: dict) -> "BaseBoard":
        new = self.copy()
        new._moves_stack = _copy.deepcopy(self._moves_stack, memo)
        return new

    def features(self) -> BoardFeatures:
        wm, wk = _popcount(self.white_men), _popcount(self.white_kings)
        bm, bk = _popcount(self.black_men), _popcount(self.black_kings)
        total = wm + wk + bm + bk
        phase = "opening" if total >= self.SQUARES_COUNT * 0.6 else ("endgame" if total <= 8 else "midgame")
        return BoardFeatures(wm, wk, bm, bk, 1 if self.turn == Color.WHITE else -1, len(self.legal_moves),
                             (wm + 2 * wk) - (bm + 2 * bk), phase)

    def move_to_index(self, move: Move) -> int:
        return move.square_list[0] * self.SQUARES_COUNT + move.square_list[-1]

    def index_to_move(self, index: int) -> Move:
        n = self.SQUARES_COUNT
        src, dst = divmod(int(index), n)
        legal = self.legal_moves
        for m in legal:
            if m.square_list[0] == src and m.square_list[-1] == dst:
                return m
        raise ValueError(f"No legal move from square {src + 1} to {dst + 1}. Legal moves: {list(map(str, legal))}")


# -------------------------------------------------------------------------------------------------
# variants
# -------------------------------------------------------------------------------------------------
def _rows(width: int, count: int):
    return [((1 << width) - 1) << (i * width) for i in range(count)]


_ROW50, _ROW32 = _rows(5, 10), _rows(4, 8)
_NAMES32 = ["b8", "d8", "f8", "h8", "a7", "c7", "e7", "g7", "b6", "d6", "f6", "h6", "a5", "c5", "e5", "g5",
            "b4", "d4", "f4", "h4", "a3", "c3", "e3", "g3", "b2", "d2", "f2", "h2", "a1", "c1", "e1", "g1"]


class _Board50(BaseBoard):
    SQUARES_COUNT = 50
    PROMO_WHITE, PROMO_BLACK = _ROW50[0], _ROW50[9]
    STARTING_POSITION = np.array([1] * 20 + [0] * 10 + [-1] * 20, dtype=np.int8)
    ROW_IDX = {v: v // 5 for v in range(50)}
    COL_IDX = {v: v % 10 for v in range(50)}
    __slots__ = ()


class _Board32(BaseBoard):
    SQUARES_COUNT = 32
    PROMO_WHITE, PROMO_BLACK = _ROW32[0], _ROW32[7]
    STARTING_POSITION = np.array([1] * 12 + [0] * 8 + [-1] * 12, dtype=np.int8)
    ROW_IDX = {v: v // 4 for v in range(32)}
    COL_IDX = {v: v % 8 for v in range(32)}
    SQUARE_NAMES = _NAMES32
    __slots__ = ()


class StandardBoard(_Board50):
    """International draughts (standard.py:73-311)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 20, "Standard (international) checkers", "standard"
    __slots__ = ()

    @property
    def is_draw(self) -> bool:
        return self.is_25_moves_rule or self.is_threefold_repetition or self.is_5_moves_rule or self.is_16_moves_rule

    @property
    def is_25_moves_rule(self) -> bool:
        return self.halfmove_clock >= 50

    @property
    def is_16_moves_rule(self) -> bool:
        if self.halfmove_clock < 32 or _popcount(self._all()) > 4:
            return False
        return _popcount(self.white_kings | self.black_kings) * 2 + _popcount(self.white_men | self.black_men) >= 6

    @property
    def is_5_moves_rule(self) -> bool:
        if _popcount(self._all()) > 3:
            return False
        return (_popcount(self.white_kings | self.black_kings) * 2 + _popcount(self.white_men | self.black_men) >= 5
                and self.halfmove_clock >= 10)


class AntidraughtsBoard(StandardBoard):
    """Standard rules, inverted result (antidraughts.py:19-37)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 20, "Antidraughts", "antidraughts"
    __slots__ = ()

    @property
    def result(self):
        if self.is_draw:
            return "1/2-1/2"
        if not self.legal_moves:
            return "1-0" if self.turn == Color.WHITE else "0-1"
        return "-"


class BreakthroughBoard(StandardBoard):
    """Standard rules, first king wins (breakthrough.py:16-39)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 20, "Breakthrough", "breakthrough"
    __slots__ = ()

    @property
    def game_over(self) -> bool:
        if self.white_kings or self.black_kings:
            return True
        return self.is_draw or not self.legal_moves

    @property
    def result(self):
        if self.white_kings:
            return "1-0"
        if self.black_kings:
            return "0-1"
        return super().result


class FrisianBoard(_Board50):
    """Frisian draughts (frisian.py:211-659)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 40, "Frisian", "frisian"
    __slots__ = ()

    @property
    def is_draw(self) -> bool:
        return self.is_threefold_repetition or self.is_5_moves_rule or self.is_16_moves_rule or self.is_25_moves_rule

    @property
    def is_25_moves_rule(self) -> bool:
        return self.halfmove_clock >= 50

    @property
    def is_16_moves_rule(self) -> bool:
        if self.halfmove_clock < 14:
            return False
        return (_popcount(self.white_men | self.black_men) == 0
                and _popcount(self.white_kings | self.black_kings) == 3)

    @property
    def is_5_moves_rule(self) -> bool:
        if _popcount(self.white_men | self.black_men) == 0 and _popcount(self.white_kings | self.black_kings) == 2:
            if _popcount(self.white_kings) == 1 and _popcount(self.black_kings) == 1:
                return self.halfmove_clock >= 4
        return False


class FryskBoard(FrisianBoard):
    """Frisian rules, five men a side (frysk.py:16-37)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 40, "Frysk!", "frysk"
    STARTING_POSITION = np.array([1] * 5 + [0] * 40 + [-1] * 5, dtype=np.int8)
    __slots__ = ()


class AmericanBoard(_Board32):
    """American checkers (american.py:59-256)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 23, "American checkers", "american"
    __slots__ = ()

    @property
    def is_draw(self) -> bool:
        return self.halfmove_clock >= 40 or self.is_threefold_repetition


class RussianBoard(_Board32):
    """Russian draughts (russian.py:82-379)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 25, "Russian draughts", "russian"
    __slots__ = ()

    @property
    def is_draw(self) -> bool:
        return self.is_threefold_repetition or self.is_15_moves_rule or self.is_3_kings_vs_1_rule

    @property
    def is_15_moves_rule(self) -> bool:
        return self.halfmove_clock >= 30

    @property
    def is_3_kings_vs_1_rule(self) -> bool:
        if self.white_men or self.black_men:
            return False
        wk, bk = _popcount(self.white_kings), _popcount(self.black_kings)
        if (wk >= 3 and bk == 1) or (bk >= 3 and wk == 1):
            return self.halfmove_clock >= 30
        return False


class BrazilianBoard(RussianBoard):
    """Brazilian draughts (brazilian.py:18-77)."""
    GAME_TYPE, VARIANT_NAME, VARIANT = 26, "Brazilian draughts", "brazilian"
    __slots__ = ()


BOARDS = {"standard": StandardBoard, "american": AmericanBoard, "frisian": FrisianBoard, "russian": RussianBoard,
          "brazilian": BrazilianBoard, "antidraughts": AntidraughtsBoard, "breakthrough": BreakthroughBoard,
          "frysk": FryskBoard}
assert list(BOARDS) == _lib.VARIANTS and all(VARIANT_ID[k] == i for i, k in enumerate(BOARDS))


def get_board(variant: str, fen: str | None = None) -> BaseBoard:
    cls = BOARDS[variant]
    return cls.from_fen(fen) if fen else cls()
