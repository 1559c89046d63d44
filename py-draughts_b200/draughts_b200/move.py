"""Move value type with the reference's public surface (draughts/move.py:8-184): visited squares,
captured squares + piece codes, promotion flag, saved halfmove clock, UCI printing / parsing and the
reference's loose equality.  On the device a move is the packed 32-byte DrgMove record
(include/drg.h); `Move.from_record` turns one back into this view."""
from __future__ import annotations

from typing import Iterable

from . import geometry


class Move:
    __slots__ = ("square_list", "captured_list", "captured_entities", "is_promotion", "halfmove_clock", "_len",
                 "_value", "_is_king_move")
    _EMPTY_LIST: list = []

    def __init__(self, visited_squares, captured_list=None, captured_entities=None, is_promotion: bool = False):
        self.square_list = visited_squares
        self.captured_list = captured_list if captured_list else Move._EMPTY_LIST
        self.captured_entities = captured_entities if captured_entities else Move._EMPTY_LIST
        self.is_promotion = is_promotion
        self.halfmove_clock = 0
        self._len = len(captured_list) + 1 if captured_list else 1
        self._value = 0
        self._is_king_move = False

    # -- device record -> reference view ---------------------------------------------------------
    @classmethod
    def from_record(cls, rec, cells, S: int) -> "Move":
        """rec: one MOVE_DTYPE element; cells: callable sq -> cell code of the position the move was
        generated on (base.py:152-163 priority).  Victim k is the captured square lying between
        path[k] and path[k+1] that no earlier step took."""
        n = int(rec["n_sq"])
        path = [int(x) for x in rec["path"][:n]]
        cap_mask = int(rec["captured"])
        caps: list[int] = []
        if cap_mask:
            for a, b in zip(path, path[1:]):
                for sq in geometry.between(a, b, S):
                    if (cap_mask >> sq) & 1 and sq not in caps:
                        caps.append(sq)
                        break
        ents = [cells(sq) for sq in caps]
        m = cls(path, caps, ents, bool(rec["flags"] & 0x01))
        m._is_king_move = bool(rec["flags"] & 0x02)
        m._value = sum(199 if abs(e) == 2 else 100 for e in ents)  # frisian.py:35-36
        return m

    @classmethod
    def list_from_records(cls, recs, cells, S: int) -> list["Move"]:
        """from_record over an array of MOVE_DTYPE records, with the fields pulled out of numpy once (a quiet move
        costs one list slice and one constructor call)."""
        if len(recs) == 0:
            return []
        ns, flags = recs["n_sq"].tolist(), recs["flags"].tolist()
        caps_all, paths = recs["captured"].tolist(), recs["path"].tolist()
        out = []
        for n, fl, cap_mask, path in zip(ns, flags, caps_all, paths):
            path = path[:n]
            if cap_mask:
                caps: list[int] = []
                for a, b in zip(path, path[1:]):
                    for sq in geometry.between(a, b, S):
                        if (cap_mask >> sq) & 1 and sq not in caps:
                            caps.append(sq)
                            break
                ents = [cells(sq) for sq in caps]
                m = cls(path, caps, ents, bool(fl & 0x01))
                m._value = sum(199 if abs(e) == 2 else 100 for e in ents)  # frisian.py:35-36
            else:
                m = cls(path, None, None, bool(fl & 0x01))
            m._is_king_move = bool(fl & 0x02)
            out.append(m)
        return out

    # -- reference surface -----------------------------------------------------------------------
    def __str__(self) -> str:
        if not self.captured_list:
            return f"{self.square_list[0] + 1}-{self.square_list[-1] + 1}"
        return "x".join(str(sq + 1) for sq in self.square_list)

    def __repr__(self) -> str:
        return "Move: " + "->".join(str(s + 1) for s in self.square_list)

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Move):
            return False
        a, b = self.square_list, other.square_list
        if a[0] != b[0] or a[-1] != b[-1]:
            return False
        longer, shorter = (a, b) if len(a) >= len(b) else (b, a)
        return all(sq in longer for sq in shorter)

    def __len__(self) -> int:
        return self._len

    def __add__(self, other: "Move") -> "Move":
        if self.square_list[-1] != other.square_list[0]:
            raise ValueError(f"Cannot append moves {self} and {other}. "
                             "Last square of first move should equal first square of second move.")
        caps = self.captured_list + other.captured_list
        m = Move(self.square_list + other.square_list[1:], caps, self.captured_entities + other.captured_entities,
                 self.is_promotion or other.is_promotion)
        m._len = len(caps) + 1
        return m

    def __hash__(self) -> int:
        return hash((self.square_list[0], self.square_list[-1]))

    @classmethod
    def from_uci(cls, move: str, legal_moves: Iterable["Move"]) -> "Move":
        move = move.lower()
        if "-" in move:
            steps = move.split("-")
        elif "x" in move:
            steps = move.split("x")
        else:
            raise ValueError(f"Invalid move format: {move}")
        probe = Move([int(s) - 1 for s in steps])
        legal = list(legal_moves)
        matches = [m for m in legal if m == probe]
        if len(matches) == 1:
            return matches[0]
        if not matches:
            raise ValueError(f"{probe} is correct format, but not legal in this position.\n"
                             f"Legal moves: {list(map(str, legal))}")
        exact = [m for m in matches if m.square_list == probe.square_list]
        if len(exact) == 1:
            return exact[0]
        raise ValueError(f"{move} is ambiguous: it matches multiple capture paths {list(map(str, matches))}. "
                         "Specify the full path including the intermediate squares.")
