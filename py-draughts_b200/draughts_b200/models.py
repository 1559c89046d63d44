"""Colour / piece codes of the reference API (draughts/models.py:12-39): WHITE = -1, BLACK = 1;
cell codes -2 WK, -1 WM, 0 empty, 1 BM, 2 BK."""
from __future__ import annotations

from enum import Enum, IntEnum


class Color(Enum):
    WHITE = -1
    BLACK = 1


class Figure(IntEnum):
    BLACK_KING = 2
    BLACK_MAN = 1
    WHITE_KING = -2
    WHITE_MAN = -1
    KING = 2
    MAN = 1
    EMPTY = 0


FIGURE_REPR = {Figure.BLACK_MAN: "b", Figure.WHITE_MAN: "w", Figure.EMPTY: ".", Figure.BLACK_KING: "B",
               Figure.WHITE_KING: "W"}
