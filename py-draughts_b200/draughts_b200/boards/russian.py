"""draughts/boards/russian.py:82: `Board` of this variant."""
from ..models import Color, Figure  # noqa: F401
from ..move import Move  # noqa: F401
from ._impl import BaseBoard  # noqa: F401
from ._impl import RussianBoard as Board  # noqa: F401

# square-name constants (0-based square indices), as the reference's 8x8 modules define them
SQUARES = [B8, D8, F8, H8, A7, C7, E7, G7, B6, D6, F6, H6, A5, C5, E5, G5,
           B4, D4, F4, H4, A3, C3, E3, G3, B2, D2, F2, H2, A1, C1, E1, G1] = range(32)
