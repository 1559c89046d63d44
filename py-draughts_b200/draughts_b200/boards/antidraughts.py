"""draughts/boards/antidraughts.py:19: `Board` of this variant."""
from ..models import Color, Figure  # noqa: F401
from ..move import Move  # noqa: F401
from ._impl import BaseBoard  # noqa: F401
from ._impl import AntidraughtsBoard as Board  # noqa: F401
