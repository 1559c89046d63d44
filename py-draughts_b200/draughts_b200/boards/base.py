"""draughts/boards/base.py: BaseBoard, BoardFeatures and the names it re-exports (Color, Figure, Move)."""
from ..models import Color, Figure  # noqa: F401
from ..move import Move  # noqa: F401
from ._impl import BaseBoard, BoardFeatures  # noqa: F401
