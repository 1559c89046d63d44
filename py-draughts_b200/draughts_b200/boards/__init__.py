"""Board classes with the reference's module layout (draughts/boards/__init__.py:3-11): `draughts_b200.boards.base`,
`.standard`, `.american`, `.frisian`, `.russian`, `.brazilian`, `.antidraughts`, `.breakthrough`, `.frysk`, each with
its `Board` class (and the square-name constants of the 8x8 modules).  The implementation lives in `_impl.py`."""
from ._impl import BOARDS, BoardFeatures, format_fen, get_board, parse_fen  # noqa: F401
from .american import Board as AmericanBoard
from .antidraughts import Board as AntidraughtsBoard
from .base import BaseBoard
from .brazilian import Board as BrazilianBoard
from .breakthrough import Board as BreakthroughBoard
from .frisian import Board as FrisianBoard
from .frysk import Board as FryskBoard
from .russian import Board as RussianBoard
from .standard import Board as StandardBoard

__all__ = ["BaseBoard", "StandardBoard", "AmericanBoard", "FrisianBoard", "RussianBoard", "BrazilianBoard",
           "AntidraughtsBoard", "BreakthroughBoard", "FryskBoard"]
