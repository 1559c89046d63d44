"""Host-side board geometry (coordinates only), used to turn packed move records back into the
reference's Move fields.  The kernels have their own tables (csrc/drg_geom.h)."""
from __future__ import annotations

from functools import lru_cache


def coords(sq: int, S: int) -> tuple[int, int]:
    """(row, board column) of a 0-based dark-square index; row 0 is black's back rank."""
    w = 5 if S == 50 else 4
    row, c = divmod(sq, w)
    return row, 2 * c + (0 if row & 1 else 1)


def square_at(row: int, bcol: int, S: int) -> int:
    w = 5 if S == 50 else 4
    return row * w + bcol // 2


@lru_cache(maxsize=None)
def between(a: int, b: int, S: int) -> tuple[int, ...]:
    """Dark squares strictly between a and b on their common diagonal / orthogonal line."""
    ra, ca = coords(a, S)
    rb, cb = coords(b, S)
    dr, dc = rb - ra, cb - ca
    if dr == 0 and dc == 0:
        return ()
    if abs(dr) == abs(dc):
        n, sr, sc = abs(dr), (dr > 0) - (dr < 0), (dc > 0) - (dc < 0)
    elif dr == 0 or dc == 0:  # orthogonal lines only touch every second cell
        n = max(abs(dr), abs(dc)) // 2
        sr, sc = 2 * ((dr > 0) - (dr < 0)), 2 * ((dc > 0) - (dc < 0))
    else:
        return ()
    return tuple(square_at(ra + sr * k, ca + sc * k, S) for k in range(1, n))
