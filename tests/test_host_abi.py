"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol that
include/drg.h declares (no compute call is made -- there is no GPU here), it refuses to compute without a
device, the kernels' geometry table equals the reference's tables, and the pure-host helpers (FEN text, Move
value type, record decoding) behave like the reference's."""
import ctypes as C
import json
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def L():
    from draughts_b200 import _lib
    _lib.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(L):
    from draughts_b200 import _lib
    header = (ROOT / "include" / "drg.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(drg_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/drg.h but not exported"
    assert sorted(declared) == _lib.EXPORTED_SYMBOLS, "ctypes signature table out of sync with include/drg.h"
    assert L.drg_abi_version() == 3
    assert [L.drg_squares(v) for v in range(8)] == [50, 32, 50, 32, 32, 50, 50, 50]
    assert L.drg_squares(8) < 0 and b"variant" in L.drg_last_error()


def test_record_layout_matches_header():
    from draughts_b200._lib import MOVE_DTYPE
    assert MOVE_DTYPE.itemsize == 32 and MOVE_DTYPE == O.MOVE_DTYPE
    assert MOVE_DTYPE.fields["captured"][1] == 0 and MOVE_DTYPE.fields["n_sq"][1] == 8
    assert MOVE_DTYPE.fields["flags"][1] == 9 and MOVE_DTYPE.fields["path"][1] == 10


def test_no_cpu_fallback(L):
    """without a CUDA device every compute entry point fails loudly (DRG_E_CUDA), nothing is computed on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import draughts_b200 as d
    from draughts_b200._lib import DrgError
    with pytest.raises(DrgError) as e:
        d.Board().legal_moves
    assert e.value.code == -2
    counts = np.zeros(1, np.uint32)
    z = np.zeros(1, np.uint64)
    rc = L.drg_legal_moves_host(0, 1, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data,
                                np.zeros(1, np.uint8).ctypes.data, counts.ctypes.data, None, None, 0, None, None, None)
    assert rc == -2


def test_geometry_table_equals_reference(L, golden_dir):
    """NB[sq][d] reproduces MOVE_TGT / JUMP_TGT / KING_RAYS / ORTHO_RAYS / ORTHO_JUMP of the reference."""
    g = json.loads((golden_dir / "tables.json").read_text())
    for name, S in (("standard", 50), ("frisian", 50), ("american", 32), ("russian", 32)):
        nb = np.zeros(512, np.uint8)
        assert L.drg_geometry_table(S, nb.ctypes.data) == 0
        nb = nb.reshape(64, 8).astype(int)
        nb[nb == 255] = -1
        assert nb[:S, :4].tolist() == [list(x) for x in g[name]["MOVE_TGT"]]
        jump = [[(nb[nb[s, d], d] if nb[s, d] >= 0 else -1) for d in range(4)] for s in range(S)]
        assert jump == [list(x) for x in g[name]["JUMP_TGT"]]
        def chain(s, d):
            out, t = [], nb[s, d]
            while t >= 0:
                out.append(int(t))
                t = nb[t, d]
            return out
        if "KING_RAYS" in g[name]:
            assert [[chain(s, d) for d in range(4)] for s in range(S)] == [[list(r) for r in sq] for sq in g[name]["KING_RAYS"]]
        if "ORTHO_RAYS" in g[name]:
            assert [[chain(s, 4 + d) for d in range(4)] for s in range(S)] == [[list(r) for r in sq] for sq in g[name]["ORTHO_RAYS"]]
            for s in range(S):
                tg, ld = g[name]["ORTHO_JUMP"][s]
                for d in range(4):
                    t = nb[s, 4 + d]
                    l2 = nb[t, 4 + d] if t >= 0 else -1
                    assert ((int(t), int(l2)) if (t >= 0 and l2 >= 0) else (-1, -1)) == (tg[d], ld[d])
        assert (nb[S:] == -1).all()
    assert L.drg_geometry_table(40, np.zeros(512, np.uint8).ctypes.data) < 0


def test_fen_text_roundtrip(golden_dir):
    from draughts_b200.boards import format_fen, parse_fen, StandardBoard
    fx = json.loads((golden_dir / "ref_fixtures.json").read_text())
    for rec in fx["random_positions"]:
        assert parse_fen(rec["src_fen"], 50) == O.parse_fen(rec["src_fen"], 50)
        assert format_fen(*parse_fen(rec["src_fen"], 50), 50)[6:-2] == rec["fen"]
    # test_standard_board.py:36-74: legacy prefix, empty side, wrapper, tails
    assert parse_fen("W:B:W31:B1", 50) == parse_fen("B:W31:B1", 50)
    assert parse_fen("B:W:B1", 50) == (0, 0, 1, 0, 1)
    assert parse_fen('[FEN "W:WK4:B1:H0:F2"]', 50) == (0, 8, 1, 0, 0)
    with pytest.raises(ValueError):
        parse_fen("garbage", 50)
    b = StandardBoard.from_fen("W:WK4,31:B1,K2")
    assert b.fen == '[FEN "W:W31,K4:B1,K2"]' or b.fen == '[FEN "W:WK4,31:B1,K2"]'
    assert b.position.tolist()[:4] == [1, 2, 0, -2] and b.position.dtype == np.int8
    assert StandardBoard().fen == format_fen(*O.start_position("standard")[:4], 0, 50)


def test_move_value_type_and_record_decoding(golden_dir):
    """Move.from_record on records produced by the oracle reproduces the reference's Move fields (order of victims!)."""
    from draughts_b200 import Move, BOARDS
    g = json.loads((golden_dir / "movelists_russian.json").read_text())
    checked = 0
    for rec in g["positions"]:
        if "moves" not in rec or not any(len(m[1]) >= 3 for m in rec["moves"]):
            continue
        wm, wk, bm, bk, turn = O.parse_fen(rec["fen"], 32)
        out = O.batch_emit("russian", [wm], [wk], [bm], [bk], [turn])
        b = BOARDS["russian"].from_fen(rec["fen"])
        got = []
        for r in out["moves"]:
            m = Move.from_record(r, b._get, 32)
            got.append([str(m), list(m.captured_list), list(m.captured_entities), bool(m.is_promotion)])
        assert got == rec["moves"], rec["fen"]
        checked += 1
    assert checked > 20
    a, b2 = Move([30, 26]), Move([30, 26])
    assert a == b2 and hash(a) == hash(b2) and str(a) == "31-27" and repr(a) == "Move: 31->27" and len(a) == 1
    c = Move([3, 26], [13], [1]) + Move([26, 37], [31], [1])
    assert str(c) == "4x27x38" and c.captured_list == [13, 31] and len(c) == 3
    with pytest.raises(ValueError):
        Move([1, 2]) + Move([3, 4])
    legal = [Move([3, 26, 37, 14], [1, 2, 3], [1, 1, 1]), Move([3, 30, 41, 14], [4, 5, 6], [1, 1, 1])]
    with pytest.raises(ValueError, match="ambiguous"):
        Move.from_uci("4x15", legal)  # test_standard_board.py:145-161
    assert Move.from_uci("4x31x42x15", legal) is legal[1]
    with pytest.raises(ValueError):
        Move.from_uci("1-2", legal)
    with pytest.raises(ValueError):
        Move.from_uci("nonsense", legal)


def test_pop_and_terminal_predicates_host_logic():
    """pop() bookkeeping and the draw predicates are host code: exercise them without the GPU."""
    from draughts_b200 import AmericanBoard, Color, FrisianBoard, Move, RussianBoard, StandardBoard
    b = StandardBoard.from_fen("W:W31:B1")
    m = Move([30, 26])
    m.halfmove_clock = 7
    b._moves_stack.append(m)                     # state as if push(31-27) had happened
    b.white_men, b.turn = 1 << 26, Color.BLACK
    assert b.pop() is m and b.white_men == 1 << 30 and b.turn == Color.WHITE and b.halfmove_clock == 7
    with pytest.raises(IndexError):
        b.pop()
    s = StandardBoard.from_fen("W:WK1,K2:BK50")    # 3 pieces, kings*2 + men = 6 >= 5 (standard.py:302-311)
    s.halfmove_clock = 9
    assert not s.is_draw
    s.halfmove_clock = 10
    assert s.is_5_moves_rule and s.is_draw
    f = FrisianBoard.from_fen("W:WK1:BK50")
    f.halfmove_clock = 4
    assert f.is_draw
    a = AmericanBoard()
    a.halfmove_clock = 40
    assert a.is_draw
    r = RussianBoard.from_fen("W:WK1,K2,K3:BK32")
    r.halfmove_clock = 30
    assert r.is_3_kings_vs_1_rule and r.is_draw
    mv = [Move([1, 2]), Move([3, 4]), Move([5, 6]), Move([7, 8])]
    r2 = RussianBoard()
    r2._moves_stack = [mv[i % 4] for i in range(9)]
    assert r2.is_threefold_repetition


def test_bulk_fen_codec_matches_reference_strings(golden_dir):
    """drg_fen_decode / drg_fen_encode (host code, no device): every FEN the reference wrote into the golden files
    decodes to the bitboards the regex restatement gives and encodes back to the reference's exact `fen` string
    (base.py:408-495); non-canonical spellings are handed to the full parser, malformed text raises ValueError."""
    from draughts_b200 import BoardBatch
    from draughts_b200.boards import format_fen, parse_fen
    from draughts_b200._lib import VARIANTS, squares
    for variant in VARIANTS:
        g = json.loads((golden_dir / f"movelists_{variant}.json").read_text())
        S = squares(variant)
        fens = []
        for rec in g["positions"]:
            fens.append(rec["fen"])
            fens.extend(a for a in (rec.get("after") or [])[:4] if isinstance(a, str))
        fens += ['[FEN "%s"]' % f for f in fens[:50]]
        b = BoardBatch.from_fens(variant, fens)
        want = [parse_fen(f, S) for f in fens]
        assert [tuple(int(x[i]) for x in (b.wm, b.wk, b.bm, b.bk, b.turn)) for i in range(len(fens))] == want
        assert b.fens() == [format_fen(*w, S) for w in want]
    # the reference's own tolerated spellings (test_standard_board.py:36-74) take the full parser
    odd = ["W:B:W31:B1", "B:W:B1", '[FEN "W:WK4:B1:H0:F2"]', "w:wk4,31:b1,k2", "W:W31,G5,P7:B1", "W:W31,,32,:B,1", "W:W007:BK050",
           "W:W5,K5:BK5", "B:W:B"]
    b = BoardBatch.from_fens("standard", odd)
    assert [tuple(int(x[i]) for x in (b.wm, b.wk, b.bm, b.bk, b.turn)) for i in range(len(odd))] == [parse_fen(f, 50) for f in odd]
    for bad in ["garbage", "W:W31", "W:WK:B1", "W:W51:B1"]:
        with pytest.raises((ValueError, IndexError)):
            BoardBatch.from_fens("standard", ["W:W31:B1", bad])
    # overlapping colours (Russian Q4) and a man shadowing a king of its own colour print like BaseBoard.fen
    q = BoardBatch("russian", np.array([1 << 5 | 1 << 9], np.uint64), np.array([1 << 5], np.uint64), np.array([1 << 5], np.uint64),
                   np.array([0], np.uint64), np.array([1], np.uint8))
    assert q.fens() == ['[FEN "B:W6,10:B6"]'] == [format_fen(1 << 5 | 1 << 9, 1 << 5, 1 << 5, 0, 1, 32)]
    assert BoardBatch.from_fens("standard", []).fens() == []
    # large ragged batch: encode -> decode is the identity
    rng = np.random.default_rng(3)
    occ = rng.integers(0, 1 << 50, 20000, dtype=np.uint64) & rng.integers(0, 1 << 50, 20000, dtype=np.uint64)
    a, c, e = (rng.integers(0, 1 << 50, 20000, dtype=np.uint64) for _ in range(3))
    big = BoardBatch("frisian", occ & a & c, occ & a & ~c, occ & ~a & e, occ & ~a & ~e, rng.integers(0, 2, 20000).astype(np.uint8))
    back = BoardBatch.from_fens("frisian", big.fens())
    for k in ("wm", "wk", "bm", "bk", "turn"):
        assert np.array_equal(getattr(back, k), getattr(big, k))


def test_piece_square_tables_match_reference(golden_dir):
    """drg_pst_table (host code of the library, no device) == AlphaBetaEngine._get_pst_tables of the reference."""
    from draughts_b200 import _lib
    L = _lib.load()
    g = json.loads((golden_dir / "evaluate.json").read_text())
    for variant, rec in g["variants"].items():
        out = np.zeros(4 * rec["S"], np.float64)
        assert L.drg_pst_table(_lib.VARIANT_ID[variant], out.ctypes.data) == 4 * rec["S"]
        want = np.array([float.fromhex(x) for t in rec["pst"] for x in t])
        assert out.tobytes() == want.tobytes(), variant
    assert L.drg_pst_table(99, np.zeros(256).ctypes.data) < 0


def test_new_entry_points_reject_bad_arguments_without_a_device():
    """Argument errors are reported by code before any CUDA call (or, for the device entry points, as 'not
    initialised' on a box without a GPU) -- never a crash, never a silent success."""
    from draughts_b200 import _lib
    L = _lib.load()
    off = np.zeros(2, np.int64)
    one = np.zeros(1, np.uint64)
    t = np.zeros(1, np.uint8)
    # FEN codec is pure host code: full checks run here
    assert L.drg_fen_encode(99, 1, one.ctypes.data, one.ctypes.data, one.ctypes.data, one.ctypes.data, t.ctypes.data, None, 0, off.ctypes.data) < 0
    assert L.drg_fen_encode(0, 1, None, None, None, None, None, None, 0, off.ctypes.data) == _lib.E_ARG
    assert L.drg_fen_encode(0, 1, one.ctypes.data, one.ctypes.data, one.ctypes.data, one.ctypes.data, t.ctypes.data, None, 0, off.ctypes.data) == 0
    assert off.tolist() == [0, len('[FEN "W:W:B"]')]
    buf = np.zeros(4, np.uint8)   # too small for the row
    assert L.drg_fen_encode(0, 1, one.ctypes.data, one.ctypes.data, one.ctypes.data, one.ctypes.data, t.ctypes.data, buf.ctypes.data, 4, off.ctypes.data) == _lib.E_OVERFLOW
    assert L.drg_fen_decode(0, 1, None, None, None, None, None, None, None, None) == _lib.E_ARG
    assert L.drg_fen_decode(0, -1, None, None, None, None, None, None, None, None) == _lib.E_ARG
    assert L.drg_pst_table(0, None) < 0
    import torch
    if not torch.cuda.is_available():
        # device entry points: no context here -> DRG_E_CUDA with a message, not a fallback
        assert L.drg_evaluate(0, 1, one.ctypes.data, one.ctypes.data, one.ctypes.data, one.ctypes.data, t.ctypes.data, one.ctypes.data, None) == _lib.E_CUDA
        assert L.drg_children(0, 1, *([one.ctypes.data] * 4), t.ctypes.data, t.ctypes.data, off.ctypes.data, 1, one.ctypes.data,
                              *([one.ctypes.data] * 4), t.ctypes.data, t.ctypes.data, None, None, None) == _lib.E_CUDA
        assert b"no CPU fallback" in L.drg_last_error() or b"drg_init" in L.drg_last_error()


def test_fen_decoder_never_disagrees_with_the_full_parser():
    """Mutated FEN text: whenever the native decoder accepts a row (status 0) the regex restatement of from_fen must give
    the same bitboards; rows it hands over (status 1) are simply left to that parser.  No input may crash it."""
    from draughts_b200 import _lib
    from draughts_b200.boards import parse_fen
    L = _lib.load()
    rng = np.random.default_rng(12)
    base = ["W:W31,32,K5:B1,2,K50", "B:WK1:BK2,3,4", '[FEN "W:W21,22:B9,10"]', "W:W:B", "B:W5:B", "W:W1,2,3,4,5:B46,47,48,49,50"]
    alphabet = list("WBK0123456789,:[]\"FEN wbkGP-x")
    rows = []
    for _ in range(4000):
        s = list(base[int(rng.integers(len(base)))])
        for _ in range(int(rng.integers(0, 4))):
            op = int(rng.integers(3))
            pos = int(rng.integers(len(s) + 1))
            if op == 0 and s:
                del s[min(pos, len(s) - 1)]
            elif op == 1:
                s.insert(pos, alphabet[int(rng.integers(len(alphabet)))])
            elif s:
                s[min(pos, len(s) - 1)] = alphabet[int(rng.integers(len(alphabet)))]
        rows.append("".join(s))
    raw = [r.encode() for r in rows]
    off = np.zeros(len(raw) + 1, np.int64)
    np.cumsum([len(r) for r in raw], out=off[1:])
    text = np.frombuffer(b"".join(raw) + b"\0", np.uint8)
    n = len(rows)
    wm, wk, bm, bk = (np.zeros(n, np.uint64) for _ in range(4))
    turn, status = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    for variant, S in (("standard", 50), ("american", 32)):
        assert L.drg_fen_decode(_lib.VARIANT_ID[variant], n, text.ctypes.data, off.ctypes.data, wm.ctypes.data, wk.ctypes.data,
                                bm.ctypes.data, bk.ctypes.data, turn.ctypes.data, status.ctypes.data) == 0
        accepted = 0
        for i in np.flatnonzero(status == 0):
            assert parse_fen(rows[i], S) == (int(wm[i]), int(wk[i]), int(bm[i]), int(bk[i]), int(turn[i])), rows[i]
            accepted += 1
        assert accepted > 500   # the mutations leave plenty of canonical rows


def test_result_pool_recycles_by_lifetime_and_never_aliases_live_arrays():
    """_lib._PinnedPool (the result buffers of BoardBatch.legal_moves): plain numpy memory for a size class's first
    request and below 1 MiB, pool blocks from the second on, a block comes back only when the last view of its array
    is gone, and the pool stops pinning at its limit (exercised with a malloc stand-in for cudaMallocHost)."""
    import ctypes
    import gc
    from draughts_b200 import _lib
    libc = ctypes.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = ctypes.c_void_p, [ctypes.c_size_t]
    allocs = []

    def alloc(nbytes):
        allocs.append(nbytes)
        return libc.malloc(nbytes)

    pool = _lib._PinnedPool(alloc=alloc)
    pool.limit = (9 << 20) // 2
    addr = lambda a: a.__array_interface__["data"][0]
    assert pool.empty(100, np.uint32).nbytes == 400 and not allocs          # small: numpy memory
    a = pool.empty((1 << 18, 5), np.uint8)                                    # 1.25 MiB, first of its class: numpy memory
    assert not allocs and a.shape == (1 << 18, 5)
    b = pool.empty((1 << 18, 5), np.uint8)                                    # second: a pool block
    assert len(allocs) == 1 and allocs[0] >= b.nbytes and allocs[0] <= b.nbytes * 1.13
    c = pool.empty((1 << 18, 5), np.uint8)
    assert len(allocs) == 2 and addr(b) != addr(c)                            # live arrays never share memory
    view = b[7:9, 1:3]
    pb = addr(b)
    del b
    gc.collect()
    assert addr(pool.empty((1 << 18, 5), np.uint8)) != pb and len(allocs) == 3  # a view of b is alive: b's block stays out
    del view
    gc.collect()
    d = pool.empty(5 << 18, np.uint8)
    assert addr(d) == pb and len(allocs) == 3                                 # ... and comes back with the last view
    d[:] = 7
    assert pool.held == 3 * allocs[0] and pool.held + allocs[0] > pool.limit
    e = pool.empty(5 << 18, np.uint8)                                         # limit reached: numpy memory again
    assert len(allocs) == 3 and e.nbytes == 5 << 18
    s = _lib._PinnedPool.size_class
    assert all(n <= s(n) <= n * 1.126 + (1 << 16) for n in (1 << 20, 3_000_001, 2_621_440_000, 33_554_437))
