"""API-compatibility gate for the drop-in classes: the reference's test areas -- generic make/unmake,
per-variant movegen, invariants under random play, UCI round trips, PDN replay of real games, ML exports, game end --
exercised on draughts_b200.Board & co, whose legal_moves / push / mask / tensor run on the GPU through the C-ABI.
Expected values come from tests/golden (written by the unmodified reference) or from the oracle."""
import copy
import json
import random

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
VARIANTS = O.VARIANTS


@pytest.fixture(scope="module")
def d():
    import draughts_b200
    from draughts_b200 import _lib
    _lib.lib()
    return draughts_b200


def play(d, variant, seed, plies):
    """test/_test_helpers.py:44-63 on the shim"""
    b = d.BOARDS[variant]()
    rng = random.Random(seed)
    for _ in range(plies):
        lm = list(b.legal_moves)
        if not lm:
            break
        b.push(rng.choice(lm))
    return b


def test_exports_and_class_attributes(d):
    assert d.Board is d.StandardBoard and d.Color.WHITE.value == -1 and d.Figure.BLACK_KING == 2
    for name, cls in d.BOARDS.items():
        b = cls()
        S = cls.SQUARES_COUNT
        assert S == O.SQUARES[name] and b.shape == ((10, 10) if S == 50 else (8, 8))
        assert b.turn == d.Color.WHITE and b.halfmove_clock == 0 and b._moves_stack == []
        assert cls.STARTING_POSITION.dtype == np.int8 and len(cls.STARTING_POSITION) == S
        assert np.array_equal(b.position, cls.STARTING_POSITION)
        wm, wk, bm, bk, _ = O.start_position(name)
        assert (b.white_men, b.white_kings, b.black_men, b.black_kings) == (wm, wk, bm, bk)
        assert cls(cls.STARTING_POSITION).fen == b.fen
    assert d.StandardBoard.GAME_TYPE == 20 and d.AmericanBoard.GAME_TYPE == 23 and d.RussianBoard.GAME_TYPE == 25
    assert d.BrazilianBoard.GAME_TYPE == 26 and d.FrisianBoard.GAME_TYPE == 40 and d.FryskBoard.VARIANT_NAME == "Frysk!"
    assert d.AmericanBoard.SQUARE_NAMES[0] == "b8" and d.AmericanBoard.PROMO_WHITE == 0xF


def test_generic_push_pop_hand_built_moves(d):
    """test_board.py:21-68 pattern: hand-built Move objects on an American board"""
    b = d.AmericanBoard()
    m = d.Move([20, 16])
    b.push(m)
    assert b[16] == -1 and b[20] == 0 and b.turn == d.Color.BLACK
    assert b.pop() is m and b[20] == -1 and b[16] == 0 and b.turn == d.Color.WHITE
    b2 = d.AmericanBoard.from_fen("W:W18:B14")
    cap = d.Move([17, 10], [13], [1])
    b2.push(cap)
    assert b2[10] == -1 and b2[13] == 0 and b2[17] == 0
    b2.pop()
    assert b2[17] == -1 and b2[13] == 1 and b2.fen == '[FEN "W:W18:B14"]'
    with pytest.raises(ValueError):
        d.AmericanBoard().push(d.Move([0, 4]))      # black piece, white to move
    with pytest.raises(ValueError):
        d.AmericanBoard().push(d.Move([15, 11]))    # empty square
    with pytest.raises(IndexError):
        d.AmericanBoard().pop()
    with pytest.raises(ValueError):
        d.Board().push_uci("31-22")
    with pytest.raises(ValueError):
        d.Board.from_fen("not a fen")


@pytest.mark.parametrize("variant", VARIANTS)
def test_invariants_under_random_play(d, variant):
    """test_board_invariants.py:25-80: push/pop restores turn, clock, position, stack length and FEN exactly; men and
    kings never overlap; the shim's legal_moves equal the oracle's at every ply."""
    for seed in range(3):
        b = play(d, variant, seed, 25)
        for _ in range(10):
            lm = b.legal_moves
            want = [O.move_str(m) for m in O.legal_moves(variant, b.white_men, b.white_kings, b.black_men, b.black_kings,
                                                         0 if b.turn == d.Color.WHITE else 1)]
            assert [str(m) for m in lm] == want
            if not lm:
                break
            before = (b.turn, b.halfmove_clock, b.position.tolist(), len(b._moves_stack), b.fen)
            for m in lm:
                b.push(m)
                assert (b.white_men | b.black_men) & (b.white_kings | b.black_kings) == 0
                assert b.pop() is m
                assert (b.turn, b.halfmove_clock, b.position.tolist(), len(b._moves_stack), b.fen) == before
            b.push(lm[seed % len(lm)])


@pytest.mark.parametrize("variant", ["standard", "american", "russian", "frisian"])
def test_uci_round_trip(d, variant):
    """test_move.py:7-18"""
    for seed in range(3):
        b = play(d, variant, 100 + seed, 20)
        for m in b.legal_moves:
            same = [x for x in b.legal_moves if str(x) == str(m)]
            if len(same) == 1:
                assert d.Move.from_uci(str(m), b.legal_moves).square_list == m.square_list


@pytest.mark.parametrize("variant", VARIANTS)
def test_pdn_replay_of_real_games(d, golden_dir, variant):
    """test_pdn.py:25-46 + exact parity: every recorded human game is replayable, and the shim parses each PDN into the
    same move sequence / final FEN / result / clock / re-emitted PDN as the reference."""
    g = json.loads((golden_dir / "pdn_games.json").read_text())["games"][variant]
    cls = d.BOARDS[variant]
    for game in g:
        b = cls.from_pdn(game["pdn"])
        assert [str(m) for m in b._moves_stack] == game["moves"]
        assert (b.fen, b.result, b.halfmove_clock) == (game["fen"], game["result"], game["clock"])
        assert b.pdn == game["pdn_out"]
        # unwind the whole game
        while b._moves_stack:
            b.pop()
        assert b.fen == cls().fen


def test_ai_features(d):
    """test_ai_features.py: copy semantics, tensor / mask shapes and dtypes, index round trip"""
    b = d.Board()
    c = b.copy()
    c.push_uci("31-27")
    assert len(b._moves_stack) == 0 and len(c._moves_stack) == 1 and b.fen != c.fen
    assert copy.copy(b).fen == b.fen and copy.deepcopy(c)._moves_stack[0].square_list == [30, 26]
    t = b.to_tensor()
    assert t.shape == (4, 50) and t.dtype == np.float32 and t.sum() == 40 and t[1].sum() == 0
    tb = b.to_tensor(d.Color.BLACK)
    assert np.array_equal(tb[0], t[2]) and np.array_equal(tb[2], t[0])
    a = d.AmericanBoard()
    assert a.to_tensor().shape == (4, 32) and a.legal_moves_mask().shape == (1024,)
    mask = b.legal_moves_mask()
    assert mask.dtype == np.bool_ and mask.shape == (2500,) and mask.sum() == len(b.legal_moves) == 9
    for m in b.legal_moves:
        idx = b.move_to_index(m)
        assert mask[idx] and b.index_to_move(idx) == m
    with pytest.raises(ValueError):
        b.index_to_move(0)
    f = b.features()
    assert (f.white_men, f.black_men, f.mobility, f.phase, f.turn) == (20, 20, 9, "opening", 1)
    # duplicate routes collapse in the mask and index_to_move returns the first route
    q = d.Board.from_fen("W:WK4:B9,22,32,43,20")
    lm = q.legal_moves
    assert q.legal_moves_mask().sum() < len(lm)
    assert str(q.index_to_move(q.move_to_index(lm[-1]))) in [str(m) for m in lm]


def test_game_over_and_results(d):
    b = d.Board.from_fen("W:W31:B")           # black has nothing
    b.turn = d.Color.BLACK
    assert b.game_over and b.result == "1-0"
    a = d.AntidraughtsBoard.from_fen("B:W31:B")
    assert a.result == "0-1"                   # antidraughts.py:31-37: the side that cannot move wins
    k = d.BreakthroughBoard.from_fen("W:WK1,31:B20")
    assert k.game_over and k.result == "1-0"   # breakthrough.py:28-39
    r = d.RussianBoard.from_fen("W:WK1:BK32")
    r.halfmove_clock = 30
    assert r.is_draw and r.result == "1/2-1/2"
    s = d.Board()
    assert not s.game_over and s.result == "-"
    # seeded self-play terminates (test_frysk_board.py:34-47)
    f = d.FryskBoard()
    rng = random.Random(1)
    for _ in range(400):
        if f.game_over:
            break
        f.push(rng.choice(f.legal_moves))
    assert f.game_over and f.result in ("1-0", "0-1", "1/2-1/2")


def test_russian_mid_capture_promotion(d):
    """test_russian_board.py:179-247"""
    b = d.RussianBoard.from_fen("W:W10:B6,K9,K15,K17,K23")
    lm = b.legal_moves
    assert all(m.is_promotion for m in lm if 1 - 1 in m.square_list[1:] or 0 in m.square_list[1:])
    m = next(m for m in lm if str(m) == "10x1x19x26x13x6")
    assert m.is_promotion and m.captured_list == [5, 14, 22, 16, 8]
    b.push(m)
    assert b.fen == '[FEN "B:WK6:B6"]'
    b.pop()
    assert b.fen == '[FEN "W:W10:B6,K9,K15,K17,K23"]'
    p = d.RussianBoard.from_fen("W:W9:B5,6")
    assert [str(x) for x in p.legal_moves] == [O.move_str(x) for x in O.legal_moves("russian", *O.parse_fen("W:W9:B5,6", 32))]
