"""Pins the CPU oracle (oracle/drg_oracle.c) to vectors produced by the unmodified Python reference
(tools/gen_golden.py -> tests/golden).  CPU only."""
import hashlib
import json

import numpy as np
import pytest

from oracle import oracle as O

VARIANTS = O.VARIANTS


def load(golden_dir, name):
    return json.loads((golden_dir / name).read_text())


def oracle_position_record(variant, fen, clock=0):
    """Same shape as tools/gen_golden.py:position_record, computed by the oracle."""
    S = O.SQUARES[variant]
    wm, wk, bm, bk, turn = O.parse_fen(fen, S)
    moves = O.legal_moves(variant, wm, wk, bm, bk, turn)
    entries = [[O.move_str(m), m["cap"], m["ent"], m["promo"]] for m in moves]
    after = []
    for a in O.push_all(variant, wm, wk, bm, bk, turn, clock):
        if a is None:
            after.append("ValueError")
        else:
            after.append(f"{O.make_fen(a[0], a[1], a[2], a[3], a[4], S)[6:-2]}|{a[5]}|{int(a[6])}")
    mask = sorted({m["sq"][0] * S + m["sq"][-1] for m in moves})
    return entries, after, mask


def check_record(variant, rec):
    entries, after, mask = oracle_position_record(variant, rec["fen"], rec.get("clock", 0))
    if "moves" in rec:
        assert entries == [[m[0], m[1], m[2], m[3]] for m in rec["moves"]], rec["fen"]
    else:
        assert len(entries) == rec["n_moves"], rec["fen"]
        sha = hashlib.sha256(json.dumps(entries, separators=(",", ":")).encode()).hexdigest()[:16]
        assert sha == rec["moves_sha"], rec["fen"]
    if "after" in rec:
        assert after == rec["after"], rec["fen"]
    assert hashlib.sha256("\n".join(after).encode()).hexdigest()[:16] == rec["after_sha"], rec["fen"]
    assert mask == rec["mask"], rec["fen"]


def test_tables_match_reference(golden_dir):
    g = load(golden_dir, "tables.json")
    L = O.lib()
    for name, S in (("standard", 50), ("american", 32), ("russian", 32), ("frisian", 50)):
        mt = np.zeros((S, 4), np.int32)
        jt = np.zeros((S, 4), np.int32)
        rl = np.zeros((S, 4), np.int32)
        ray = np.zeros((S, 4, 9), np.int32)
        orl = np.zeros((S, 4), np.int32)
        oray = np.zeros((S, 4, 5), np.int32)
        ojt = np.zeros((S, 4), np.int32)
        ojl = np.zeros((S, 4), np.int32)
        L.orc_tables(S, *[a.ctypes.data for a in (mt, jt, rl, ray, orl, oray, ojt, ojl)])
        assert mt.tolist() == [list(x) for x in g[name]["MOVE_TGT"]]
        assert jt.tolist() == [list(x) for x in g[name]["JUMP_TGT"]]
        if "KING_RAYS" in g[name]:
            got = [[[int(v) for v in ray[s, d, : rl[s, d]]] for d in range(4)] for s in range(S)]
            assert got == [[list(r) for r in sq] for sq in g[name]["KING_RAYS"]]
        if "ORTHO_RAYS" in g[name]:
            got = [[[int(v) for v in oray[s, d, : orl[s, d]]] for d in range(4)] for s in range(S)]
            assert got == [[list(r) for r in sq] for sq in g[name]["ORTHO_RAYS"]]
            assert [[ojt[s].tolist(), ojl[s].tolist()] for s in range(S)] == \
                [[list(t), list(l)] for t, l in g[name]["ORTHO_JUMP"]]


def test_reference_fixture_legal_moves_len(golden_dir):
    """test/games/standard/legal_moves_len.json (test_standard_board.py:76-82)."""
    fx = load(golden_dir, "ref_fixtures.json")
    for fen, n in fx["legal_moves_len"].items():
        wm, wk, bm, bk, turn = O.parse_fen(fen, 50)
        assert len(O.legal_moves("standard", wm, wk, bm, bk, turn)) == n, fen


def test_reference_random_positions_fingerprint(golden_dir):
    """SURVEY 8(c) fingerprint over the 73 FENs of tools/random_positions.json."""
    fx = load(golden_dir, "ref_fixtures.json")
    lines = []
    for rec in fx["random_positions"]:
        wm, wk, bm, bk, turn = O.parse_fen(rec["src_fen"], 50)
        lines.append(rec["src_fen"] + "|" + ",".join(O.move_str(m) for m in O.legal_moves("standard", wm, wk, bm, bk, turn)))
        check_record("standard", rec)
        # FEN round trip (test_standard_board.py:27-34)
        assert O.make_fen(wm, wk, bm, bk, turn, 50)[6:-2] == rec["fen"]
    sha = hashlib.sha256("\n".join(lines).encode()).hexdigest()
    assert sha == fx["random_positions_sha256"] == "70de4bfbcaae7435514cf81f1fcd0b8eafa9a8a651c1250f1d5414cce3175feb"


def test_reference_named_positions(golden_dir):
    fx = load(golden_dir, "ref_fixtures.json")
    for variant, recs in fx["named"].items():
        for rec in recs:
            check_record(variant, rec)
    # exact strings asserted by the reference's own tests
    def strs(v, fen):
        return [O.move_str(m) for m in O.legal_moves(v, *O.parse_fen(fen, O.SQUARES[v]))]
    assert strs("standard", "W:WK43:B8,9,11,14,19,38,39,42,47,48").count("43x30x13x2x16x43") == 1
    q4 = strs("russian", "W:W10:B6,K9,K15,K17,K23")
    assert "10x1x19x26x13x6" in q4  # SURVEY Q4 witness
    assert len(strs("frisian", "W:WK1,K15,K20:BK4,7,9,10,12,17,21,26,29,34,40,41,44")) == 9832


@pytest.mark.parametrize("variant", VARIANTS)
def test_perft_start(golden_dir, variant):
    pf = load(golden_dir, "perft.json")
    s = O.start_position(variant)
    want = pf["start"][variant]
    got = [O.perft(variant, *s, d) for d in range(1, len(want) + 1)]
    assert got == want


def test_perft_fens_and_deep(golden_dir):
    pf = load(golden_dir, "perft.json")
    for e in pf["fen"]:
        v = e["variant"]
        b = O.parse_fen(e["fen"], O.SQUARES[v])
        assert [O.perft(v, *b, d) for d in range(1, len(e["perft"]) + 1)] == e["perft"], e
    # SURVEY 8(c) deep values (reference, multi-process) within CPU-suite budget
    deep = pf["survey_deep"]
    T = O.host_threads()
    for v, d in (("standard", "8"), ("russian", "9"), ("frisian", "8"), ("frysk", "8"), ("american", "9")):
        s = O.start_position(v)
        wm, wk, bm, bk = O.expand(v, *[np.array([x], np.uint64) for x in s[:4]], s[4])
        wm, wk, bm, bk = O.expand(v, wm, wk, bm, bk, 1 - s[4])
        assert O.perft_batch(v, wm, wk, bm, bk, s[4], int(d) - 2, threads=T) == deep[v][d], (v, d)


@pytest.mark.parametrize("variant", VARIANTS)
def test_movelists(golden_dir, variant):
    g = load(golden_dir, f"movelists_{variant}.json")
    assert g["variant"] == variant
    for rec in g["positions"]:
        check_record(variant, rec)


@pytest.mark.parametrize("variant", VARIANTS)
def test_tensor_and_mask_batch(golden_dir, variant):
    """orc_batch_emit's mask / tensor / packed records against the reference dumps."""
    g = load(golden_dir, f"movelists_{variant}.json")
    S = O.SQUARES[variant]
    boards = [O.parse_fen(r["fen"], S) for r in g["positions"]]
    arr = [np.array([b[i] for b in boards], np.uint64) for i in range(4)] + [np.array([b[4] for b in boards], np.uint8)]
    out = O.batch_emit(variant, *arr)
    for i, rec in enumerate(g["positions"]):
        assert np.flatnonzero(out["mask"][i]).tolist() == rec["mask"]
        assert np.flatnonzero(out["tensor"][i].ravel()).tolist() == rec["tensor"]
        n = rec["n_moves"] if "n_moves" in rec else len(rec["moves"])
        assert out["counts"][i] == n
        if "moves" in rec:
            recs = out["moves"][int(out["offsets"][i]): int(out["offsets"][i + 1])]
            for r, m in zip(recs, rec["moves"]):
                path = r["path"][: r["n_sq"]].tolist()
                if m[1]:
                    assert "x".join(str(s + 1) for s in path) == m[0]
                else:
                    assert f"{path[0] + 1}-{path[-1] + 1}" == m[0]
                cap = 0
                for c in m[1]:
                    cap |= 1 << c
                assert int(r["captured"]) == cap
                assert bool(r["flags"] & 1) == m[3]


def test_rollouts(golden_dir):
    g = load(golden_dir, "rollouts.json")
    res_code = {"-": 0, "1-0": 1, "0-1": 2, "1/2-1/2": 3}
    for variant, games in g["games"].items():
        S = O.SQUARES[variant]
        for gm in games:
            r = O.rollout(variant, 1, gm["seed"], gm["game_id"], gm["max_plies"], flags=1 if gm["draw_rules"] else 0)
            fen = O.make_fen(int(r["wm"][0]), int(r["wk"][0]), int(r["bm"][0]), int(r["bk"][0]), int(r["turn"][0]), S)[6:-2]
            assert (int(r["plies"][0]), fen, int(r["clock"][0])) == (gm["plies"], gm["fen"], gm["clock"]), (variant, gm)
            assert int(r["result"][0]) == res_code[gm["result"]], (variant, gm)


def test_rollout_threads_and_game_id_independence():
    a = O.rollout("standard", 64, 5, game_id0=100, threads=1)
    b = O.rollout("standard", 64, 5, game_id0=100, threads=4)
    c = O.rollout("standard", 32, 5, game_id0=132, threads=2)
    for k in a:
        assert np.array_equal(a[k], b[k])
        assert np.array_equal(a[k][32:], c[k])


def test_leaf_evaluation_matches_reference_bit_for_bit(golden_dir):
    """orc_pst / orc_evaluate against AlphaBetaEngine._get_pst_tables / evaluate run by the reference
    (tests/golden/evaluate.json, doubles stored as float.hex): exact equality, signed zeros included."""
    g = json.loads((golden_dir / "evaluate.json").read_text())
    for variant, rec in g["variants"].items():
        S = rec["S"]
        want_pst = np.array([[float.fromhex(x) for x in t] for t in rec["pst"]])
        assert O.pst(S).tobytes() == want_pst.tobytes(), variant
        boards = [O.parse_fen(f, S) for f, _ in rec["positions"]]
        cols = list(zip(*boards))
        got = O.evaluate(variant, *(np.array(c, np.uint64) for c in cols[:4]), np.array(cols[4], np.uint8))
        want = np.array([float.fromhex(h) for _, h in rec["positions"]])
        assert got.tobytes() == want.tobytes(), variant
