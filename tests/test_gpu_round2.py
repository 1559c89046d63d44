"""GPU parity tests of the round-2 pieces: split capture-tree walks (king-endgame monsters), hash-sharded perft,
the on-device policy-head -> move mapping, guard bands around every wide output, 1 Mi-game rollouts."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drg():
    import draughts_b200
    from draughts_b200 import _lib
    _lib.lib()
    return draughts_b200


def _synthetic(variant, n, seed):
    from fuzz_parity import synthetic
    return synthetic(variant, O.SQUARES[variant], n, np.random.default_rng(seed))


def _fused(drg, variant, arrs, want, guard=0):
    """DeviceBoardBatch.movegen (fused two-stream call) with mask + tensor; optional guard bands around every output"""
    import torch
    n, S = len(arrs[0]), O.SQUARES[variant]
    total = int(want["offsets"][-1])
    db = drg.DeviceBoardBatch.from_host(drg.BoardBatch(variant, *arrs), "cuda")

    def banded(shape, dtype):
        flat = int(np.prod(shape))
        unit = torch.empty(0, dtype=dtype).element_size()
        g = (guard * 32 // unit) if guard else 0  # guard * 32 bytes each side keeps every view 32-byte aligned
        buf = torch.full((flat + 2 * g,), 0x5A if dtype == torch.uint8 else 7, dtype=dtype, device="cuda")
        return buf, buf[g:g + flat].view(shape), g

    bufs = {}
    out = {}
    for key, shape, dtype in (("counts", (n,), torch.int32), ("offsets", (n + 1,), torch.int64),
                              ("moves", (max(total, 1), 32), torch.uint8), ("mask", (n, S * S), torch.uint8),
                              ("tensor", (n, 4, S), torch.float32)):
        bufs[key] = banded(shape, dtype)
        out[key] = bufs[key][1]
    out["mask"] = out["mask"].view(torch.bool)
    out["overflow"] = torch.zeros(1, dtype=torch.int64, device="cuda")
    db.movegen(out)
    torch.cuda.synchronize()
    assert int(out["overflow"].item()) == 0
    assert np.array_equal(out["counts"].cpu().numpy().astype(np.uint32), want["counts"])
    assert np.array_equal(out["offsets"].cpu().numpy().astype(np.uint64), want["offsets"])
    assert out["moves"].cpu().numpy()[:total].tobytes() == want["moves"].tobytes()
    assert np.array_equal(out["mask"].cpu().numpy().view(np.uint8), want["mask"])
    assert out["tensor"].cpu().numpy().tobytes() == want["tensor"].tobytes()
    if guard:
        for key, (buf, _, g) in bufs.items():
            if g == 0:
                continue
            fill = 0x5A if buf.dtype == torch.uint8 else 7
            assert bool((buf[:g] == fill).all()) and bool((buf[-g:] == fill).all()), f"write outside {key}"


@pytest.mark.parametrize("variant", ["frisian", "frysk", "standard", "russian", "brazilian", "american", "antidraughts"])
def test_king_endgame_batches_bit_exact(drg, variant):
    """BASELINE configs[4]: capture-heavy king endgames (trees of up to 10^5 nodes for Frisian): host path and fused
    device path equal the oracle bit for bit -- counts, offsets, every record in canonical order, mask, tensor."""
    arrs = _synthetic(variant, 12000, 100)
    want = O.batch_emit(variant, *arrs, threads=O.host_threads())
    got = drg.BoardBatch(variant, *arrs).legal_moves(want_mask=True, want_tensor=True)
    assert np.array_equal(got["counts"], want["counts"]) and np.array_equal(got["offsets"], want["offsets"])
    assert got["moves"].tobytes() == want["moves"].tobytes()
    assert np.array_equal(got["mask"], want["mask"]) and got["tensor"].tobytes() == want["tensor"].tobytes()
    _fused(drg, variant, arrs, want, guard=64)
    if variant == "frisian":
        assert int(want["counts"].max()) > 1000  # the batch really contains fan-out monsters


@pytest.mark.parametrize("variant,n_monsters", [("frisian", 3), ("frisian", 40), ("frysk", 300), ("standard", 40), ("russian", 40)])
def test_few_monsters_among_ordinary_boards(drg, variant, n_monsters):
    """the regimes of round_budget(): a handful of monster trees in a batch of ordinary positions leaves the lanes of
    the split walk nearly empty (budgets 6 and 12, several frames handed over child by child), a few hundred fill them
    (budgets 32 / 64); whatever the budgets, every record sits where the oracle puts it"""
    base = drg.random_positions(variant, 30_000, seed=5, max_ply=120)
    mon = _synthetic(variant, 4 * n_monsters, 31)
    cnt = O.batch_emit(variant, *mon, want_mask=False, want_tensor=False, threads=O.host_threads())["counts"]
    heavy = np.argsort(cnt)[::-1][:n_monsters]          # the boards with the longest move lists
    arrs = [np.concatenate([getattr(base, k), m[heavy]]) for k, m in zip(("wm", "wk", "bm", "bk", "turn"), mon)]
    perm = np.random.default_rng(3).permutation(len(arrs[0]))
    arrs = [a[perm] for a in arrs]
    want = O.batch_emit(variant, *arrs, threads=O.host_threads())
    _fused(drg, variant, arrs, want, guard=64)
    got = drg.BoardBatch(variant, *arrs).legal_moves(want_mask=True, want_tensor=True)
    assert got["moves"].tobytes() == want["moves"].tobytes() and np.array_equal(got["mask"], want["mask"])


def _many_kings(variant, n, seed):
    """boards where the side to move has 3-9 kings and the other side a few men: quiet-move lists of 30-100 entries,
    mixed with capture boards"""
    S = O.SQUARES[variant]
    rng = np.random.default_rng(seed)
    wm = np.zeros(n, np.uint64); wk = np.zeros(n, np.uint64); bm = np.zeros(n, np.uint64); bk = np.zeros(n, np.uint64)
    for i in range(n):
        sq = rng.choice(S, size=int(rng.integers(5, 16)), replace=False)
        nk = int(rng.integers(3, 10))
        for t, q in enumerate(sq):
            bit = np.uint64(1 << int(q))
            if t < nk: wk[i] |= bit
            elif t % 3 == 0: wm[i] |= bit
            else: bm[i] |= bit
    turn = rng.integers(0, 2, n).astype(np.uint8)
    sw = turn == 1
    wm[sw], bm[sw] = bm[sw].copy(), wm[sw].copy()
    bk[sw], wk[sw] = wk[sw].copy(), bk[sw].copy()
    return wm, wk, bm, bk, turn


@pytest.mark.parametrize("variant", ["standard", "american", "frisian", "russian"])
def test_many_king_boards_with_long_quiet_lists(drg, variant):
    """boards whose side to move has 3-9 kings (30-100 quiet moves, long flying-king rays) mixed with ordinary ones:
    the fused call equals the oracle, also when the record buffer clips in the middle of such a list"""
    import torch
    mk = _many_kings(variant, 6000, 17)
    base = drg.random_positions(variant, 6000, seed=9, max_ply=120)
    arrs = [np.concatenate([getattr(base, k), m]) for k, m in zip(("wm", "wk", "bm", "bk", "turn"), mk)]
    perm = np.random.default_rng(4).permutation(len(arrs[0]))
    arrs = [a[perm] for a in arrs]
    want = O.batch_emit(variant, *arrs, threads=O.host_threads())
    assert variant == "american" or int((want["counts"] > 32).sum()) > 500  # short kings never reach 32 moves
    _fused(drg, variant, arrs, want, guard=64)
    n, S = len(arrs[0]), O.SQUARES[variant]
    total = int(want["offsets"][-1])
    db = drg.DeviceBoardBatch.from_host(drg.BoardBatch(variant, *arrs), "cuda")
    for cap in (total - 7, total // 3):
        moves = torch.full((cap + 64, 32), 0xAB, dtype=torch.uint8, device="cuda")
        out = {"counts": torch.empty(n, dtype=torch.int32, device="cuda"),
               "offsets": torch.empty(n + 1, dtype=torch.int64, device="cuda"), "moves": moves[:cap],
               "overflow": torch.zeros(1, dtype=torch.int64, device="cuda"),
               "mask": torch.empty((n, S * S), dtype=torch.bool, device="cuda")}
        db.movegen(out)
        torch.cuda.synchronize()
        got = moves.cpu().numpy()
        assert (got[cap:] == 0xAB).all()
        assert got[:cap].tobytes() == want["moves"][:cap].tobytes()
        assert np.array_equal(out["mask"].cpu().numpy().view(np.uint8), want["mask"])
        off = want["offsets"].astype(np.int64)
        room = np.minimum(off[1:], cap) - np.minimum(off[:-1], cap)
        assert int(out["overflow"].item()) == int((want["counts"].astype(np.int64) > room).sum())


def test_split_walk_with_a_full_record_pool(drg, monkeypatch):
    """a record pool far too small for the batch: walks that cannot hand over continue without a budget; same result"""
    arrs = _synthetic("frisian", 6000, 7)
    want = O.batch_emit("frisian", *arrs, threads=O.host_threads())
    monkeypatch.setenv("DRG_WALK_POOL_CAP", "3000")
    _fused(drg, "frisian", arrs, want)
    got = drg.BoardBatch("frisian", *arrs).legal_moves(want_mask=True, want_tensor=True)
    assert got["moves"].tobytes() == want["moves"].tobytes() and np.array_equal(got["mask"], want["mask"])
    monkeypatch.setenv("DRG_WALK_POOL_CAP", "0")
    _fused(drg, "frisian", arrs, want)


def test_standalone_emit_splits_too(drg):
    """drg_movegen_count -> drg_scan_counts -> drg_movegen_emit (the emit step builds its own tree-search list and
    walks it): same records as the fused call"""
    import torch
    arrs = _synthetic("frisian", 8000, 11)
    want = O.batch_emit("frisian", *arrs, threads=O.host_threads())
    db = drg.DeviceBoardBatch.from_host(drg.BoardBatch("frisian", *arrs), "cuda")
    r = db.legal_moves(want_mask=True, want_tensor=True)
    total = int(want["offsets"][-1])
    assert np.array_equal(r["counts"].cpu().numpy().astype(np.uint32), want["counts"])
    assert r["moves"].cpu().numpy()[:total].tobytes() == want["moves"].tobytes()
    assert np.array_equal(r["mask"].cpu().numpy().view(np.uint8), want["mask"])
    assert int(r["overflow"].item()) == 0


@pytest.mark.parametrize("variant,depth,expect", [("standard", 9, 41022614), ("frisian", 8, 2907905),
                                                  ("american", 10, 906645019), ("russian", 9, 4571392),
                                                  ("standard", 2, 81), ("standard", 1, 9)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_perft_sums_to_perft(drg, golden_dir, variant, depth, expect, world):
    """drg_perft_sharded: the ranks' shares (dealt by content hash on the device) add up to the perft value; every
    rank of a `world`-GPU job is emulated here, one after the other, on one GPU"""
    import json
    import torch
    deep = json.loads((golden_dir / "perft.json").read_text())["survey_deep"]  # values the reference itself produced
    if depth > 2:
        assert deep[variant][str(depth)] == expect
    b = drg.BoardBatch.start(variant, 1)
    db = drg.DeviceBoardBatch.from_host(b, "cuda")
    total, shares = 0, []
    for rank in range(world):
        ctr_dev = torch.zeros(8, dtype=torch.int64, device="cuda")
        nodes, ctr = db.perft(depth, 0, rank=rank, world=world, counters_dev=ctr_dev)
        assert ctr_dev.cpu().numpy().astype(np.uint64).tolist() == ctr.tolist()
        shares.append(nodes)
        total += nodes
    assert total == expect, (variant, depth, world, shares)
    if depth > 2:
        assert min(shares) > 0 and max(shares) < 2.0 * expect / world  # an even deal


def _first_choice(moves, offsets, counts, action, S):
    """numpy restatement of BaseBoard.index_to_move (base.py:861-891) on packed records"""
    n = len(counts)
    out = np.full(n, -1, np.int32)
    frm = moves["path"][:, 0].astype(np.int64)
    to = moves["path"][np.arange(len(moves)), moves["n_sq"].astype(np.int64) - 1].astype(np.int64)
    act = frm * S + to
    for i in range(n):
        lo, m = int(offsets[i]), int(counts[i])
        hit = np.flatnonzero(act[lo:lo + m] == action[i])
        if len(hit):
            out[i] = hit[0]
    return out


@pytest.mark.parametrize("variant", ["standard", "american", "frisian"])
def test_policy_head_to_move_on_device(drg, variant):
    """drg_sample_action / drg_action_to_choice against the reference's recipe (examples/reinforcement_learning.py:
    103-112 + base.py:861-891): argmax equals torch.argmax of the masked logits, a sampled action is legal and the
    chosen record is the FIRST one with that (from, to), identical calls give identical samples, and on one board
    sampled 20 000 times the frequencies follow softmax(logits / temperature)."""
    import torch
    n, S = 100_000, O.SQUARES[variant]
    rng = np.random.default_rng(3)
    stop = rng.integers(0, 120, n).astype(np.uint16)
    r = O.rollout(variant, n, 5, stop_ply=stop, flags=0, threads=O.host_threads())
    arrs = (r["wm"], r["wk"], r["bm"], r["bk"], r["turn"])
    db = drg.DeviceBoardBatch.from_host(drg.BoardBatch(variant, *arrs), "cuda")
    mv = db.legal_moves(want_mask=True)
    torch.manual_seed(1)
    logits = torch.randn((n, S * S), dtype=torch.float32, device="cuda")
    mask = mv["mask"]
    has = mask.any(dim=1)
    # argmax
    action, choice = db.sample_actions(logits, mv, mode="argmax")
    masked = logits.masked_fill(~mask, -1e9)
    want_a = torch.where(has, masked.argmax(dim=1).to(torch.int32), torch.full_like(action, -1))
    assert torch.equal(action, want_a)
    moves = mv["moves"].cpu().numpy().view(O.MOVE_DTYPE).reshape(-1)
    counts, offsets = mv["counts"].cpu().numpy(), mv["offsets"].cpu().numpy()
    assert np.array_equal(choice.cpu().numpy(), _first_choice(moves, offsets, counts, action.cpu().numpy(), S))
    # sampling
    a1, c1 = db.sample_actions(logits, mv, mode="sample", temperature=0.5, seed=9)
    a2, c2 = db.sample_actions(logits, mv, mode="sample", temperature=0.5, seed=9)
    a3, _ = db.sample_actions(logits, mv, mode="sample", temperature=0.5, seed=10)
    assert torch.equal(a1, a2) and torch.equal(c1, c2) and not torch.equal(a1, a3)
    legal = mask.gather(1, a1.clamp(min=0).to(torch.int64).unsqueeze(1)).squeeze(1)
    assert bool((legal | ~has).all()) and bool(((a1 >= 0) == has).all())
    assert np.array_equal(c1.cpu().numpy(), _first_choice(moves, offsets, counts, a1.cpu().numpy(), S))
    assert torch.equal(db.action_to_choice(a1, mv), c1)
    # an action that is not legal: the reference raises ValueError, the batch call answers -1
    bad = (~mask).to(torch.uint8).argmax(dim=1).to(torch.int32)  # the first index that is not a legal action
    assert bool((db.action_to_choice(bad, mv) == -1).all())
    # the drop-in class agrees (index_to_move through the shim, 300 boards)
    cls = drg.BOARDS[variant]
    fens = drg.BoardBatch(variant, *(a[:300] for a in arrs)).fens()
    for i in range(300):
        if int(a1[i]) < 0:
            continue
        m = cls.from_fen(fens[i]).index_to_move(int(a1[i]))
        rec = moves[int(offsets[i]) + int(c1[i])]
        assert m.square_list == rec["path"][: rec["n_sq"]].tolist()
    # frequencies on one board
    i0 = int(torch.nonzero(mv["counts"] >= 6)[0])
    reps = 20_000
    one = drg.DeviceBoardBatch(variant, *(t[i0:i0 + 1].repeat(reps) for t in (db.wm, db.wk, db.bm, db.bk, db.turn)))
    mv1 = one.legal_moves(want_mask=True)
    lg = logits[i0:i0 + 1].repeat(reps, 1).contiguous()
    acts, _ = one.sample_actions(lg, mv1, mode="sample", temperature=0.7, seed=4)
    p = torch.softmax(masked[i0] / 0.7, dim=0).cpu().numpy().astype(np.float64)
    freq = np.bincount(acts.cpu().numpy(), minlength=S * S) / reps
    assert np.abs(freq - p).max() < 0.02 and freq[p < 1e-12].sum() == 0


def test_selfplay_with_policy_logits_and_clipped_buffers(drg):
    """selfplay_export driven by a policy head on the device: uniform logits reproduce a uniform choice over ACTIONS,
    every game ends, results are reproducible; a record buffer that is too small is reported, not read past."""
    import torch
    from draughts_b200 import _lib
    n = 4096

    def logits_fn(o):
        return torch.zeros((len(o["counts"]), 32 * 32), dtype=torch.float32, device="cuda")

    runs = []
    for _ in range(2):
        db = drg.DeviceBoardBatch.from_host(drg.BoardBatch.start("american", n), "cuda")
        runs.append(db.selfplay_export(seed=3, policy_logits=logits_fn, temperature=1.0, compact_min=512))
    assert torch.equal(runs[0]["result"], runs[1]["result"]) and torch.equal(runs[0]["plies"], runs[1]["plies"])
    assert int((runs[0]["result"] == 0).sum()) == 0 and int(runs[0]["plies"].max()) > 20
    db = drg.DeviceBoardBatch.from_host(drg.BoardBatch.start("standard", n), "cuda")
    with pytest.raises(_lib.DrgError):
        db.selfplay_export(seed=3, moves_per_board=4, check_every=4)


def test_million_game_rollouts(drg):
    """BASELINE configs[3] at full size: 1 Mi Standard games to game_over.  A 64 Ki-game slice equals the oracle bit
    for bit (results, plies, final boards); result and ply histograms of the whole run are those of its slices'
    distribution (games are keyed by id, so the slice is a sub-sample of the same run)."""
    n, k = 1 << 20, 1 << 16
    b = drg.BoardBatch.start("standard", n)
    r = b.rollout(seed=11)
    w = O.rollout("standard", k, 11, threads=O.host_threads())
    assert np.array_equal(r["result"][:k], w["result"]) and np.array_equal(r["plies"][:k], w["plies"])
    for key in ("wm", "wk", "bm", "bk", "turn", "clock"):
        assert np.array_equal(getattr(b, key)[:k], w[key])
    assert int(r["counters"][6]) == 0 and int((r["result"] == 0).sum()) == 0
    # a second, disjoint slice from the far end of the batch (game ids n-k .. n-1)
    w2 = O.rollout("standard", k, 11, game_id0=n - k, threads=O.host_threads())
    assert np.array_equal(r["result"][n - k:], w2["result"]) and np.array_equal(r["plies"][n - k:], w2["plies"])
    h_all = np.bincount(r["result"], minlength=4) / n
    h_slice = np.bincount(w["result"], minlength=4) / k
    assert np.abs(h_all - h_slice).max() < 0.01
    assert abs(r["plies"].mean() - w["plies"].mean()) < 1.0


def test_selfplay_compact_packs_running_games_and_files_finished_ones(drg):
    """drg_selfplay_compact against numpy: finished games land in the result arrays at ids[i], running games are packed
    in order (boards, plies, ids, game ids, the [54][n] move history re-strided to [54][n_running]); dst = NULL files all"""
    import ctypes as C
    import torch
    from draughts_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)
    n, total = 10_000, 25_000
    spec = {"wm": torch.int64, "wk": torch.int64, "bm": torch.int64, "bk": torch.int64, "turn": torch.uint8,
            "clock": torch.uint8, "plies": torch.int16, "state": torch.uint8}

    def rand(dtype, size):
        hi = {torch.int64: 1 << 50, torch.uint8: 200, torch.int16: 3000, torch.int32: 1 << 30}[dtype]
        return torch.from_numpy(rng.integers(0, hi, size=size, dtype=np.int64)).to(dtype).cuda()

    src = {k: rand(dt, n) for k, dt in spec.items()}
    src["state"] = torch.from_numpy(rng.choice([0, 0, 0, 1, 2, 3, 4], size=n).astype(np.uint8)).cuda()
    src["hist"] = rand(torch.int32, 54 * n)
    src["ids"] = torch.from_numpy(rng.permutation(total)[:n].astype(np.int64)).cuda()
    src["game_ids"] = src["ids"] + 77
    dst = {k: torch.full_like(v, 1) for k, v in src.items()}
    fin = {k: torch.zeros(total, dtype=dt, device="cuda") for k, dt in spec.items()}

    def games(d):
        return _lib.DrgGames(**{k: (d[k].data_ptr() if k in d else None) for k, _ in _lib.DrgGames._fields_})

    left = C.c_int64(-1)
    _lib.check(L.drg_selfplay_compact(n, C.byref(games(src)), C.byref(games(dst)), C.byref(games(fin)), C.byref(left), None))
    run = (src["state"] == 0).cpu().numpy()
    m = int(run.sum())
    assert left.value == m and 0 < m < n
    for k in spec:
        a = src[k].cpu().numpy()
        assert np.array_equal(dst[k].cpu().numpy()[:m], a[run]), k
        want = np.zeros(total, a.dtype)
        want[src["ids"].cpu().numpy()[~run]] = a[~run]
        assert np.array_equal(fin[k].cpu().numpy(), want), k
    assert np.array_equal(dst["ids"].cpu().numpy()[:m], src["ids"].cpu().numpy()[run])
    assert np.array_equal(dst["game_ids"].cpu().numpy()[:m], src["game_ids"].cpu().numpy()[run])
    assert np.array_equal(dst["hist"].cpu().numpy()[: 54 * m].reshape(54, m), src["hist"].cpu().numpy().reshape(54, n)[:, run])
    assert bool((dst["wm"][m:] == 1).all()) and bool((dst["hist"][54 * m:] == 1).all())  # nothing past the packed games
    # last call of a job: every game is filed, nothing is packed
    _lib.check(L.drg_selfplay_compact(n, C.byref(games(src)), None, C.byref(games(fin)), None, None))
    for k in spec:
        want = np.zeros(total, src[k].cpu().numpy().dtype)
        want[src["ids"].cpu().numpy()] = src[k].cpu().numpy()
        assert np.array_equal(fin[k].cpu().numpy(), want), k
    # argument errors
    assert L.drg_selfplay_compact(-1, C.byref(games(src)), None, C.byref(games(fin)), None, None) == _lib.E_ARG
    assert L.drg_selfplay_compact(n, None, None, C.byref(games(fin)), None, None) == _lib.E_ARG
    assert L.drg_selfplay_compact(n, C.byref(games(src)), C.byref(games(dst)), C.byref(games(fin)), None, None) == _lib.E_ARG
    bad = dict(src); bad.pop("hist")
    assert L.drg_selfplay_compact(n, C.byref(games(bad)), None, C.byref(games(fin)), None, None) == _lib.E_ARG
    assert L.drg_selfplay_compact(0, C.byref(games(src)), None, C.byref(games(fin)), None, None) == 0


def test_library_owned_nccl_allreduce_single_rank(drg):
    """drg_nccl_unique_id / drg_comm_init / drg_allreduce_counters / drg_comm_destroy on a world of one rank (the
    multi-rank sum is tests/test_gpu_multirank.py, which needs two GPUs): NCCL is bound at run time, the uint64 vector
    comes back unchanged, argument errors are reported"""
    import ctypes as C
    import torch
    from draughts_b200 import _lib
    L = _lib.lib()
    uid = np.zeros(128, np.uint8)
    _lib.check(L.drg_nccl_unique_id(_lib.ptr(uid)))
    assert uid.any()
    comm = C.c_void_p(None)
    _lib.check(L.drg_comm_init(_lib.ptr(uid), 0, 1, C.byref(comm)))
    assert comm.value
    v = torch.tensor([1, 2 ** 40 + 5, 0, 7, 2 ** 62, 9, 10, 11], dtype=torch.int64, device="cuda")
    want = v.clone()
    _lib.check(L.drg_allreduce_counters(comm, v.data_ptr(), v.numel(), None))
    torch.cuda.synchronize()
    assert torch.equal(v, want)
    assert L.drg_allreduce_counters(None, v.data_ptr(), 8, None) == _lib.E_ARG
    assert L.drg_allreduce_counters(comm, None, 8, None) == _lib.E_ARG
    assert L.drg_allreduce_counters(comm, v.data_ptr(), 0, None) == 0
    assert L.drg_comm_init(_lib.ptr(uid), 1, 1, C.byref(C.c_void_p(None))) == _lib.E_ARG
    assert L.drg_nccl_unique_id(None) == _lib.E_ARG
    _lib.check(L.drg_comm_destroy(comm))
    assert L.drg_comm_destroy(None) == 0
