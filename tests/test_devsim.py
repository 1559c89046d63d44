"""Host replay of the split walk (tools/devsim) against the oracle -- no GPU.

The tree-search kernels run every board's capture search as budgeted walks whose unexplored parts are handed to
later rounds (csrc/drg_core.cuh: Walker / WalkLane; csrc/drg_kernels.cuh: k_walk_*).  tools/devsim compiles that
per-thread code with g++ and schedules it round by round the way the kernels do; here its counts and records must
equal the oracle's on king endgames (deep trees, Frisian value filter and king preference, Russian mid-capture
promotion), with budgets small enough that almost every tree is split, and with a record pool too small to hold
them (walks then continue without a budget).  The GPU tests check the kernels themselves."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from oracle import oracle as O  # noqa: E402

FAM = {"standard": 0, "american": 1, "frisian": 2, "russian": 3, "brazilian": 4, "antidraughts": 0, "frysk": 2}


@pytest.fixture(scope="module")
def sim():
    so = ROOT / "tools" / "devsim" / "libdevsim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(so), str(ROOT / "tools" / "devsim" / "devsim.cpp")])
    return C.CDLL(str(so))


def _split(sim, variant, arrs, budgets, cap, explode):
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    wm, wk, bm, bk, turn = arrs
    n = len(wm)
    b = np.asarray(budgets, np.int32)
    stats = np.zeros(98, np.uint64)
    counts = np.zeros(n, np.uint32)
    sim.sim_set_explode(explode)
    rc = sim.sim_split(FAM[variant], C.c_int64(n), p(wm), p(wk), p(bm), p(bk), p(turn), p(b), len(b), C.c_int64(cap), p(counts),
                       None, None, p(stats))
    assert rc == 0, rc
    offsets = np.zeros(n + 1, np.uint64)
    np.cumsum(counts, out=offsets[1:])
    moves = np.zeros(int(offsets[-1]) + 1, O.MOVE_DTYPE)
    c2 = np.zeros(n, np.uint32)
    rc = sim.sim_split(FAM[variant], C.c_int64(n), p(wm), p(wk), p(bm), p(bk), p(turn), p(b), len(b), C.c_int64(cap), p(c2),
                       p(offsets), p(moves), p(stats))
    assert rc == 0, rc
    return counts, moves[:-1], stats


@pytest.mark.parametrize("variant", ["standard", "american", "frisian", "russian", "brazilian", "frysk"])
@pytest.mark.parametrize("budgets,cap,explode", [((2, 3, 4, 4, 6, 8, 8, 16, 32, 64, 0), 1 << 22, 1),
                                                  ((1, 1, 2, 2, 4, 4, 8, 8, 16, 16, 32, 64, 0), 1 << 22, 2),
                                                  ((16, 24, 32, 32, 32, 0), 1 << 22, 0),
                                                  ((3, 3, 3, 0), 500, 1)])
def test_split_walk_equals_oracle(sim, variant, budgets, cap, explode):
    from fuzz_parity import synthetic
    S = O.SQUARES[variant]
    arrs = synthetic(variant, S, 3000, np.random.default_rng(5))
    counts, moves, stats = _split(sim, variant, arrs, budgets, cap, explode)
    want = O.batch_emit(variant, *arrs, want_mask=False, want_tensor=False, threads=O.host_threads())
    assert np.array_equal(counts, want["counts"])
    assert moves.tobytes() == want["moves"].tobytes()
    if cap > 1000 and variant in ("frisian", "frysk", "standard"):
        assert stats[97] > 0  # trees really were split
