"""N > 1 host logic on CPU: world_size-2 gloo process group.  The product's sharding helpers
(draughts_b200.drivers.shard_indices / allreduce_counters) are exercised with the ORACLE standing in for the GPU
perft of each shard (there is no GPU here): the deal must be a partition, and the all-reduced total must equal
the single-rank perft -- the same additivity the GPU driver relies on."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (ROOT, ROOT / "py-draughts_b200"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from draughts_b200 import drivers, _lib
    from oracle import oracle as O
    variant, depth, split = "standard", 6, 3
    s = O.start_position(variant)
    fr = [np.array([x], np.uint64) for x in s[:4]]
    turn = s[4]
    for _ in range(split):
        fr = list(O.expand(variant, *fr, turn))
        turn ^= 1
    mine = drivers.shard_indices(len(fr[0]), rank, world)
    ctr = np.zeros(_lib.N_COUNTERS, np.uint64)
    ctr[_lib.C_NODES] = O.perft_batch(variant, *[a[mine] for a in fr], turn, depth - split)
    ctr[_lib.C_MOVEGEN] = len(mine)
    total = drivers.allreduce_counters(ctr)
    # rollout sharding: contiguous game-id ranges, results keyed by global game id
    n_games = 101
    lo, hi = n_games * rank // world, n_games * (rank + 1) // world
    r = O.rollout(variant, hi - lo, 9, game_id0=lo, max_plies=60)
    wins = np.zeros(_lib.N_COUNTERS, np.uint64)
    for code, slot in ((1, _lib.C_WHITE_WINS), (2, _lib.C_BLACK_WINS), (3, _lib.C_DRAWS), (0, _lib.C_UNFINISHED)):
        wins[slot] = int((r["result"] == code).sum())
    wins[_lib.C_NODES] = int(r["plies"].astype(np.int64).sum())
    wins = drivers.allreduce_counters(wins)
    if rank == 0:
        out.put((int(total[_lib.C_NODES]), int(total[_lib.C_MOVEGEN]), len(fr[0]), wins.tolist()))
    dist.destroy_process_group()


def test_shard_indices_partition():
    from draughts_b200.drivers import shard_indices
    for n in (0, 1, 7, 658):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if parts else np.array([])
            assert allidx.tolist() == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@pytest.mark.timeout(300)
def test_world2_gloo_perft_and_rollout_additivity():
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nodes, roots, n_frontier, wins = res
    assert nodes == 167140 and roots == n_frontier == 658  # Standard perft(6), 658 depth-3 subtrees
    full = O.rollout("standard", 101, 9, game_id0=0, max_plies=60)
    assert wins[0] == int(full["plies"].astype(np.int64).sum())
    assert wins[2] + wins[3] + wins[4] + wins[5] == 101
    assert wins[2] == int((full["result"] == 1).sum()) and wins[5] == int((full["result"] == 0).sum())
