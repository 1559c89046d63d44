"""perft and random-rollout drivers (new callers; the reference has neither).

Multi-GPU: one process per GPU (torchrun); work shards by root subtree (perft) or by game index
range (rollouts) with NO data-path exchange; the only collective is one all-reduce of the uint64 counter
vector: drg_allreduce_counters (NCCL over NVLink through a communicator the library owns) when the process
group is nccl, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .batch import BoardBatch, start_bitboards
from .boards import BaseBoard


def _as_batch(root, variant: str | None) -> BoardBatch:
    if isinstance(root, BoardBatch):
        return root
    if isinstance(root, BaseBoard):
        return root._batch1()
    if root is None:
        if variant is None:
            raise ValueError("variant needed")
        return BoardBatch.start(variant, 1)
    if isinstance(root, str):
        return BoardBatch.from_fens(variant, [root])
    raise TypeError(type(root))


def expand(batch: BoardBatch) -> BoardBatch:
    """All children of every board (canonical order), as a new batch (legal_moves + push on the GPU)."""
    out = batch.legal_moves()
    counts = out["counts"].astype(np.int64)
    idx = np.repeat(np.arange(len(batch)), counts)
    child = BoardBatch(batch.variant, batch.wm[idx], batch.wk[idx], batch.bm[idx], batch.bk[idx], batch.turn[idx],
                       batch.clock[idx])
    if len(child):
        status = child.push(out["moves"])
        if status.any():  # the reference raises ValueError on these (two-pieces-on-one-square quirk); perft skips them
            keep = status == 0
            child = BoardBatch(child.variant, child.wm[keep], child.wk[keep], child.bm[keep], child.bk[keep],
                               child.turn[keep], child.clock[keep])
    return child


def shard_indices(n: int, rank: int, world: int) -> np.ndarray:
    """Round-robin deal of n root subtrees: rank r owns i with i % world == r."""
    return np.arange(rank, n, world)


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


_LIBRARY_COMM = {}  # id of the process group -> library-owned NCCL communicator (void*), or None when it could not be made


def library_comm(dist):
    """The NCCL communicator the library owns for this process group (include/drg.h drg_comm_init), made on first use:
    rank 0's unique id travels over the existing group, then every rank joins.  Whether it can be made at all (an NCCL
    library the loader finds) is agreed on by ALL ranks first, so nobody waits alone in a collective; None -> the
    caller uses torch.distributed for the all-reduce instead."""
    import ctypes as C
    import torch
    key = id(dist.group.WORLD)
    if key in _LIBRARY_COMM:
        return _LIBRARY_COMM[key]
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    uid = np.zeros(128, np.uint8)
    ok = L.drg_nccl_unique_id(_lib.ptr(uid)) == 0  # every rank probes the binding; rank 0's id is the one that counts
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm = None
    if int(flag.item()) == 1:
        t = torch.from_numpy(uid).to(dev)
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy()
        handle = C.c_void_p(None)
        rc = L.drg_comm_init(_lib.ptr(uid), dist.get_rank(), dist.get_world_size(), C.byref(handle))
        joined = torch.tensor([1 if rc == 0 and handle.value else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(joined, op=dist.ReduceOp.MIN)
        if int(joined.item()) == 1:
            comm = handle
    _LIBRARY_COMM[key] = comm
    return comm


def allreduce_counters_device(ctr_dev, dist):
    """in-place sum of a CUDA int64 / uint64 counter vector over the ranks: drg_allreduce_counters (NCCL through the
    library's own communicator) on the current stream, torch.distributed if the library could not bind NCCL"""
    import ctypes as C
    import torch
    comm = library_comm(dist)
    if comm is None:
        dist.all_reduce(ctr_dev, op=dist.ReduceOp.SUM)
        return ctr_dev
    stream = C.c_void_p(torch.cuda.current_stream(ctr_dev.device).cuda_stream)
    _lib.check(_lib.lib().drg_allreduce_counters(comm, ctr_dev.data_ptr(), ctr_dev.numel(), stream))
    return ctr_dev


def allreduce_counters(counters: np.ndarray) -> np.ndarray:
    """Sum a uint64 counter vector over all ranks (no-op without an initialised process group)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return counters
    import torch
    t = torch.from_numpy(counters.astype(np.int64))
    if dist.get_backend() == "nccl":
        return allreduce_counters_device(t.cuda(), dist).cpu().numpy().astype(np.uint64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def perft(root=None, depth: int = 1, variant: str | None = None, split_depth: int | None = None,
          return_counters: bool = False):
    """Pure-movegen perft of a Board / FEN / BoardBatch / start position.

    With an initialised torch.distributed group of W > 1 ranks the work is dealt below the root and the per-rank
    counter vectors are summed with ONE all-reduce; every rank returns the same total.  NCCL groups: the deal
    happens inside the library on the device (drg_perft_sharded).  Other groups / an explicit `split_depth`: the
    frontier at that depth is dealt round-robin by the host."""
    batch = _as_batch(root, variant)
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    if world == 1:
        nodes, ctr = batch.perft(depth)
        return (nodes, ctr) if return_counters else nodes
    if dist.get_backend() == "nccl" and split_depth is None:
        # every rank gives the library the same roots; the library expands the shallow plies identically on every
        # GPU, deals the first wide frontier by a hash of the board content and walks this rank's share
        # (drg_perft_sharded); the counter vector stays on the device for the one all-reduce
        import torch
        from .batch import DeviceBoardBatch
        dev = torch.device("cuda", torch.cuda.current_device())
        db = DeviceBoardBatch.from_host(batch, dev)
        ctr_dev = torch.zeros(_lib.N_COUNTERS, dtype=torch.int64, device=dev)
        turn = int(batch.turn[0]) if len(batch) else 0
        db.perft(depth, turn, rank=rank, world=world, counters_dev=ctr_dev)
        allreduce_counters_device(ctr_dev, dist)
        ctr = ctr_dev.cpu().numpy().astype(np.uint64)
        nodes = int(ctr[_lib.C_NODES])
        return (nodes, ctr) if return_counters else nodes
    if depth <= 1:
        nodes, ctr = batch.perft(depth)
        return (nodes, ctr) if return_counters else nodes
    # host-side deal (gloo process groups of the CPU tests, or an explicit split_depth): root subtrees round-robin
    frontier, d = batch, 0
    limit = depth - 1 if split_depth is None else min(split_depth, depth - 1)
    while d < limit and (split_depth is not None or len(frontier) < 64 * world):
        frontier = expand(frontier)
        d += 1
    mine = shard_indices(len(frontier), rank, world)
    shard = BoardBatch(frontier.variant, frontier.wm[mine], frontier.wk[mine], frontier.bm[mine], frontier.bk[mine],
                       frontier.turn[mine])
    ctr = np.zeros(_lib.N_COUNTERS, np.uint64)
    if len(shard):
        _, ctr = shard.perft(depth - d)
    ctr = allreduce_counters(ctr)
    nodes = int(ctr[_lib.C_NODES])
    return (nodes, ctr) if return_counters else nodes


def perft_divide(root=None, depth: int = 1, variant: str | None = None) -> dict[str, int]:
    """Node count below each legal move of a single root (debugging aid): {str(move): nodes}."""
    from .boards import BOARDS
    if isinstance(root, BaseBoard):
        board = root
    else:
        v = variant or "standard"
        board = BOARDS[v].from_fen(root) if isinstance(root, str) else BOARDS[v]()
    out = {}
    for m in board.legal_moves:
        child = board.copy()
        child.push(m)
        out[str(m)] = 1 if depth <= 1 else child._batch1().perft(depth - 1)[0]
    return out


def rollout(variant: str, n_games: int, seed: int = 0, max_plies: int = 1000, draw_rules: bool = True,
            game_id0: int = 0):
    """n_games random self-play games from the start position.  Sharded by contiguous game-id ranges over the
    ranks of an initialised process group (results are keyed by global game id, so they do not depend on the
    number of GPUs).  -> dict(result, plies, final BoardBatch of THIS rank's games, lo, hi, counters summed over ranks)."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    lo, hi = n_games * rank // world, n_games * (rank + 1) // world
    b = BoardBatch.start(variant, hi - lo)
    r = b.rollout(seed, game_id0=game_id0 + lo, max_plies=max_plies, draw_rules=draw_rules)
    r["counters"] = allreduce_counters(r["counters"])
    r.update(boards=b, lo=lo, hi=hi)
    return r
