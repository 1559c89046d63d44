"""Batched board API over structure-of-arrays bitboards -- the throughput surface of the hot path.

`BoardBatch` holds HOST numpy arrays and calls the `_host` entry points of libdrg_b200.so (H2D +
kernels + D2H inside).  `DeviceBoardBatch` holds torch CUDA tensors and calls the device entry
points on torch's current stream, so tensor / mask exports stay on the GPU for RL consumers
(examples/reinforcement_learning.py:95-137 pattern) without a host round trip.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MOVE_DTYPE, VARIANT_ID, check, ptr, squares

_MOVES_PER_BOARD: dict[str, float] = {}  # last large legal_moves() call per variant: records per board (buffer sizing hint)

START = {  # standard.py:93-97, american.py:87-91, frysk.py:33-37  (wm, wk, bm, bk)
    50: (((1 << 20) - 1) << 30, 0, (1 << 20) - 1, 0),
    32: (((1 << 12) - 1) << 20, 0, (1 << 12) - 1, 0),
    "frysk": (((1 << 5) - 1) << 45, 0, (1 << 5) - 1, 0),
}


def start_bitboards(variant: str):
    return START["frysk"] if variant == "frysk" else START[squares(variant)]


class BoardBatch:
    """n independent boards of one variant as host arrays wm, wk, bm, bk (uint64), turn, clock (uint8)."""

    def __init__(self, variant: str, wm, wk, bm, bk, turn, clock=None):
        if variant not in VARIANT_ID:
            raise ValueError(f"unknown variant {variant!r}")
        self.variant = variant
        self.vid = VARIANT_ID[variant]
        self.S = squares(variant)
        self.wm = np.ascontiguousarray(wm, np.uint64)
        self.wk = np.ascontiguousarray(wk, np.uint64)
        self.bm = np.ascontiguousarray(bm, np.uint64)
        self.bk = np.ascontiguousarray(bk, np.uint64)
        self.turn = np.ascontiguousarray(turn, np.uint8)
        n = len(self.wm)
        self.clock = np.zeros(n, np.uint8) if clock is None else np.ascontiguousarray(clock, np.uint8)
        for a in (self.wk, self.bm, self.bk, self.turn, self.clock):
            if len(a) != n:
                raise ValueError("all board arrays must have the same length")

    def __len__(self) -> int:
        return len(self.wm)

    @classmethod
    def start(cls, variant: str, n: int) -> "BoardBatch":
        s = start_bitboards(variant)
        return cls(variant, *(np.full(n, v, np.uint64) for v in s), np.zeros(n, np.uint8))

    def copy(self) -> "BoardBatch":
        return BoardBatch(self.variant, self.wm.copy(), self.wk.copy(), self.bm.copy(), self.bk.copy(),
                          self.turn.copy(), self.clock.copy())

    # ------------------------------------------------------------------------------------------
    def legal_moves(self, want_mask: bool = False, want_tensor: bool = False, out=None):
        """-> dict(counts u32[n], offsets u64[n+1], moves MOVE_DTYPE[total], mask u8[n,S*S]|None, tensor f32[n,4,S]|None).
        Board i's moves, in the reference's canonical order, are moves[offsets[i]:offsets[i+1]].
        `out` may carry preallocated (e.g. pinned) arrays under the same keys.
        want_mask="bits": the mask rows come back packed (key "mask_bits": uint8[ceil(n*S*S/8)], one flat little-endian
        bit stream; `unpack_mask` widens it) -- no host-side widening, the call is bound by PCIe."""
        L = _lib.lib()
        n, S = len(self), self.S
        out = dict(out or {})
        if want_mask == "bits":
            return self._legal_moves_bits(want_tensor, out)
        empty = _lib.result_pool.empty  # pinned blocks recycled by object lifetime for large results, numpy memory else
        counts = out.get("counts")
        if counts is None:
            counts = empty(n, np.uint32)
        mask = out.get("mask") if want_mask else None
        if want_mask and mask is None:
            mask = empty((n, S * S), np.uint8)
        tensor = out.get("tensor") if want_tensor else None
        if want_tensor and tensor is None:
            tensor = empty((n, 4, S), np.float32)
        moves = out.get("moves")
        offsets = out.get("offsets")
        if offsets is None:
            offsets = empty(n + 1, np.uint64)
        total = C.c_int64(0)

        def call(buf):
            return L.drg_legal_moves_host(self.vid, n, ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                                          ptr(self.turn), ptr(counts), ptr(offsets), ptr(buf), 0 if buf is None else len(buf),
                                          C.byref(total), ptr(mask), ptr(tensor))
        if moves is None and n <= 128:
            # small batch (Board.legal_moves is a batch of 1): one call with a generous buffer; the library reports
            # the exact total with DRG_E_OVERFLOW when a fan-out monster does not fit
            moves = np.empty(64 * n + 448, MOVE_DTYPE)
            rc = call(moves)
            if rc == _lib.E_OVERFLOW and total.value > len(moves):
                moves = np.empty(total.value, MOVE_DTYPE)
                rc = call(moves)
            check(rc)
        else:
            rc = None
            hint = _MOVES_PER_BOARD.get(self.variant)
            if moves is None and hint is not None and n >= (1 << 14):
                # large batches in a loop: a record buffer sized from the previous call's moves per board saves the
                # count-only sizing pass (an upload and a count per call); if it is too small the library says how many
                moves = empty(int(n * hint * 1.25) + 4096, MOVE_DTYPE)
                rc = call(moves)
                if rc == _lib.E_OVERFLOW and total.value > len(moves):
                    moves, rc = None, None
            if moves is None:
                # size the record buffer with a count-only pass
                check(L.drg_legal_moves_host(self.vid, n, ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                                             ptr(self.turn), ptr(counts), None, None, 0, C.byref(total), None, None))
                moves = empty(max(total.value, 1), MOVE_DTYPE)
            check(call(moves) if rc is None else rc)
            if n >= (1 << 14):
                _MOVES_PER_BOARD[self.variant] = total.value / n
        return {"counts": counts, "offsets": offsets, "moves": moves[: total.value], "mask": mask, "tensor": tensor}

    def _legal_moves_bits(self, want_tensor, out):
        L = _lib.lib()
        n, S = len(self), self.S
        empty = _lib.result_pool.empty
        counts = out.get("counts") if out.get("counts") is not None else empty(n, np.uint32)
        offsets = out.get("offsets") if out.get("offsets") is not None else empty(n + 1, np.uint64)
        bits = out.get("mask_bits") if out.get("mask_bits") is not None else empty((n * S * S + 7) // 8, np.uint8)
        tensor = out.get("tensor") if want_tensor else None
        if want_tensor and tensor is None:
            tensor = empty((n, 4, S), np.float32)
        moves = out.get("moves")
        total = C.c_int64(0)
        if moves is None:
            check(L.drg_legal_moves_host(self.vid, n, ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                                         ptr(self.turn), ptr(counts), None, None, 0, C.byref(total), None, None))
            moves = empty(max(total.value, 1), MOVE_DTYPE)
        check(L.drg_legal_moves_host_bits(self.vid, n, ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                                          ptr(self.turn), ptr(counts), ptr(offsets), ptr(moves), len(moves),
                                          C.byref(total), ptr(bits), ptr(tensor)))
        return {"counts": counts, "offsets": offsets, "moves": moves[: total.value], "mask": None, "mask_bits": bits,
                "tensor": tensor}

    def unpack_mask(self, mask_bits: np.ndarray) -> np.ndarray:
        """packed rows ("mask_bits") -> bool [n, S*S], the reference's legal_moves_mask layout (base.py:807-838)"""
        n, S = len(self), self.S
        return np.unpackbits(mask_bits, bitorder="little")[: n * S * S].reshape(n, S * S).view(np.bool_)

    def counts(self) -> np.ndarray:
        L = _lib.lib()
        counts = np.empty(len(self), np.uint32)
        total = C.c_int64(0)
        check(L.drg_legal_moves_host(self.vid, len(self), ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                                     ptr(self.turn), ptr(counts), None, None, 0, C.byref(total), None, None))
        return counts

    def evaluate(self) -> np.ndarray:
        """AlphaBetaEngine.evaluate (alpha_beta.py:205-244) of every board -> float64[n] (drg_evaluate_host)."""
        out = np.empty(len(self), np.float64)
        check(_lib.lib().drg_evaluate_host(self.vid, len(self), ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                                           ptr(self.turn), ptr(out)))
        return out

    def legal_moves_mask(self) -> np.ndarray:
        """bool [n, S*S] (base.py:807-838 per board)."""
        return self.legal_moves(want_mask=True)["mask"].view(np.bool_)

    def to_tensor(self) -> np.ndarray:
        """float32 [n, 4, S], perspective = side to move (base.py:753-805 per board)."""
        return self.legal_moves(want_tensor=True)["tensor"]

    def push(self, chosen) -> np.ndarray:
        """board.push(chosen[i]) in place for every board; returns status (1 = the reference would raise ValueError)."""
        L = _lib.lib()
        chosen = np.ascontiguousarray(chosen, MOVE_DTYPE)
        if len(chosen) != len(self):
            raise ValueError("one move per board expected")
        status = np.zeros(len(self), np.uint8)
        check(L.drg_apply_host(self.vid, len(self), ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk),
                               ptr(self.turn), ptr(self.clock), ptr(chosen), ptr(status)))
        return status

    def perft(self, depth: int, turn: int | None = None):
        """pure-movegen perft summed over the batch (all boards must have the same side to move) -> (nodes, counters)."""
        L = _lib.lib()
        if turn is None:
            turn = int(self.turn[0]) if len(self) else 0
        if len(self) and not np.all(self.turn == turn):
            raise ValueError("perft roots must share the side to move")
        nodes = C.c_uint64(0)
        counters = np.zeros(_lib.N_COUNTERS, np.uint64)
        check(L.drg_perft_host(self.vid, len(self), ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk), int(turn),
                               int(depth), C.byref(nodes), ptr(counters)))
        return int(nodes.value), counters

    def rollout(self, seed: int, game_id0: int = 0, max_plies: int = 1000, stop_ply=None, draw_rules: bool = True):
        """random self-play in place (include/drg.h drg_rollout) -> dict(result u8[n], plies u16[n], counters u64[8])."""
        L = _lib.lib()
        n = len(self)
        result = np.zeros(n, np.uint8)
        plies = np.zeros(n, np.uint16)
        counters = np.zeros(_lib.N_COUNTERS, np.uint64)
        sp = None if stop_ply is None else np.ascontiguousarray(stop_ply, np.uint16)
        check(L.drg_rollout_host(self.vid, n, C.c_uint64(seed), C.c_uint64(game_id0), int(max_plies), ptr(sp),
                                 _lib.RF_DRAW_RULES if draw_rules else 0, ptr(self.wm), ptr(self.wk), ptr(self.bm),
                                 ptr(self.bk), ptr(self.turn), ptr(self.clock), ptr(result), ptr(plies),
                                 ptr(counters)))
        return {"result": result, "plies": plies, "counters": counters}

    # ------------------------------------------------------------------------------------------
    @classmethod
    def from_fens(cls, variant: str, fens) -> "BoardBatch":
        """Bulk `Board.from_fen` (base.py:435-495).  Canonical rows are decoded by the threaded native codec
        (drg_fen_decode); any other spelling goes through boards.parse_fen, which raises ValueError like the
        reference when the text is not a FEN at all."""
        from .boards import parse_fen
        S = squares(variant)
        fens = list(fens)
        n = len(fens)
        raw = [f.encode() for f in fens]
        off = np.zeros(n + 1, np.int64)
        if n:
            np.cumsum(np.fromiter((len(r) for r in raw), np.int64, n), out=off[1:])
        text = np.frombuffer(b"".join(raw) + b"\0", np.uint8)
        wm, wk, bm, bk = (np.zeros(n, np.uint64) for _ in range(4))
        turn, status = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        rc = _lib.load().drg_fen_decode(VARIANT_ID[variant], n, ptr(text), ptr(off), ptr(wm), ptr(wk), ptr(bm), ptr(bk),
                                        ptr(turn), ptr(status))
        if rc:
            raise _lib.DrgError(rc, "drg_fen_decode: bad arguments")
        for i in np.flatnonzero(status):
            wm[i], wk[i], bm[i], bk[i], turn[i] = parse_fen(fens[i], S)
        return cls(variant, wm, wk, bm, bk, turn)

    def fens_flat(self):
        """Bulk `Board.fen` (base.py:408-433) -> (text u8[total], off i64[n+1]); row i is text[off[i]:off[i+1]]."""
        n = len(self)
        L = _lib.load()
        off = np.zeros(n + 1, np.int64)
        args = (self.vid, n, ptr(self.wm), ptr(self.wk), ptr(self.bm), ptr(self.bk), ptr(self.turn))
        rc = L.drg_fen_encode(*args, None, 0, ptr(off))
        text = np.empty(max(int(off[n]), 1), np.uint8)
        rc = rc or L.drg_fen_encode(*args, ptr(text), len(text), ptr(off))
        if rc:
            raise _lib.DrgError(rc, "drg_fen_encode failed")
        return text[: int(off[n])], off

    def fens(self) -> list[str]:
        text, off = self.fens_flat()
        blob = text.tobytes().decode("ascii")
        o = off.tolist()
        return [blob[o[i]:o[i + 1]] for i in range(len(self))]


def random_positions(variant: str, n: int, seed: int = 0, max_ply: int = 80, game_id0: int = 0) -> BoardBatch:
    """SURVEY 8(d).2 sampling: n seeded random-play games from the start position (draw rules
    ignored, stop only when no legal move), game g snapshotted after k_g ~ U{0..max_ply} plies, with
    k_g = (splitmix64(seed ^ 0xD1B54A32D192ED03 + g) >> 32) * (max_ply + 1) >> 32.  Played on the GPU."""
    g = np.arange(game_id0, game_id0 + n, dtype=np.uint64)
    stop = ((splitmix64_np(np.uint64(seed ^ 0xD1B54A32D192ED03) + g) >> np.uint64(32)) * np.uint64(max_ply + 1)
            >> np.uint64(32)).astype(np.uint16)
    b = BoardBatch.start(variant, n)
    b.rollout(seed, game_id0=game_id0, max_plies=max_ply, stop_ply=stop, draw_rules=False)
    b.clock[:] = 0
    return b


def splitmix64_np(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


class DeviceBoardBatch:
    """Boards resident in HBM as torch CUDA tensors (int64 storage reinterpreted as uint64 bitboards).
    All calls are enqueued on torch's current stream; outputs are torch tensors on the same device."""

    def __init__(self, variant: str, wm, wk, bm, bk, turn, clock=None):
        import torch
        self.torch = torch
        self.variant, self.vid, self.S = variant, VARIANT_ID[variant], squares(variant)
        self.wm, self.wk, self.bm, self.bk, self.turn = wm, wk, bm, bk, turn
        self.clock = clock if clock is not None else torch.zeros_like(turn)
        for t in (wm, wk, bm, bk):
            assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous()
        assert turn.dtype == torch.uint8 and turn.is_cuda
        self.device = wm.device
        _lib.lib(self.device.index if self.device.index is not None else 0)

    def __len__(self):
        return self.wm.numel()

    @classmethod
    def from_host(cls, batch: BoardBatch, device="cuda", pin: bool = False) -> "DeviceBoardBatch":
        import torch
        def up(a, dt):
            t = torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a)
            return t.to(device, non_blocking=True).to(dt)
        return cls(batch.variant, up(batch.wm, torch.int64), up(batch.wk, torch.int64), up(batch.bm, torch.int64),
                   up(batch.bk, torch.int64), up(batch.turn, torch.uint8), up(batch.clock, torch.uint8))

    def to_host(self) -> BoardBatch:
        c = lambda t: t.cpu().numpy()
        return BoardBatch(self.variant, c(self.wm).view(np.uint64), c(self.wk).view(np.uint64),
                          c(self.bm).view(np.uint64), c(self.bk).view(np.uint64), c(self.turn), c(self.clock))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def counts(self):
        torch = self.torch
        n = len(self)
        counts = torch.empty(n, dtype=torch.int32, device=self.device)
        check(_lib.lib().drg_movegen_count(self.vid, n, self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(),
                                           self.bk.data_ptr(), self.turn.data_ptr(), counts.data_ptr(), self._stream()))
        return counts

    def legal_moves(self, want_mask=False, want_tensor=False, out=None, counts=None):
        """count -> scan -> emit on the current stream.  Returns torch tensors: counts i32[n], offsets i64[n+1],
        moves u8[total,32] (MOVE_DTYPE records), mask bool[n,S*S]|None, tensor f32[n,4,S]|None.  One host sync
        (reading the total) unless `out["moves"]` is supplied."""
        torch = self.torch
        L = _lib.lib()
        n, S, dev = len(self), self.S, self.device
        out = dict(out or {})
        if counts is None:
            counts = self.counts()
        offsets = out.get("offsets")
        if offsets is None:
            offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
        check(L.drg_scan_counts(n, counts.data_ptr(), offsets.data_ptr(), self._stream()))
        moves = out.get("moves")
        if moves is None:
            total = int(offsets[-1].item())
            moves = torch.empty((max(total, 1), 32), dtype=torch.uint8, device=dev)
        mask = out.get("mask") if want_mask else None
        if want_mask and mask is None:
            mask = torch.empty((n, S * S), dtype=torch.bool, device=dev)
        tensor = out.get("tensor") if want_tensor else None
        if want_tensor and tensor is None:
            tensor = torch.empty((n, 4, S), dtype=torch.float32, device=dev)
        ovf = out.get("overflow")
        if ovf is None:
            ovf = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.drg_movegen_emit(self.vid, n, self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(),
                                 self.bk.data_ptr(), self.turn.data_ptr(), offsets.data_ptr(), moves.data_ptr(),
                                 mask.data_ptr() if mask is not None else None,
                                 tensor.data_ptr() if tensor is not None else None, None, ovf.data_ptr(),
                                 self._stream()))
        return {"counts": counts, "offsets": offsets, "moves": moves, "mask": mask, "tensor": tensor, "overflow": ovf}

    def movegen(self, out):
        """count -> scan -> emit into preallocated tensors, stream-ordered, NO host synchronisation (drg_movegen).
        out: counts i32[n], offsets i64[n+1], moves u8[cap,32], overflow i64[1], optional mask bool[n,S*S] and
        tensor f32[n,4,S].  Boards whose records would not fit `moves` are clipped and counted in out["overflow"]."""
        mask, tensor = out.get("mask"), out.get("tensor")
        check(_lib.lib().drg_movegen(self.vid, len(self), self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(),
                                     self.bk.data_ptr(), self.turn.data_ptr(), out["counts"].data_ptr(),
                                     out["offsets"].data_ptr(), out["moves"].data_ptr(), out["moves"].shape[0],
                                     mask.data_ptr() if mask is not None else None,
                                     tensor.data_ptr() if tensor is not None else None, out["overflow"].data_ptr(),
                                     self._stream()))
        return out

    def push(self, chosen, status=None):
        """chosen: u8[n,32] records on the device; boards updated in place."""
        check(_lib.lib().drg_apply(self.vid, len(self), self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(),
                                   self.bk.data_ptr(), self.turn.data_ptr(), self.clock.data_ptr(), chosen.data_ptr(),
                                   status.data_ptr() if status is not None else None, self._stream()))

    # -- search-leaf service -------------------------------------------------------------------
    def evaluate(self, out=None):
        """AlphaBetaEngine.evaluate (alpha_beta.py:205-244) of every board -> float64[n] on the device,
        side-to-move relative, bit-exact with the reference."""
        if out is None:
            out = self.torch.empty(len(self), dtype=self.torch.float64, device=self.device)
        check(_lib.lib().drg_evaluate(self.vid, len(self), self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(),
                                      self.bk.data_ptr(), self.turn.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def children(self, moves=None, evaluate: bool = False, out=None, slots: int | None = None):
        """All children of every board, in `legal_moves` order (what a batched MCTS expansion or an alpha-beta
        frontier asks for).  -> dict(children DeviceBoardBatch, parent i32, status u8, counts, offsets, moves,
        score f64 | None).  `moves` may be the dict a previous legal_moves() / movegen() call returned; otherwise
        the move lists are generated here.  By default the outputs are sized exactly (one host read of the total).
        With `slots` (or a previous result passed back as `out`) they hold that many child slots, only the first
        offsets[-1] are written, and nothing synchronises with the host.  status[j] = 1 marks a move the
        reference's push would reject (the child is then the unchanged parent)."""
        torch = self.torch
        L = _lib.lib()
        n, dev = len(self), self.device
        mv = moves if moves is not None else self.legal_moves()
        if out is not None:
            slots = len(out["children"])
        elif slots is None:
            slots = int(mv["offsets"][-1].item()) if n else 0
        slots = min(int(slots), int(mv["moves"].shape[0]))
        if out is None:
            u64 = lambda: torch.empty(slots, dtype=torch.int64, device=dev)
            u8 = lambda: torch.empty(slots, dtype=torch.uint8, device=dev)
            out = {"children": DeviceBoardBatch(self.variant, u64(), u64(), u64(), u64(), u8(), u8()),
                   "parent": torch.empty(slots, dtype=torch.int32, device=dev), "status": u8(),
                   "score": torch.empty(slots, dtype=torch.float64, device=dev) if evaluate else None}
        kid = out["children"]
        check(L.drg_children(self.vid, n, self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(), self.bk.data_ptr(),
                             self.turn.data_ptr(), self.clock.data_ptr(), mv["offsets"].data_ptr(), slots,
                             mv["moves"].data_ptr(), kid.wm.data_ptr(), kid.wk.data_ptr(), kid.bm.data_ptr(),
                             kid.bk.data_ptr(), kid.turn.data_ptr(), kid.clock.data_ptr(), out["parent"].data_ptr(),
                             out["status"].data_ptr(), self._stream()))
        score = None
        if evaluate:
            score = out.get("score")
            if score is None:
                score = torch.empty(slots, dtype=torch.float64, device=dev)
            kid.evaluate(out=score)
        return {"children": kid, "parent": out["parent"], "status": out["status"], "counts": mv["counts"],
                "offsets": mv["offsets"], "moves": mv["moves"], "score": score}

    def perft(self, depth: int, turn: int, rank: int = 0, world: int = 1, counters_dev=None):
        """pure-movegen perft summed over the batch.  world > 1: THIS rank's share (drg_perft_sharded: the frontier of
        the first wide ply is dealt by a hash of the board content); counters_dev (int64[8] CUDA tensor) receives the
        rank's counter vector on the device, ready for an all-reduce."""
        nodes = C.c_uint64(0)
        counters = np.zeros(_lib.N_COUNTERS, np.uint64)
        check(_lib.lib().drg_perft_sharded(self.vid, len(self), self.wm.data_ptr(), self.wk.data_ptr(), self.bm.data_ptr(),
                                           self.bk.data_ptr(), int(turn), int(depth), int(rank), int(world), C.byref(nodes),
                                           ptr(counters), counters_dev.data_ptr() if counters_dev is not None else None,
                                           self._stream()))
        return int(nodes.value), counters

    def sample_actions(self, logits, mv, mode: str = "sample", temperature: float = 1.0, seed: int = 0,
                       game_id0: int = 0, game_ids=None, plies=None, out=None):
        """The policy-head end of the RL loop (examples/reinforcement_learning.py:103-112) on the device:
        masked softmax over the legal actions of every board -> sampled (or argmax) action -> index of the first
        record with that (from, to) in the board's move list (= BaseBoard.index_to_move).
        logits: float32 [n, S*S] CUDA; mv: what movegen() / legal_moves() returned for these boards.
        -> (action int32[n] = from*S+to or -1, choice int32[n] for drg_selfplay_step / moves[offsets[i]+choice[i]])."""
        torch = self.torch
        n = len(self)
        assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous() and logits.shape == (n, self.S * self.S)
        if out is None:
            out = (torch.empty(n, dtype=torch.int32, device=self.device), torch.empty(n, dtype=torch.int32, device=self.device))
        action, choice = out
        check(_lib.lib().drg_sample_action(self.vid, n, logits.data_ptr(), mv["counts"].data_ptr(), mv["offsets"].data_ptr(),
                                           mv["moves"].data_ptr(), mv["moves"].shape[0], C.c_float(temperature),
                                           {"sample": 0, "argmax": 1}[mode], C.c_uint64(seed), C.c_uint64(game_id0),
                                           game_ids.data_ptr() if game_ids is not None else None,
                                           plies.data_ptr() if plies is not None else None, action.data_ptr(),
                                           choice.data_ptr(), self._stream()))
        return action, choice

    def action_to_choice(self, action, mv, out=None):
        """BaseBoard.index_to_move (base.py:861-891) for a batch: choice[i] = index in board i's move list of the first
        record with (from, to) = divmod(action[i], S); -1 where the reference raises ValueError."""
        torch = self.torch
        n = len(self)
        assert action.is_cuda and action.dtype == torch.int32 and action.is_contiguous()
        choice = out if out is not None else torch.empty(n, dtype=torch.int32, device=self.device)
        check(_lib.lib().drg_action_to_choice(self.vid, n, action.data_ptr(), mv["counts"].data_ptr(), mv["offsets"].data_ptr(),
                                              mv["moves"].data_ptr(), mv["moves"].shape[0], choice.data_ptr(), self._stream()))
        return choice

    def rollout(self, seed, game_id0=0, max_plies=1000, stop_ply=None, draw_rules=True):
        torch = self.torch
        n, dev = len(self), self.device
        result = torch.zeros(n, dtype=torch.uint8, device=dev)
        plies = torch.zeros(n, dtype=torch.int16, device=dev)
        counters = torch.zeros(_lib.N_COUNTERS, dtype=torch.int64, device=dev)
        check(_lib.lib().drg_rollout(self.vid, n, C.c_uint64(seed), C.c_uint64(game_id0), int(max_plies),
                                     stop_ply.data_ptr() if stop_ply is not None else None,
                                     _lib.RF_DRAW_RULES if draw_rules else 0, self.wm.data_ptr(), self.wk.data_ptr(),
                                     self.bm.data_ptr(), self.bk.data_ptr(), self.turn.data_ptr(),
                                     self.clock.data_ptr(), result.data_ptr(), plies.data_ptr(), counters.data_ptr(),
                                     self._stream()))
        return {"result": result, "plies": plies, "counters": counters}

    def selfplay_export(self, seed, game_id0=0, max_plies=1000, draw_rules=True, ring=2, want_mask=True,
                        want_tensor=True, policy=None, on_ply=None, compact=True, check_every=4, moves_per_board=24,
                        compact_min=4096, policy_logits=None, temperature=1.0, policy_mode="sample", compact_frac=0.97):
        """Self-play with per-ply RL export, ply-synchronous over the whole batch (BASELINE configs[3]).

        Every ply: count -> scan -> emit (move records + legal-move mask + tensor of the CURRENT positions, written
        into slot `ply % ring` of device ring buffers) -> drg_selfplay_step (terminal tests, pick a move, push).
        The move is legal_moves[idx] with idx from the counter RNG of drg_rollout, from `policy(slot_out) -> int32[n]`
        (a move-list index per game), or from `policy_logits(slot_out) -> float32[n, S*S]` (a policy head: the masked
        softmax sample / argmax and the index_to_move mapping run on the device, drg_sample_action -- the loop of
        examples/reinforcement_learning.py:103-112 without a host round trip per ply).  `on_ply(ply, slot_out)` lets a consumer read the
        exports before the slot is reused.  Final boards / results equal drg_rollout's for the same seed.
        Every `check_every` plies the host reads 16 bytes (moves made, records clipped); when fewer than `compact_frac`
        of the working games still run (and more than `compact_min` games are left) drg_selfplay_compact writes the
        finished ones to the result arrays and packs the others into a second working set (one flag-scan-gather chain,
        0.1 ms per Mi games; the torch gather per array it replaces took 3.5 ms, which is why it used to wait for 50 %).
        -> dict(result u8[n] (0 unfinished), plies i16[n], ring, ply_steps, exported_bytes)."""
        torch = self.torch
        L = _lib.lib()
        n0, S, dev = len(self), self.S, self.device
        # results of ALL games, indexed by original position in the batch; the working set below shrinks as finished
        # games are dropped from it (drg_selfplay_compact writes their final boards / results into these arrays)
        fin = {k: getattr(self, k).clone() for k in ("wm", "wk", "bm", "bk", "turn", "clock")}
        fin["state"] = torch.zeros(n0, dtype=torch.uint8, device=dev)
        fin["plies"] = torch.zeros(n0, dtype=torch.int16, device=dev)

        def game_set():
            """buffers of one working set, sized for all games (two of them: compaction packs from one into the other)"""
            w = {k: torch.empty_like(getattr(self, k)) for k in ("wm", "wk", "bm", "bk", "turn", "clock")}
            w["state"] = torch.zeros(n0, dtype=torch.uint8, device=dev)
            w["plies"] = torch.zeros(n0, dtype=torch.int16, device=dev)
            w["hist"] = torch.zeros(54 * n0, dtype=torch.int32, device=dev)
            w["ids"] = torch.empty(n0, dtype=torch.int64, device=dev)
            w["game_ids"] = torch.empty(n0, dtype=torch.int64, device=dev)
            return w

        def c_games(w):
            return _lib.DrgGames(**{k: (w[k].data_ptr() if k in w else None) for k, _ in _lib.DrgGames._fields_})

        # both working sets up front: a run that was warmed up (allocator cache) then never allocates in its loop
        cur, other = game_set(), (game_set() if compact else None)
        for k in ("wm", "wk", "bm", "bk", "turn", "clock"):
            cur[k].copy_(getattr(self, k))
        torch.arange(n0, dtype=torch.int64, device=dev, out=cur["ids"])
        torch.add(cur["ids"], int(game_id0), out=cur["game_ids"])
        c_fin = c_games(fin)

        def views(w, n):
            """the first n games of a working set: board batch, state, plies, hist [54][n], ids, game ids"""
            return (DeviceBoardBatch(self.variant, *(w[k][:n] for k in ("wm", "wk", "bm", "bk", "turn", "clock"))),
                    w["state"][:n], w["plies"][:n], w["hist"][: 54 * n], w["ids"][:n], w["game_ids"][:n])

        work, state, plies, hist, ids, gids = views(cur, n0)
        # moves made so far | records clipped: one two-word tensor, so that a host check is ONE device-to-host read; every
        # ring slot reports clipping into the same word (each slot is written on other plies)
        ctl = torch.zeros(2, dtype=torch.int64, device=dev)
        running, overflow = ctl[0:1], ctl[1:2]
        slots = []
        for _ in range(ring):
            slots.append({
                "counts": torch.empty(n0, dtype=torch.int32, device=dev),
                "offsets": torch.empty(n0 + 1, dtype=torch.int64, device=dev),
                "moves": torch.empty((max(n0 * moves_per_board, 1), 32), dtype=torch.uint8, device=dev),
                "mask": torch.empty((n0, S * S), dtype=torch.bool, device=dev) if want_mask else None,
                "tensor": torch.empty((n0, 4, S), dtype=torch.float32, device=dev) if want_tensor else None,
                "overflow": overflow,
            })
        per_ply = (S * S if want_mask else 0) + (16 * S if want_tensor else 0)
        ply = steps = exported = 0
        flags = _lib.RF_DRAW_RULES if draw_rules else 0

        def drop_finished():
            """finished games out to the result arrays, running games packed into the other working set (one native call)"""
            nonlocal cur, other, work, state, plies, hist, ids, gids
            if other is None:
                other = game_set()
            left = C.c_int64(0)
            check(L.drg_selfplay_compact(len(work), C.byref(c_games(cur)), C.byref(c_games(other)), C.byref(c_fin),
                                         C.byref(left), self._stream()))
            cur, other = other, cur
            work, state, plies, hist, ids, gids = views(cur, int(left.value))

        last_total = -1
        while True:
            n = len(work)
            slot = slots[ply % ring]
            o = {"counts": slot["counts"][:n], "offsets": slot["offsets"][: n + 1], "moves": slot["moves"],
                 "overflow": slot["overflow"], "mask": slot["mask"][:n] if want_mask else None,
                 "tensor": slot["tensor"][:n] if want_tensor else None, "ids": ids}
            work.movegen(o)  # count -> scan -> emit, no host synchronisation
            choice = policy(o) if policy is not None else None
            if policy_logits is not None:
                o["plies"] = plies
                _, choice = work.sample_actions(policy_logits(o), o, mode=policy_mode, temperature=temperature, seed=seed,
                                                game_ids=gids, plies=plies)
            check(L.drg_selfplay_step(self.vid, n, C.c_uint64(seed), C.c_uint64(game_id0), int(max_plies), None, flags,
                                      work.wm.data_ptr(), work.wk.data_ptr(), work.bm.data_ptr(), work.bk.data_ptr(),
                                      work.turn.data_ptr(), work.clock.data_ptr(), plies.data_ptr(), state.data_ptr(),
                                      hist.data_ptr(), o["counts"].data_ptr(), o["offsets"].data_ptr(),
                                      o["moves"].data_ptr(), o["moves"].shape[0],
                                      choice.data_ptr() if choice is not None else None,
                                      gids.data_ptr(), running.data_ptr(), self._stream()))
            if on_ply is not None:
                on_ply(ply, o)
            exported += n * per_ply
            ply += 1
            if ply % check_every:
                continue
            # one host sync every `check_every` plies: total moves made so far; unchanged -> every game is over
            total, clipped = ctl.tolist()
            if clipped:
                raise _lib.DrgError(_lib.E_OVERFLOW, f"move buffer of {moves_per_board} records per board overflowed")
            if total == last_total:
                break
            alive_est = (total - last_total) / check_every if last_total >= 0 else n
            last_total = total
            if compact and alive_est < compact_frac * n and n > compact_min:  # drop finished games from the working set
                drop_finished()
        steps = last_total
        check(L.drg_selfplay_compact(len(work), C.byref(c_games(cur)), None, C.byref(c_fin), None, self._stream()))
        for k in ("wm", "wk", "bm", "bk", "turn", "clock"):
            getattr(self, k).copy_(fin[k])
        result = torch.where(fin["state"] == 4, torch.zeros_like(fin["state"]), fin["state"])
        return {"result": result, "plies": fin["plies"], "ring": slots, "ply_steps": steps, "rounds": ply,
                "exported_bytes": exported}
