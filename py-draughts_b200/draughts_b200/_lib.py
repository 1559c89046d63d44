"""ctypes binding of libdrg_b200.so (include/drg.h).  There is no CPU fallback: if the shared
library is missing, or no CUDA device can be bound, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
import weakref
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_CSRC = _HERE.parent / "csrc"
# DRG_SO: another build of the library (measurement aid: variants compiled with DRG_NVCC_EXTRA="-DDRG_SCOUT_WARPS=2" ...)
SO_PATH = Path(os.environ["DRG_SO"]).resolve() if os.environ.get("DRG_SO") else _HERE / "libdrg_b200.so"

VARIANTS = ["standard", "american", "frisian", "russian", "brazilian", "antidraughts", "breakthrough", "frysk"]
VARIANT_ID = {n: i for i, n in enumerate(VARIANTS)}

MOVE_DTYPE = np.dtype([("captured", "<u8"), ("n_sq", "u1"), ("flags", "u1"), ("path", "u1", (22,))])
MAX_PATH = 22
MF_PROMO, MF_KINGMOVE, MF_CAPTURE, MF_TRUNC = 0x01, 0x02, 0x04, 0x80
N_COUNTERS = 8
C_NODES, C_MOVEGEN, C_WHITE_WINS, C_BLACK_WINS, C_DRAWS, C_UNFINISHED, C_OVERFLOW, C_ILLEGAL = range(8)
RF_DRAW_RULES = 0x1

E_ARG, E_CUDA, E_OVERFLOW, E_NOMEM = -1, -2, -3, -4

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-ldl"]


class DrgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libdrg_b200 error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> Path:
    """Compile csrc/drg_abi.cu for sm_100a into libdrg_b200.so (in-tree, next to this file)."""
    srcs = [_CSRC / n for n in ("drg_abi.cu", "drg_host.cpp", "drg_fen.cpp", "drg_kernels.cuh", "drg_core.cuh", "drg_records.cuh", "drg_geom.h")]
    srcs.append(_HERE.parent.parent / "include" / "drg.h")
    if not force and SO_PATH.exists() and all(SO_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    subprocess.check_call([nvcc, *NVCC_FLAGS, *os.environ.get("DRG_NVCC_EXTRA", "").split(), "-Xcompiler", "-pthread", "-o", str(SO_PATH), str(_CSRC / "drg_abi.cu"),
                           str(_CSRC / "drg_host.cpp"), str(_CSRC / "drg_fen.cpp")])
    return SO_PATH


_lib = None
_inited_device = None

_i64, _u64, _vp, _int, _u32 = C.c_int64, C.c_uint64, C.c_void_p, C.c_int, C.c_uint32

class DrgGames(C.Structure):
    """include/drg.h DrgGames: a working set of self-play games (device pointers)"""
    _fields_ = [(k, _vp) for k in ("wm", "wk", "bm", "bk", "turn", "clock", "plies", "state", "hist", "ids", "game_ids")]


_SIGNATURES = {
    "drg_abi_version": (_int, []),
    "drg_init": (_int, [_int]),
    "drg_shutdown": (_int, []),
    "drg_last_error": (C.c_char_p, []),
    "drg_squares": (_int, [_int]),
    "drg_launch_count": (_i64, []),
    "drg_geometry_table": (_int, [_int, _vp]),
    "drg_profile": (_int, [_int, _vp]),
    "drg_memset_async": (_int, [_vp, _int, C.c_size_t, _vp]),
    "drg_host_alloc": (_vp, [C.c_size_t]),
    "drg_host_free": (_int, [_vp]),
    "drg_movegen_count": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_scan_counts": (_int, [_i64, _vp, _vp, _vp]),
    "drg_movegen_emit": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_movegen": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "drg_apply": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_perft": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _vp]),
    "drg_rollout": (_int, [_int, _i64, _u64, _u64, _int, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_perft_sharded": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "drg_selfplay_step": (_int, [_int, _i64, _u64, _u64, _int, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _i64, _vp, _vp, _vp, _vp]),
    "drg_selfplay_compact": (_int, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "drg_nccl_unique_id": (_int, [_vp]),
    "drg_comm_init": (_int, [_vp, _int, _int, _vp]),
    "drg_allreduce_counters": (_int, [_vp, _vp, _int, _vp]),
    "drg_comm_destroy": (_int, [_vp]),
    "drg_sample_action": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _i64, C.c_float, _int, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "drg_action_to_choice": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "drg_legal_moves_host": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "drg_legal_moves_host_bits": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "drg_apply_host": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_perft_host": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _int, _int, _vp, _vp]),
    "drg_rollout_host": (_int, [_int, _i64, _u64, _u64, _int, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_children": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_evaluate": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_pst_table": (_int, [_int, _vp]),
    "drg_evaluate_host": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_fen_decode": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "drg_fen_encode": (_int, [_int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
}
EXPORTED_SYMBOLS = sorted(_SIGNATURES)


def load():
    """dlopen the library and set the signatures (no CUDA call yet)."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise RuntimeError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  draughts_b200 has no CPU fallback.")
        L = C.CDLL(str(SO_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return load().drg_last_error().decode(errors="replace")


def check(rc: int) -> int:
    if rc != 0:
        raise DrgError(rc, last_error())
    return rc


def lib(device: int | None = None):
    """The initialised library, bound to `device` (default: LOCAL_RANK or 0)."""
    global _inited_device
    L = load()
    if device is None:
        device = _inited_device if _inited_device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    if _inited_device != device:
        check(L.drg_init(int(device)))
        _inited_device = device
    return L


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data


def squares(variant: str) -> int:
    return 32 if variant in ("american", "russian", "brazilian") else 50


class _PinnedPool:
    """Result buffers of the host-buffer calls, recycled by object lifetime.

    A large `BoardBatch.legal_moves()` result (2.6 GB of mask rows per Mi Standard boards) written into fresh pageable
    numpy memory is bound by page faults and staged copies (127 ms per Mi boards against 23 ms into pinned memory).
    Callers may pass pinned arrays through `out=`; for those who do not, arrays come from this pool: cudaMallocHost
    blocks in size classes, handed out as numpy arrays and taken back when the LAST view of an array has been garbage
    collected (a weakref finalizer on the owning ctypes object) -- so a loop that drops its previous result reuses the
    same pinned memory and no live result is ever aliased.  A size class is only pinned from the second request on
    (a one-shot call is not worth the cost of pinning), and the pool never holds more than DRG_PINNED_POOL_MB
    (default: a quarter of the host's memory, at most 16 GiB, divided by LOCAL_WORLD_SIZE); beyond that, or below 1 MiB, plain numpy memory is used."""

    def __init__(self, alloc=None):
        self.alloc = alloc  # bytes -> address (default: drg_host_alloc = cudaMallocHost); tests pass their own
        self.lock = threading.Lock()
        self.free: dict[int, list[int]] = {}
        self.seen: dict[int, int] = {}
        self.held = 0
        self.limit = None

    def _limit(self) -> int:
        if self.limit is None:
            mb = os.environ.get("DRG_PINNED_POOL_MB")
            if mb is not None:
                self.limit = int(mb) << 20
            else:
                try:
                    ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
                except (ValueError, OSError):
                    ram = 8 << 30
                ranks = max(int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1), 1)  # one process per GPU shares the host
                self.limit = min(ram // 4, 16 << 30) // ranks
        return self.limit

    @staticmethod
    def size_class(nbytes: int) -> int:
        step = max(1 << 16, 1 << max(nbytes.bit_length() - 4, 0))  # at most 1/8 above the request
        return (nbytes + step - 1) // step * step

    def _give_back(self, addr: int, cls: int):
        with self.lock:
            self.free.setdefault(cls, []).append(addr)

    def empty(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = count * dtype.itemsize
        if nbytes < (1 << 20):
            return np.empty(shape, dtype)
        cls = self.size_class(nbytes)
        with self.lock:
            stack = self.free.get(cls)
            addr = stack.pop() if stack else None
            if addr is None:
                self.seen[cls] = self.seen.get(cls, 0) + 1
                if self.seen[cls] < 2 or self.held + cls > self._limit():
                    return np.empty(shape, dtype)
                self.held += cls
        if addr is None:
            addr = self.alloc(cls) if self.alloc else lib().drg_host_alloc(cls)
            if not addr:
                with self.lock:
                    self.held -= cls
                return np.empty(shape, dtype)
        owner = type("PinnedBlock", (C.c_uint8 * cls,), {}).from_address(addr)
        weakref.finalize(owner, self._give_back, addr, cls)
        return np.frombuffer(owner, dtype=dtype, count=count).reshape(shape)


result_pool = _PinnedPool()


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by cudaMallocHost memory (freed with the process)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib().drg_host_alloc(max(n, 1))
    if not p:
        raise DrgError(E_NOMEM, last_error())
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
