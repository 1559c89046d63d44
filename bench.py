#!/usr/bin/env python3
"""bench.py -- measurement of the hot path (BASELINE.json: "perft nodes/sec and legal-movegen positions/sec").

Headline workload ("step", BASELINE configs[1]): batched legal_moves + legal-move-mask export for 1 Mi random Standard
(10x10) positions per GPU through ONE stream-ordered C-ABI call (drg_movegen).  Positions are synthetic: seeded
random self-play snapshots (SURVEY 8(d).2 sampling, played on the GPU by the rollout kernel; the sampler is
parity-tested against the oracle).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--quick]

One process per GPU (torchrun for N > 1); rank 0 prints ONE JSON line.
  value        positions/s, inputs resident in HBM, CUDA-event time, max over ranks, L2 flushed between steps
  e2e          same metric through the host-buffer C-ABI call (pinned host inputs -> H2D -> kernels -> D2H), result in
               the reference's layout (one bool byte per mask index); e2e.packed: the documented packed-bit mask output
  roofline     dominant kernel (the mask-row writer) algorithmic bytes / event time vs the measured HBM peak
  sustained    the same step looped for >= 2 s under the clock sampler
  perft        the other half of the BASELINE metric: Standard perft (d12 on one GPU, d13 on several) and American d12
               (configs[2]), the frontier dealt over the ranks on the device, one NCCL all-reduce of the counters;
               perft.sustained: Standard d13 back to back for ~2.6 s under the clock sampler;
               perft.roofline: issue-slot utilisation from the committed ncu capture x the live node rate
  selfplay     configs[3]: 1 Mi random self-play games to game_over with per-ply tensor + mask export, sharded by game id
  king_endgames configs[4]: capture-heavy king-endgame batches (Frisian, Russian, Frysk!), bit-exact vs the CPU port
  cpu_baseline the CPU oracle (C port of the reference algorithm) on all host threads;
  cpu_baseline_reference   the unmodified CPython reference (baseline/_ref) on one core and on all cores
--impl reference times the reference's CPU path on the same config: the C port on all host threads (kind "port";
the CPython reference itself is reported beside it as cpu_baseline_reference).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "py-draughts_b200", ROOT / "tools"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

N_POS = 1 << 20
VARIANT = "standard"
SEED = 0
METRIC = "legal-movegen positions/sec (legal_moves + legal-move mask export, 1Mi random Standard positions per GPU)"
UNIT = "positions/s"
WORKLOAD = "batched legal_moves + legal-move mask, 1Mi random Standard positions per GPU (BASELINE configs[1])"


def workload_config(n_pos: int, mean_moves: float) -> dict:
    """`config` of the JSON line, the same dictionary for both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "positions_per_gpu": n_pos, "variant": VARIANT, "seed": SEED, "mean_moves": mean_moves,
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB memset outside the event pairs) and 2.9 GB of outputs per "
                  "step exceed it; CPU arm: the same 2.9 GB of outputs per step exceed the host's caches"}


def load_profile(name):
    f = ROOT / "profiles" / name
    try:
        return json.loads(f.read_text())
    except Exception:
        return None


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """samples SM clock, power and throttle reasons through NVML while a timed region runs"""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.power, self.reasons, self.max_mhz = [], [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=1.0)
        s = sorted(self.samples)
        out = {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(s)}
        if self.power:
            out["power_w_median"] = sorted(self.power)[len(self.power) // 2]
            out["power_w_max"] = max(self.power)
        return out


# ---------------------------------------------------------------------------------------------------------------
# CPU baselines (rank 0 only; test infrastructure / the unmodified reference -- never part of the measured GPU path)
# ---------------------------------------------------------------------------------------------------------------
def oracle_sample(n: int, game_id0: int, threads: int):
    """the same sampling scheme as draughts_b200.random_positions, played by the CPU oracle"""
    from oracle import oracle as O
    from draughts_b200.batch import splitmix64_np
    g = np.arange(game_id0, game_id0 + n, dtype=np.uint64)
    stop = ((splitmix64_np(np.uint64(SEED ^ 0xD1B54A32D192ED03) + g) >> np.uint64(32)) * np.uint64(81) >> np.uint64(32)).astype(np.uint16)
    r = O.rollout(VARIANT, n, SEED, game_id0=game_id0, max_plies=80, stop_ply=stop, flags=0, threads=threads)
    return r["wm"], r["wk"], r["bm"], r["bk"], r["turn"]


def time_oracle(n_sample: int, steps: int, warmup: int, threads: int):
    """CPU oracle legal_moves + mask on a sample of the workload; returns (positions/s, seconds per step)."""
    from oracle import oracle as O
    arrs = oracle_sample(n_sample, 0, threads)
    counts = O.batch_count(VARIANT, *arrs, threads=threads)
    total = int(counts.sum())
    moves = np.zeros(max(total, 1), O.MOVE_DTYPE)
    mask = np.zeros((n_sample, 2500), np.uint8)
    for _ in range(warmup):
        O.batch_emit(VARIANT, *arrs, want_tensor=False, threads=threads, out=(moves, mask, None))
    t0 = time.perf_counter()
    for _ in range(steps):
        O.batch_emit(VARIANT, *arrs, want_tensor=False, threads=threads, out=(moves, mask, None))
    dt = (time.perf_counter() - t0) / steps
    time_oracle.mean_moves = total / n_sample
    return n_sample / dt, dt


def oracle_perft(variant: str, depth: int, threads: int):
    from oracle import oracle as O
    s0 = O.start_position(variant)
    fr, turn = [np.array([x], np.uint64) for x in s0[:4]], s0[4]
    for _ in range(3):
        fr, turn = list(O.expand(variant, *fr, turn)), turn ^ 1
    t0 = time.perf_counter()
    nodes = O.perft_batch(variant, *fr, turn, depth - 3, threads=threads)
    dt = time.perf_counter() - t0
    return {"variant": variant, "depth": depth, "nodes": nodes, "seconds": dt, "nodes_per_s": nodes / dt, "cores": threads,
            "kind": "port"}


def _ref_worker(args):
    """one process of the CPython-reference baseline: legal_moves + legal_moves_mask of every FEN, timed inside"""
    ref_root, fens = args
    sys.path.insert(0, ref_root)
    import draughts as ref  # the unmodified reference, staged under baseline/_ref (never imported by the product)
    boards = [ref.StandardBoard.from_fen(f) for f in fens]
    t0 = time.perf_counter()
    moves = 0
    for b in boards:
        moves += len(b.legal_moves)
        b.legal_moves_mask()
    return time.perf_counter() - t0, moves


def cpython_reference_baseline(n_sample: int = 20000):
    """the reference itself (pure Python, v1.8.3) on the same sampler: one core, and all cores (one process each)"""
    ref_root = ROOT / "baseline" / "_ref"
    if not (ref_root / "draughts" / "__init__.py").exists():
        return {"unavailable": "baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)"}
    import multiprocessing as mp
    from oracle import oracle as O
    threads = O.host_threads()
    arrs = oracle_sample(n_sample, 0, threads)
    fens = [O.make_fen(int(a), int(b), int(c), int(d), int(t), 50) for a, b, c, d, t in zip(*arrs)]
    ctx = mp.get_context("fork")
    with ctx.Pool(1) as pool:
        dt1, mv1 = pool.map(_ref_worker, [(str(ref_root), fens[: n_sample // 8])])[0]
    chunks = [fens[i::threads] for i in range(threads)]
    with ctx.Pool(threads) as pool:
        res = pool.map(_ref_worker, [(str(ref_root), c) for c in chunks])
    dt_all = max(r[0] for r in res)
    return {"what": "unmodified CPython reference (py-draughts 1.8.3, baseline/_ref): StandardBoard.legal_moves + "
                    "legal_moves_mask() per position, boards built from FEN outside the timed loop",
            "one_core": {"value": (n_sample // 8) / dt1, "unit": UNIT, "cores": 1, "sample": f"{n_sample // 8} positions"},
            "all_cores": {"value": n_sample / dt_all, "unit": UNIT, "cores": threads,
                          "sample": f"{n_sample} positions, one process per core"},
            "kind": "reference"}


def run_reference(args):
    """--impl reference: the reference's CPU path with all host threads on the same config (C port; CPython beside it)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = O.host_threads()
    n_sample = N_POS if not args.quick else 1 << 16
    rate, dt = time_oracle(n_sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(n_sample, time_oracle.mean_moves),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_sample} positions per step (same sampler as the GPU arm), count pass + emit pass with mask"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "perft": oracle_perft(VARIANT, 9, threads),
        "cpu_baseline_reference": cpython_reference_baseline(20000 if not args.quick else 2000),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--perft-depth", type=int, default=0, help="Standard perft depth (default: 12 on one GPU, 13 on several)")
    ap.add_argument("--no-perft", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / selfplay / king_endgames")
    ap.add_argument("--quick", action="store_true", help="small sizes everywhere (smoke run of this script)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_pos = N_POS if not args.quick else 1 << 16

    # CPU baselines first (rank 0 of a single-GPU run only), before CUDA is touched: the CPython reference forks
    cpu = {}
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        threads = O.host_threads()
        n_sample = 1 << 18 if not args.quick else 1 << 14
        rate, dt = time_oracle(n_sample, 3, 1, threads)
        cpu["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n_sample} positions x 3 passes (same sampler), count + emit with mask, {dt:.3f} s per pass"}
        cpu["cpu_baseline_reference"] = cpython_reference_baseline(20000 if not args.quick else 2000)
        cpu["perft_cpu"] = oracle_perft(VARIANT, 9 if not args.quick else 7, threads)

    # keep stdout to the ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1, so fd 1 is pointed at
    # stderr for the run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import ctypes
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import draughts_b200 as d
    from draughts_b200 import _lib
    L = _lib.lib(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def by_rank(x: float):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        got = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(got, t)
        return [float(g.item()) for g in got]

    # ---- inputs: 1 Mi positions per rank, distinct game ids per rank (weak scaling) --------------------------
    host = d.random_positions(VARIANT, n_pos, seed=SEED, game_id0=rank * n_pos)
    db = d.DeviceBoardBatch.from_host(host, dev)
    S = 50
    counts = db.counts()
    total = int(counts.sum().item())
    mean_moves = total / n_pos
    out = {
        "counts": torch.empty(n_pos, dtype=torch.int32, device=dev),
        "offsets": torch.empty(n_pos + 1, dtype=torch.int64, device=dev),
        "moves": torch.empty((total, 32), dtype=torch.uint8, device=dev),
        "mask": torch.empty((n_pos, S * S), dtype=torch.bool, device=dev),
        "overflow": torch.zeros(1, dtype=torch.int64, device=dev),
    }
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(timed: bool):
        # the product path: ONE stream-ordered C-ABI call (drg_movegen): mask rows + count -> scan -> records beside them
        e0, e2 = ev(), ev()
        e0.record()
        db.movegen(out)
        e2.record()
        return (e0, e2) if timed else None

    def phase_times(reps: int):
        """count / scan+emit split of the same work with the stand-alone entry points"""
        cm, em = [], []
        for _ in range(reps):
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            c = db.counts()
            e1.record()
            db.legal_moves(want_mask=True, out=out, counts=c)  # scan + k_emit + tree-search boards
            e2.record()
            torch.cuda.synchronize()
            cm.append(e0.elapsed_time(e1))
            em.append(e1.elapsed_time(e2))
        return cm, em

    for _ in range(args.warmup):
        flush.zero_()
        step(False)
    # W steps are 2 ms of GPU time: the board is still at idle power (263 W against 740 W under load) and, with 8 ranks on
    # one host, the first steps still page in code on every rank.  Untimed steps continue until 0.25 s have passed.
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    warmup_extra = 0
    while time.perf_counter() - t_w < (0.25 if not args.quick else 0.02) and warmup_extra < 2000:
        flush.zero_()
        step(False)
        warmup_extra += 1
        if warmup_extra % 16 == 0:
            torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.drg_launch_count()
    evs = []
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pairs)
        evs.append(step(True))
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps  # host time to queue one step (flush + call)
    barrier()
    launches = L.drg_launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [e0.elapsed_time(e2) for e0, e2 in evs]
    count_ms, emit_ms = phase_times(min(args.steps, 20))
    # the dominant kernel alone: CUDA events recorded by the library around the k_emit launch, on its stream,
    # while the fused product call runs (L2 flushed before each repetition)
    kem = []
    _lib.check(L.drg_profile(1, None))
    for _ in range(min(args.steps, 20)):
        flush.zero_()
        db.movegen(out)
        ms = ctypes.c_float(0)
        _lib.check(L.drg_profile(1, ctypes.byref(ms)))
        kem.append(ms.value)
    # the same kernels one after the other (DRG_SPLIT_SERIAL: nothing runs beside the mask-row kernel): what the
    # overlap costs the dominant kernel, and what it gains the step
    kem_alone, serial_ms = [], []
    os.environ["DRG_SPLIT_SERIAL"] = "1"
    for _ in range(min(args.steps, 10)):
        flush.zero_()
        e0, e1 = ev(), ev()
        e0.record()
        db.movegen(out)
        e1.record()
        ms = ctypes.c_float(0)
        _lib.check(L.drg_profile(1, ctypes.byref(ms)))
        kem_alone.append(ms.value)
        torch.cuda.synchronize()
        serial_ms.append(e0.elapsed_time(e1))
    del os.environ["DRG_SPLIT_SERIAL"]
    _lib.check(L.drg_profile(0, None))
    total_ms_max = max_over_ranks(float(sum(step_ms)))
    ms_by_rank = by_rank(float(sum(step_ms)) / args.steps)
    enqueue_by_rank = by_rank(host_enqueue_ms)
    value = world * n_pos * args.steps / (total_ms_max / 1e3)
    assert int(out["overflow"].item()) == 0

    # ---- sustained: the same step for >= 2 s under the clock sampler ----------------------------------------------
    sustained = None
    if not args.no_extras:
        target_s = 2.2 if not args.quick else 0.3
        n_loop = max(50, int(target_s / max(total_ms_max / args.steps / 1e3 + 6e-5, 1e-5)))
        barrier()
        smp = ClockSampler(local)
        smp.start()
        pairs = []
        t0 = time.perf_counter()
        for _ in range(n_loop):
            flush.zero_()
            pairs.append(step(True))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clk = smp.stop()
        in_step_ms = sum(a.elapsed_time(b) for a, b in pairs)
        in_step_max = max_over_ranks(in_step_ms)
        sustained = {"steps": n_loop, "wall_seconds": wall, "ms_per_step": in_step_max / n_loop,
                     "value": world * n_pos * n_loop / (in_step_max / 1e3), "unit": UNIT,
                     "vs_short_run": (world * n_pos * n_loop / (in_step_max / 1e3)) / value, "clocks": clk,
                     "note": "event time of the steps (the L2 flush between them is outside the event pairs, inside the wall time)"}

    # ---- e2e: host-buffer C-ABI call, pinned host memory, H2D + kernels + D2H inside the timed region -------
    pin = {"counts": _lib.pinned_empty(n_pos, np.uint32), "offsets": _lib.pinned_empty(n_pos + 1, np.uint64),
           "moves": _lib.pinned_empty(total, _lib.MOVE_DTYPE),
           "mask": _lib.pinned_empty((n_pos, S * S), np.uint8)}
    pinned_in = []
    for a in (host.wm, host.wk, host.bm, host.bk, host.turn):
        pa = _lib.pinned_empty(len(a), a.dtype)
        pa[:] = a
        pinned_in.append(pa)
    hb = d.BoardBatch(VARIANT, *pinned_in)
    e2e_steps = max(3, min(args.steps, 5))

    def time_host(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        torch.cuda.synchronize()
        return world * n_pos / max_over_ranks((time.perf_counter() - t0) / e2e_steps)

    e2e_value = time_host(lambda: hb.legal_moves(want_mask=True, out=pin))
    pin_bits = {"counts": pin["counts"], "offsets": pin["offsets"], "moves": pin["moves"],
                "mask_bits": _lib.pinned_empty((n_pos * S * S + 7) // 8, np.uint8)}
    e2e_packed = time_host(lambda: hb.legal_moves(want_mask="bits", out=pin_bits))
    sl = slice(0, 4096)
    assert np.array_equal(np.unpackbits(pin_bits["mask_bits"][: 4096 * S * S // 8], bitorder="little").reshape(4096, S * S),
                          pin["mask"][sl])
    h2d = n_pos * 33
    # counts + offsets + records + mask rows; rows cross PCIe bit-packed (k_pack_mask) and are widened on the host
    d2h = n_pos * 4 + (n_pos + 1) * 8 + total * 32 + (n_pos * S * S + 7) // 8

    # ---- perft: the frontier dealt over the ranks on the device, one counter all-reduce ----------------------------
    def time_perft(variant, depth):
        d.perft(None, depth, variant=variant)  # warm-up at the same depth: allocates every frontier buffer
        barrier()
        t0 = time.perf_counter()
        nodes = d.perft(None, depth, variant=variant)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        return {"variant": variant, "depth": depth, "nodes": nodes, "seconds": dt, "nodes_per_s": nodes / dt,
                "sharding": "one process per GPU: the first frontier of >= 16 Ki boards per rank is dealt by a hash of the "
                            "board content on the device (drg_perft_sharded), one NCCL all-reduce of the counter vector"
                            if world > 1 else "single GPU"}

    perft = None
    if not args.no_perft:
        depth = args.perft_depth or ((12 if world == 1 else 13) if not args.quick else 9)
        perft = time_perft(VARIANT, depth)
        perft["scaling"] = "strong (fixed tree, more GPUs)"
        perft["american_d12"] = time_perft("american", 12 if not args.quick else 9)  # BASELINE configs[2]
        if not args.no_extras and not args.quick:
            # sustained: Standard d13 (70.6 G nodes) back to back for >= 2 s under the clock sampler
            smp = ClockSampler(local)
            smp.start()
            barrier()
            t0 = time.perf_counter()
            runs, nodes13 = 3 * world, 0  # the same count on every rank (the walk ends in a collective): ~2.6 s
            for _ in range(runs):
                nodes13 = d.perft(None, 13, variant=VARIANT)
            barrier()
            dt13 = max_over_ranks(time.perf_counter() - t0)
            perft["sustained"] = {"variant": VARIANT, "depth": 13, "nodes": nodes13, "runs": runs, "wall_seconds": dt13,
                                  "nodes_per_s": nodes13 * runs / dt13, "clocks": smp.stop()}
        prof = load_profile("r2_perft_roofline.json")
        if prof:
            # integer-issue roofline: warp instructions per counted node from the committed ncu capture x the
            # live node rate, against 4 issue slots per SM and clock
            sm_hz = (clocks.get("sm_mhz") or 1965) * 1e6
            peak = 148 * 4 * sm_hz * world
            perft["roofline"] = {"bound": "integer issue", "warp_inst_per_node": prof["warp_inst_per_node"],
                                 "thread_inst_per_node": prof["thread_inst_per_node"],
                                 "achieved_warp_inst_per_s": perft["nodes_per_s"] * prof["warp_inst_per_node"],
                                 "peak_warp_inst_per_s": peak,
                                 "frac": perft["nodes_per_s"] * prof["warp_inst_per_node"] / peak,
                                 "issue_active_pct_by_kernel": prof.get("issue_active_pct"),
                                 "active_threads_per_inst_by_kernel": prof.get("active_threads_per_inst"),
                                 "source": prof.get("source")}

    # ---- configs[3]: 1 Mi-game self-play to game_over with per-ply tensor + mask export, games sharded by id --------
    selfplay = None
    if not args.no_extras:
        games = (1 << 20) if not args.quick else 1 << 14
        lo, hi = games * rank // world, games * (rank + 1) // world

        def fresh():
            return d.DeviceBoardBatch.from_host(d.BoardBatch.start(VARIANT, hi - lo), dev)

        fresh().selfplay_export(SEED, game_id0=lo, max_plies=8, draw_rules=True, ring=2)  # warm-up: ring buffers, kernels
        sp_times = []
        for _ in range(2):  # two whole runs (wall clock around a host-driven loop of ~200 plies): both reported, the faster one counts
            sp_db = fresh()
            barrier()
            t0 = time.perf_counter()
            r = sp_db.selfplay_export(SEED, game_id0=lo, draw_rules=True, ring=2)
            torch.cuda.synchronize()
            barrier()
            sp_times.append(max_over_ranks(time.perf_counter() - t0))
        dt = min(sp_times)
        tot = torch.tensor([r["ply_steps"], r["exported_bytes"], int((r["result"] == 0).sum().item())], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        selfplay = {"games": games, "seconds": dt, "seconds_each_run": [round(x, 5) for x in sp_times], "games_per_s": games / dt, "ply_steps": int(tot[0].item()),
                    "ply_steps_per_s": float(tot[0].item()) / dt, "exported_gb": float(tot[1].item()) / 1e9,
                    "export_gb_per_s": float(tot[1].item()) / 1e9 / dt, "unfinished": int(tot[2].item()),
                    "what": "Standard, draw rules on, every ply: legal_moves records + bool mask [n,2500] + tensor [n,4,50] "
                            "of the current positions into device ring buffers, then one move per running game",
                    "scaling": "strong (games sharded by id range)"}
        del sp_db, r
        torch.cuda.empty_cache()

    # ---- configs[4]: capture-heavy king-endgame batches (N = 1 only: one call per variant, checked against the port) --
    king = None
    if rank == 0 and world == 1 and not args.no_extras:
        import king_endgames as KE
        king = [KE.run(v, 150_000 if not args.quick else 8000, 100, check=True, cpu=not args.no_cpu, reps=3)
                for v in ("frisian", "russian", "frysk")]

    if rank == 0:
        peak, peak_src = measured_peaks()
        # pure-write ceiling of this GPU, measured live: cudaMemsetAsync of the same mask buffer (2.6 GB), CUDA events,
        # best of 5.  The emit kernels are write-only, MEASURED_PEAKS' hbm_gbs is a read+write copy figure.
        wt = []
        for _ in range(5):
            a, b = ev(), ev()
            a.record()
            _lib.check(L.drg_memset_async(out["mask"].data_ptr(), 0, out["mask"].numel(),
                                          torch.cuda.current_stream().cuda_stream))
            b.record()
            torch.cuda.synchronize()
            wt.append(a.elapsed_time(b))
        write_ceiling = out["mask"].numel() / (min(wt) / 1e3) / 1e9
        emit_avg_ms = float(np.median(emit_ms))  # median: the first stand-alone call loads its kernel variant lazily
        k_emit_ms = float(np.mean(kem))
        # the dominant kernel is the mask-row writer (k_emit<mask rows only>): state 33 B in + one S*S-byte row out per
        # board.  Records (32 B each), counts, offsets and the tree-search boards are written by the chain that runs
        # beside it; step_bytes is everything one step reads and writes by the algorithm.
        alg_bytes = n_pos * (33 + 2500)
        step_bytes = n_pos * (33 + 2500) + n_pos * (33 + 4 + 8) + total * 32
        achieved = alg_bytes / (k_emit_ms / 1e3) / 1e9
        ms_step = total_ms_max / args.steps
        traffic = load_profile("r2_mask_kernel_traffic.json") or {"dram_bytes_per_launch": 2562e6 + 43e6,
                                                                   "source": "ncu --set full of the mask-row kernel, round 1 (profiles/r1_s_kernels_ncu_selected.txt)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "warmup_extra_untimed_steps": warmup_extra, "ms_per_step": ms_step, "ms_per_step_by_rank": [round(x, 4) for x in ms_by_rank],
            "host_enqueue_ms_per_step_by_rank": [round(x, 4) for x in enqueue_by_rank], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(n_pos, mean_moves),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BoardBatch.legal_moves(want_mask=True) -> drg_legal_moves_host, pinned host buffers; mask rows bit-packed over PCIe, "
                           "widened to the reference's bool bytes by host threads (drg_host.cpp)",
                    "host_result_bytes_per_step": n_pos * 4 + (n_pos + 1) * 8 + total * 32 + n_pos * S * S,
                    "packed": {"value": e2e_packed, "unit": UNIT,
                               "api": "BoardBatch.legal_moves(want_mask='bits') -> drg_legal_moves_host_bits: the same rows as one flat "
                                      "little-endian bit stream (include/drg.h), no host-side widening; same PCIe bytes",
                               "host_result_bytes_per_step": d2h}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_emit<RulesStandard, mask rows only> (CUDA events around the launch, inside the fused "
                                                     "drg_movegen call, while the count -> scan -> records -> tree-search chain runs BESIDE it "
                                                     "on a second stream and shares the SMs)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["dram_bytes_per_launch"], "traffic_source": traffic["source"],
                         "peak_source": peak_src, "write_only_ceiling_gbs": write_ceiling,
                         "frac_of_write_only_ceiling": achieved / write_ceiling,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_emit_ms,
                         "avg_launch_ms_alone": float(np.mean(kem_alone)),
                         "frac_alone": alg_bytes / (float(np.mean(kem_alone)) / 1e3) / 1e9 / peak,
                         "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step / 1e3) / 1e9,
                                  "frac": step_bytes / (ms_step / 1e3) / 1e9 / peak,
                                  "ms_same_kernels_in_sequence": float(np.median(serial_ms))},
                         "standalone_scan_emit_deep_ms": emit_avg_ms,
                         "count_kernel_ms": float(np.median(count_ms))},
        }
        if sustained:
            line["sustained"] = sustained
        if perft:
            if "perft_cpu" in cpu:
                perft["cpu_baseline"] = cpu.pop("perft_cpu")
            line["perft"] = perft
        if selfplay:
            line["selfplay"] = selfplay
        if king:
            line["king_endgames"] = king
        cpu.pop("perft_cpu", None)
        line.update(cpu)
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
