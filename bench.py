#!/usr/bin/env python3
"""bench.py -- headline measurement of the hot path (BASELINE.json configs[1]).

Workload ("step"): batched legal_moves + legal-move-mask export for 1 Mi random Standard (10x10)
positions per GPU: count kernel -> offsets scan -> emit kernel (packed move records + bool mask
rows).  Positions are synthetic: seeded random self-play snapshots (SURVEY 8(d).2 sampling, played on
the GPU by the rollout kernel; the sampler is parity-tested against the oracle).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU (torchrun for N > 1); rank 0 prints ONE JSON line.
  value     : positions/s, inputs resident in HBM, CUDA-event time, max over ranks, L2 flushed between steps
  e2e       : same metric through the host-buffer C-ABI call (pinned host inputs -> H2D -> kernels -> D2H)
  roofline  : dominant kernel (k_emit, mask writer) algorithmic bytes / event time vs measured HBM peak
  cpu_baseline : the CPU oracle (C port of the reference algorithm) on a bounded sample, all host threads
  perft     : supplementary -- Standard perft nodes/s (root subtrees sharded over ranks, one all-reduce)
--impl reference times the CPU oracle (the reference is pure Python and cannot travel to the GPU box;
see DESIGN.md) on bounded samples of the same workload with all host threads.
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

TRAFFIC_MASK_KERNEL = 2562e6 + 43e6  # DRAM write + read per launch of the mask-row kernel, profiles/r1_s_kernels_ncu_selected.txt

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

N_POS = 1 << 20
VARIANT = "standard"
SEED = 0
METRIC = "legal-movegen positions/sec (legal_moves + legal-move mask export, 1Mi random Standard positions per GPU)"
UNIT = "positions/s"


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons through NVML while the timed region runs"""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_sample(n: int, game_id0: int, threads: int):
    """the same sampling scheme as draughts_b200.random_positions, played by the CPU oracle"""
    from oracle import oracle as O
    from draughts_b200.batch import splitmix64_np
    g = np.arange(game_id0, game_id0 + n, dtype=np.uint64)
    stop = ((splitmix64_np(np.uint64(SEED ^ 0xD1B54A32D192ED03) + g) >> np.uint64(32)) * np.uint64(81) >> np.uint64(32)).astype(np.uint16)
    r = O.rollout(VARIANT, n, SEED, game_id0=game_id0, max_plies=80, stop_ply=stop, flags=0, threads=threads)
    return r["wm"], r["wk"], r["bm"], r["bk"], r["turn"]


def time_oracle(n_sample: int, steps: int, warmup: int, threads: int):
    """CPU oracle legal_moves + mask on a bounded sample; returns (positions/s, seconds per step)."""
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
    return n_sample / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    threads = O.host_threads()
    n_sample = 1 << 18
    rate, dt = time_oracle(n_sample, args.steps, args.warmup, threads)
    # supplementary, like the GPU arm: Standard perft on all host threads (depth-3 root split, dynamic scheduling)
    s0 = O.start_position(VARIANT)
    fr, turn = [np.array([x], np.uint64) for x in s0[:4]], s0[4]
    for _ in range(3):
        fr, turn = list(O.expand(VARIANT, *fr, turn)), turn ^ 1
    t0 = time.perf_counter()
    nodes = O.perft_batch(VARIANT, *fr, turn, 9 - 3, threads=threads)
    pdt = time.perf_counter() - t0
    perft = {"variant": VARIANT, "depth": 9, "nodes": nodes, "seconds": pdt, "nodes_per_s": nodes / pdt, "cores": threads}
    line = {
        "perft": perft,
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "batched legal_moves + legal-move mask, random Standard positions (BASELINE configs[1])",
                   "positions_per_step": n_sample, "variant": VARIANT, "seed": SEED},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_sample} positions per step (same sampler as the GPU arm), count pass + emit pass with mask"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--perft-depth", type=int, default=11)
    ap.add_argument("--no-perft", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    # keep stdout to the ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1, so fd 1 is pointed at
    # stderr for the run and the JSON line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    # ---- inputs: 1 Mi positions per rank, distinct game ids per rank (weak scaling) --------------------------
    host = d.random_positions(VARIANT, N_POS, seed=SEED, game_id0=rank * N_POS)
    db = d.DeviceBoardBatch.from_host(host, dev)
    S = 50
    counts = db.counts()
    total = int(counts.sum().item())
    mean_moves = total / N_POS
    out = {
        "counts": torch.empty(N_POS, dtype=torch.int32, device=dev),
        "offsets": torch.empty(N_POS + 1, dtype=torch.int64, device=dev),
        "moves": torch.empty((total, 32), dtype=torch.uint8, device=dev),
        "mask": torch.empty((N_POS, S * S), dtype=torch.bool, device=dev),
        "overflow": torch.zeros(1, dtype=torch.int64, device=dev),
    }
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    ev = lambda: torch.cuda.Event(enable_timing=True)
    emit_ms, count_ms = [], []

    def step(timed: bool):
        # the product path: ONE stream-ordered C-ABI call (drg_movegen): count (+ tree-search boards) -> scan -> emit
        e0, e2 = ev(), ev()
        e0.record()
        db.movegen(out)
        e2.record()
        return (e0, e2) if timed else None

    def phase_times(reps: int):
        """count / scan+emit split of the same work with the stand-alone entry points (roofline of the emit kernels)"""
        cm, em = [], []
        for _ in range(reps):
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            c = db.counts()
            e1.record()
            db.legal_moves(want_mask=True, out=out, counts=c)  # scan + k_emit + k_emit_deep
            e2.record()
            torch.cuda.synchronize()
            cm.append(e0.elapsed_time(e1))
            em.append(e1.elapsed_time(e2))
        return cm, em

    for _ in range(args.warmup):
        flush.zero_()
        step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.drg_launch_count()
    evs = []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pairs)
        evs.append(step(True))
    barrier()
    launches = L.drg_launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [e0.elapsed_time(e2) for e0, e2 in evs]
    count_ms, emit_ms = phase_times(min(args.steps, 20))
    # the dominant kernel alone: CUDA events recorded by the library around the k_emit launch, on its stream,
    # while the fused product call runs (L2 flushed before each repetition)
    import ctypes
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
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * N_POS * args.steps / (total_ms_max / 1e3)
    assert int(out["overflow"].item()) == 0

    # ---- e2e: host-buffer C-ABI call, pinned host memory, H2D + kernels + D2H inside the timed region -------
    pin = {"counts": _lib.pinned_empty(N_POS, np.uint32), "offsets": _lib.pinned_empty(N_POS + 1, np.uint64),
           "moves": _lib.pinned_empty(total, _lib.MOVE_DTYPE),
           "mask": _lib.pinned_empty((N_POS, S * S), np.uint8)}
    pinned_in = []
    for a in (host.wm, host.wk, host.bm, host.bk, host.turn):
        pa = _lib.pinned_empty(len(a), a.dtype)
        pa[:] = a
        pinned_in.append(pa)
    hb = d.BoardBatch(VARIANT, *pinned_in)
    e2e_steps = max(3, min(args.steps, 5))
    hb.legal_moves(want_mask=True, out=pin)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = hb.legal_moves(want_mask=True, out=pin)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N_POS / float(t.item())
    h2d = N_POS * 33
    # counts + offsets + records + mask rows; rows cross PCIe bit-packed (k_pack_mask) and are widened on the host
    d2h = N_POS * 4 + (N_POS + 1) * 8 + total * 32 + (N_POS * S * S + 7) // 8

    # ---- supplementary: Standard perft, root subtrees sharded over the ranks, one counter all-reduce --------
    perft = None
    if not args.no_perft:
        d.perft(None, args.perft_depth, variant=VARIANT)  # warm-up at the same depth: allocates every frontier buffer
        barrier()
        t0 = time.perf_counter()
        nodes = d.perft(None, args.perft_depth, variant=VARIANT)
        barrier()
        dt = time.perf_counter() - t0
        perft = {"variant": VARIANT, "depth": args.perft_depth, "nodes": nodes, "seconds": dt,
                 "nodes_per_s": nodes / dt, "sharding": "root subtrees round-robin over ranks, 1 all-reduce"}

    if rank == 0:
        peak, peak_src = measured_peaks()
        # pure-write ceiling of this GPU, measured live: cudaMemsetAsync of the same mask buffer (2.6 GB), CUDA events,
        # best of 5.  The emit kernels are write-only, MEASURED_PEAKS' hbm_gbs is a read+write copy figure; this
        # number says what a pure writer can reach here (tools/micro/write_bw.cu compares store flavours).
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
        alg_bytes = N_POS * (33 + 2500)
        step_bytes = N_POS * (33 + 2500) + N_POS * (33 + 4 + 8) + total * 32
        achieved = alg_bytes / (k_emit_ms / 1e3) / 1e9
        ms_step = total_ms_max / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "batched legal_moves + legal-move mask, 1Mi random Standard positions per GPU (BASELINE configs[1])",
                       "positions_per_gpu": N_POS, "variant": VARIANT, "seed": SEED, "mean_moves": mean_moves,
                       "l2": "flushed between timed steps (256 MiB memset outside the event pairs); outputs 2.9 GB/step exceed L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BoardBatch.legal_moves(want_mask=True) -> drg_legal_moves_host, pinned host buffers; mask rows bit-packed over PCIe, "
                           "widened to the reference's bool bytes by host threads (drg_host.cpp)",
                    "host_result_bytes_per_step": N_POS * 4 + (N_POS + 1) * 8 + total * 32 + N_POS * S * S},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_emit<RulesStandard, mask rows only> (CUDA events around the launch, inside the fused "
                                                     "drg_movegen call, while the count -> scan -> records -> tree-search chain runs BESIDE it "
                                                     "on a second stream and shares the SMs)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_MASK_KERNEL, "traffic_source": "ncu --set full of the mask-row kernel, profiles/ (dram__bytes_write + dram__bytes_read per launch)",
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
        if perft:
            line["perft"] = perft
        if not args.no_cpu:
            from oracle import oracle as O
            threads = O.host_threads()
            n_sample = 1 << 18
            rate, dt = time_oracle(n_sample, 3, 1, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_sample} positions x 3 passes (same sampler), count + emit with mask, {dt:.3f} s per pass"}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
