#!/usr/bin/env python3
"""Search-leaf service throughput: children in move-list order + static evaluation.

    python tools/leaf_service_bench.py [--variant standard] [--positions 1048576] [--repeat 20]

Per step, on positions resident in HBM: drg_movegen (counts, offsets, records) -> drg_children (one child board per
record) -> drg_evaluate of every child.  Prints one JSON line with leaves/s and children/s (CUDA events), the
achieved bandwidth of the two new kernels against their algorithmic bytes, and the CPU oracle timed on a bounded
sample of the same positions (all host threads)."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "py-draughts_b200"):
    sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="standard")
    ap.add_argument("--positions", type=int, default=1 << 20)
    ap.add_argument("--repeat", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import draughts_b200 as d
    from draughts_b200 import _lib
    host = d.random_positions(args.variant, args.positions, seed=0)
    dev = d.DeviceBoardBatch.from_host(host, "cuda")
    mv = dev.legal_moves()
    total = int(mv["offsets"][-1].item())
    ev = lambda: torch.cuda.Event(enable_timing=True)
    res = dev.children(moves=mv, evaluate=True, slots=mv["moves"].shape[0])   # buffers sized once, reused below
    t_all = []
    for it in range(args.repeat + 3):
        e0, e1 = ev(), ev()
        e0.record()
        dev.movegen(mv)
        res = dev.children(moves=mv, evaluate=True, out=res)                    # no host synchronisation in the chain
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            t_all.append(e0.elapsed_time(e1))

    def kernel_ms(fn, reps=20):   # back-to-back launches inside one event pair: device time, not Python time
        fn()
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    L = _lib.lib()
    kid = res["children"]
    t_kids = [kernel_ms(lambda: dev.children(moves=mv, out=res))]
    t_eval = [kernel_ms(lambda: kid.evaluate(out=res["score"]))]
    med = lambda x: float(np.median(x))
    n = args.positions
    kid_bytes = total * (32 + 34 + 4 + 1) + n * (34 + 8)      # record in; child, parent index, status out; parents + offsets in
    eval_bytes = len(kid) * (33 + 8)
    out = {"variant": args.variant, "leaves": n, "children": total, "ms_per_step": med(t_all),
           "leaves_per_s": n / med(t_all) * 1e3, "children_per_s": total / med(t_all) * 1e3,
           "children_kernel_ms": med(t_kids), "children_kernel_GBps": kid_bytes / med(t_kids) / 1e6,
           "evaluate_kernel_ms": med(t_eval), "evaluate_kernel_GBps": eval_bytes / med(t_eval) / 1e6,
           "note": "step = drg_movegen + drg_children + drg_evaluate into reused buffers, no host sync; kernel times from 20 back-to-back launches (evaluate runs over all child slots)"}
    if not args.no_cpu:
        from oracle import oracle as O
        m = min(n, 1 << 17)
        sub = [a[:m] for a in (host.wm, host.wk, host.bm, host.bk, host.turn)]
        th = O.host_threads()
        t0 = time.perf_counter()
        o = O.batch_emit(args.variant, *sub, want_mask=False, want_tensor=False, threads=th)
        par = np.repeat(np.arange(m), o["counts"].astype(np.int64))
        kids = O.batch_apply(args.variant, *(a[par] for a in sub), np.zeros(len(par), np.uint8), o["moves"])
        O.evaluate(args.variant, *kids[:5])
        dt = time.perf_counter() - t0
        out["cpu_oracle"] = {"leaves_per_s": m / dt, "threads_movegen": th, "sample": f"{m} positions, apply + evaluate single-threaded"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
