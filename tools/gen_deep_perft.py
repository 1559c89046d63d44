#!/usr/bin/env python3
"""Deep perft pins from the CPU oracle (test infrastructure): prints {variant: {depth: nodes}} for the start position.

    python tools/gen_deep_perft.py standard 12 [threads]

The values are merged by hand into tests/golden/perft_deep_oracle.json (the oracle is pinned to the reference by
tests/test_oracle_golden.py; American d12 = 63 633 490 505 was produced by the same call in 810 s on 8 threads)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

variant, depth = sys.argv[1], int(sys.argv[2])
threads = int(sys.argv[3]) if len(sys.argv) > 3 else O.host_threads()
s0 = O.start_position(variant)
fr, turn = [np.array([x], np.uint64) for x in s0[:4]], s0[4]
split = min(3, depth - 1)
for _ in range(split):
    fr, turn = list(O.expand(variant, *fr, turn)), turn ^ 1
t0 = time.perf_counter()
nodes = O.perft_batch(variant, *fr, turn, depth - split, threads=threads)
print(json.dumps({"variant": variant, "depth": depth, "nodes": nodes, "seconds": round(time.perf_counter() - t0, 1),
                  "threads": threads}), flush=True)
