import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "py-draughts_b200"))
import torch, numpy as np
import draughts_b200 as d
N = 1 << 20
host = d.random_positions("standard", N, seed=0)
db = d.DeviceBoardBatch.from_host(host, "cuda")
c = db.counts()
r = db.legal_moves(want_mask=True, counts=c)
out = {k: r[k] for k in ("offsets", "moves", "mask", "overflow")}
for i in range(3):
    c = db.counts()
    db.legal_moves(want_mask=True, out=out, counts=c)
torch.cuda.synchronize()
print("ok")
import time
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(5):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t0 = time.perf_counter()
    e[0].record()
    c = db.counts()
    t1 = time.perf_counter()
    e[1].record()
    db.legal_moves(want_mask=True, out=out, counts=c)
    t2 = time.perf_counter()
    e[2].record()
    torch.cuda.synchronize()
    print("count gpu %.3f host %.3f | emit gpu %.3f host %.3f" % (e[0].elapsed_time(e[1]), (t1-t0)*1e3, e[1].elapsed_time(e[2]), (t2-t1)*1e3))
