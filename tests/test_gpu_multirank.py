"""Two-GPU test of the sharded paths through NCCL (skipped on a box with one GPU; run by hand with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`): perft dealt over two ranks by the hash of the
board content with the counter all-reduce through the library's own communicator (drg_allreduce_counters), and rollouts
sharded by game id range."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
sys.path.insert(0, "ROOT"); sys.path.insert(0, "ROOT/py-draughts_b200")
import numpy as np, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import draughts_b200 as d
from draughts_b200 import drivers, _lib
out = {}
out["perft9"] = d.perft(None, 9, variant="standard")
out["perft8_frisian"] = d.perft(None, 8, variant="frisian")
out["library_comm"] = drivers.library_comm(dist) is not None
v = torch.tensor([dist.get_rank() + 1, 2 ** 40, 5], dtype=torch.int64, device="cuda")
drivers.allreduce_counters_device(v, dist)
out["sum"] = v.tolist()
r = drivers.rollout("standard", 4000, seed=3)
out["games"] = [int(x) for x in r["counters"][2:6]]
if dist.get_rank() == 0:
    print("RESULT " + json.dumps(out), flush=True)
dist.destroy_process_group()
'''


def test_two_ranks_perft_rollout_and_library_allreduce(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.replace("ROOT", str(ROOT)))
    env = dict(os.environ)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[7:])
    assert out["perft9"] == 41022614 and out["perft8_frisian"] == 2907905
    assert out["library_comm"] is True
    assert out["sum"] == [3, 2 ** 41, 10]
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as O
    w = O.rollout("standard", 4000, 3)
    res = w["result"]
    assert out["games"] == [int((res == 1).sum()), int((res == 2).sum()), int((res == 3).sum()), int((res == 0).sum())]
