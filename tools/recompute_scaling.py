#!/usr/bin/env python3
"""Full S(k) recompute, k-sharded over the ranks: time per rank with the NCCL part
broken out.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port 29511 tools/recompute_scaling.py   (or plain `python tools/recompute_scaling.py` for N=1).
Rank 0 prints one JSON line."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist
from plum_b200 import sharded, synth
from plum_b200.engine import Engine

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
FULL = "--full" in sys.argv   # stress variant S-full: 40 000 charged beads, K = 4138
r, s, types, params = (synth.load_full if FULL else synth.load)(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
eng = Engine(params, device=local_rank, capacity_beads=s.n)
eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
t_init = eng.init_energy()
res = sharded.time_sharded_recompute(eng, rank, world, t_init["recip"])
out = dict(what="k-sharded full S(k) recompute, synthetic " + ("S-full" if FULL else "S"), n_charged=int(np.count_nonzero(s.q)), **res)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
