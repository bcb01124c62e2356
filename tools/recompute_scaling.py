#!/usr/bin/env python3
"""Full S(k) recompute, k-sharded over the ranks: time per rank with the NCCL part
broken out.  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port 29511 tools/recompute_scaling.py   (or plain `python tools/recompute_scaling.py` for N=1).
Rank 0 prints one JSON line."""
import json, os, sys, time
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
n_k = eng.ewald_info().n_k_half
first, count = sharded.k_slice(n_k, rank, world)
pad = sharded.padded_count(n_k, world)
dev = torch.device("cuda", local_rank)
local = torch.zeros((pad, 2), dtype=torch.float64, device=dev)
full = torch.empty((world * pad, 2), dtype=torch.float64, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
REP = 30
t_comp, t_nccl, t_tot = [], [], []
for it in range(REP + 5):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    if count:
        eng.sk_compute_slice(first, count, local.data_ptr())      # synchronous on the engine's stream
    w1 = time.perf_counter()
    if world > 1:
        ev[0].record(); dist.all_gather_into_tensor(full, local); ev[1].record(); torch.cuda.synchronize()
        nccl_ms = ev[0].elapsed_time(ev[1])
    else:
        nccl_ms = 0.0
    w2 = time.perf_counter()
    if it >= 5:
        t_comp.append((w1 - w0) * 1e3); t_nccl.append(nccl_ms); t_tot.append((w2 - w0) * 1e3)
sk, e = sharded.sharded_sk_recompute(eng, rank, world)
ok = abs(e - t_init["recip"]) <= 1e-10 * max(1.0, abs(t_init["recip"]))


def mx(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = dict(what="k-sharded full S(k) recompute, synthetic S", n_gpus=world, n_k=int(n_k), n_charged=int(np.count_nonzero(s.q)),
           k_per_rank=int(count), compute_ms=mx(float(np.median(t_comp))), nccl_allgather_ms=mx(float(np.median(t_nccl))),
           total_ms=mx(float(np.median(t_tot))), allgather_bytes=int(world * pad * 16), energy_matches_init=bool(ok),
           timing="median of 30, max over ranks; compute = host wall around the synchronous slice kernel, NCCL = CUDA events")
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
