"""Multi-GPU use of the path, limited to where it shards naturally.

1. Independent Markov-chain replicas, one process + engine per GPU, no communication:
   `replica_seed` gives each rank its own random stream.
2. k-vector-sharded full S(k) recompute for large systems: positions are replicated on every
   rank; rank r computes S(k) for a contiguous slice of the (host-ordered, half-space) k list
   and the slices are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests), after
   which every rank holds the full S(k) and adopts it.  The reciprocal energy is the all-reduced
   sum of the per-slice energies.

Two forms of the exchange:
  * fused (the product path, `attach_peers` + `engine.recompute_sk()`): the ranks swap 64-byte handles of their exchange
    blocks once (torch.distributed, any backend); afterwards a recompute is three kernels per rank and no collective call —
    the reduction epilogue of every rank stores its slice into every peer's S(k) through peer-mapped memory and handshakes
    with flags (plum_b200/csrc/pg_sk.cu);
  * all-gather (`sharded_sk_recompute`): slices computed into a torch buffer, `all_gather_into_tensor`, adopted with
    `pg_sk_set` — kept as the plain-collective baseline the fused form is measured against, and for the gloo CPU tests.
Nothing here is on the per-move path.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def replica_seed(base_seed: int, rank: int) -> int:
    """Seed of the rank-th independent replica (what PLUM_SEED is set to for rank's plum_gpu)."""
    return int(base_seed) + 1000 * int(rank)


def k_slice(n_k: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of rank's contiguous share of n_k vectors; shares differ by at most one."""
    q, r = divmod(int(n_k), int(world))
    first = q * rank + min(rank, r)
    return first, q + (1 if rank < r else 0)


def padded_count(n_k: int, world: int) -> int:
    return (int(n_k) + int(world) - 1) // int(world)


def gather_sk(compute_slice: Callable[[int, int], "object"], n_k: int, rank: int, world: int, device="cpu", group=None):
    """All-gather of S(k) slices.  `compute_slice(first, count)` returns this rank's [count, 2] slice
    (numpy array or torch tensor on `device`).  Returns the full [n_k, 2] torch tensor on every rank."""
    import torch
    import torch.distributed as dist
    first, count = k_slice(n_k, rank, world)
    pad = padded_count(n_k, world)
    local = torch.zeros((pad, 2), dtype=torch.float64, device=device)
    if count:
        sl = compute_slice(first, count)
        local[:count] = torch.as_tensor(sl, dtype=torch.float64, device=device).reshape(count, 2)
    if world == 1:
        return local[:n_k].clone()
    full = torch.empty((world * pad, 2), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, local, group=group)
    parts = []
    for r in range(world):
        _, c = k_slice(n_k, r, world)
        parts.append(full[r * pad:r * pad + c])
    return torch.cat(parts, dim=0)


def sharded_sk_recompute(engine, rank: int, world: int, group=None):
    """GPU path: every rank fills its k slice with the engine's kernel straight into a torch CUDA
    buffer, NCCL all-gathers, the engine adopts the full S(k).  Returns (full S(k) tensor,
    reciprocal energy summed over ranks)."""
    import torch
    import torch.distributed as dist
    n_k = engine.ewald_info().n_k_half
    first, count = k_slice(n_k, rank, world)
    pad = padded_count(n_k, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.zeros((pad, 2), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    if count:
        engine.sk_compute_slice(first, count, local.data_ptr())   # synchronous on the engine's stream
    if world > 1:
        full = torch.empty((world * pad, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full, local, group=group)
        parts = []
        for r in range(world):
            _, c = k_slice(n_k, r, world)
            parts.append(full[r * pad:r * pad + c])
        sk = torch.cat(parts, dim=0).contiguous()
    else:
        sk = local[:n_k].contiguous()
    torch.cuda.synchronize()
    engine.sk_set(sk.data_ptr())
    e_local = torch.tensor([engine.sk_energy(first, count) if count else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_local, op=dist.ReduceOp.SUM, group=group)
    return sk, float(e_local.item())


def time_sharded_recompute(engine, rank: int, world: int, e_recip_init: float, reps: int = 30, warm: int = 5, group=None) -> dict:
    """Time the k-sharded full S(k) recompute of `engine`'s resident system (SURVEY.md §8e "Reported": recompute
    time at 1/2/4/8 ranks with the NCCL part broken out).  Compute = host wall clock around the synchronous slice
    kernel (`pg_sk_compute_slice`), collective = CUDA events around the NCCL all-gather on torch's current stream;
    medians over `reps`, max over ranks.  Ends with one real `sharded_sk_recompute` whose all-reduced reciprocal
    energy must reproduce `e_recip_init` (the energy initialisation's) within 1e-10."""
    import time
    import torch
    import torch.distributed as dist
    n_k = engine.ewald_info().n_k_half
    first, count = k_slice(n_k, rank, world)
    pad = padded_count(n_k, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.zeros((pad, 2), dtype=torch.float64, device=dev)
    full = torch.empty((world * pad, 2), dtype=torch.float64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_comp, t_nccl, t_tot = [], [], []
    for it in range(reps + warm):
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        if count:
            engine.sk_compute_slice(first, count, local.data_ptr())
        w1 = time.perf_counter()
        nccl_ms = 0.0
        if world > 1:
            ev[0].record()
            dist.all_gather_into_tensor(full, local, group=group)
            ev[1].record()
            torch.cuda.synchronize()
            nccl_ms = ev[0].elapsed_time(ev[1])
        w2 = time.perf_counter()
        if it >= warm:
            t_comp.append((w1 - w0) * 1e3)
            t_nccl.append(nccl_ms)
            t_tot.append((w2 - w0) * 1e3)
    _, e = sharded_sk_recompute(engine, rank, world, group=group)
    ok = abs(e - e_recip_init) <= 1e-10 * max(1.0, abs(e_recip_init))

    def mx(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    return dict(n_gpus=int(world), n_k=int(n_k), k_per_rank=int(count), compute_ms=mx(float(np.median(t_comp))),
                nccl_allgather_ms=mx(float(np.median(t_nccl))), total_ms=mx(float(np.median(t_tot))),
                allgather_bytes=int(world * pad * 16) if world > 1 else 0, energy_matches_init=bool(ok),
                timing=f"median of {reps}, max over ranks; compute = host wall around the synchronous slice kernel, NCCL = CUDA events")


def attach_peers(engine, rank: int, world: int, group=None) -> None:
    """Exchange the ranks' exchange-block handles and attach them: afterwards `engine.recompute_sk()` is k-sharded."""
    import torch.distributed as dist
    if world <= 1:
        return
    mine = engine.sk_export()
    got = [None] * world
    dist.all_gather_object(got, mine, group=group)
    engine.sk_attach(rank, world, b"".join(got))


def time_fused_recompute(engine, rank: int, world: int, e_recip_init: float, reps: int = 30, warm: int = 5, group=None) -> dict:
    """Time the (k-sharded when world > 1) full S(k) recompute with the exchange fused into the reduction epilogue:
    CUDA-event time from the first kernel to the last (the waits for the peers' flags included), median over `reps`, max
    over ranks.  The reciprocal energy of the result must reproduce `e_recip_init` within 1e-10."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    times, e = [], 0.0
    for it in range(reps + warm):
        if world > 1:
            dist.barrier(group=group)
        engine.recompute_sk_begin()
        e, ms = engine.recompute_sk_end()
        if it >= warm:
            times.append(ms)
    ok = abs(e - e_recip_init) <= 1e-10 * max(1.0, abs(e_recip_init))
    t = torch.tensor([float(np.median(times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    n_k = engine.ewald_info().n_k_half
    return dict(n_gpus=int(world), n_k=int(n_k), k_per_rank=int(k_slice(n_k, rank, world)[1]), total_ms=float(t.item()),
                exchange="slice stored into the peers' S(k) by the reduction epilogue over peer-mapped memory + flag handshake (no NCCL call)"
                if world > 1 else "none (one rank)",
                bytes_stored_to_peers=int(k_slice(n_k, rank, world)[1] * 16 * (world - 1)), energy_matches_init=bool(ok),
                timing=f"CUDA events around the rank's kernels (ready flag, k_sk_block, k_sk_finish incl. waits, energy), median of {reps}, max over ranks")
