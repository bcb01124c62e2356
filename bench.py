#!/usr/bin/env python3
"""bench.py — MC moves/s of the per-trial-move energy path on the synthetic 22k-bead
polyelectrolyte (BASELINE.json configs[4], SURVEY.md §8(d) "S").

  python bench.py --gpus N --steps K --warmup W            this repo (CUDA engine via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU algorithm

One *step* = MOVES_PER_STEP sequential Metropolis moves (move mix 0.5 ion translation /
0.1 COM / 0.3 pivot / 0.1 reptation) of one replica.  One replica per GPU, no communication
(SURVEY.md §8(e)): `value` is the aggregate over all ranks, scaling "weak".

  e2e    the loop a driver runs: host-generated trial coordinates -> pg_delta_e (H2D copy,
         fused kernel, D2H result) -> Metropolis test on the host -> pg_commit; native C++
         caller (plum_b200/host/mc_bench.cc) so no interpreter is inside the timed region.
  value  the SAME move sequence replayed device-resident (pg_replay_run): proposals already in
         HBM, acceptance taken on the device from the recorded variates, CUDA-event time.
  roofline  dominant kernel k_move: algorithmic FP64 flops (SURVEY.md §8(d) formula) over its
         CUDA-event time, against an FP64 FMA peak measured in this run (MEASURED_PEAKS.json
         carries no FP64 figure); the byte view against MEASURED_PEAKS.json's hbm_gbs is added.
  cpu_baseline  the CPU oracle port (pairwise reciprocal form = the reference's algorithm,
         oracle/plum_oracle.c) on 1 host core, bounded sample.
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

MOVES_PER_STEP = 1024
K_FULL = 3574            # k vectors the reference sums per pair on S
F_PAIR = 36.0            # algorithmic flops of one pair-configuration evaluation
WORKLOAD = ("synthetic_S: 200 chains x 100 beads (every 10th q=-1) + 2000 counter-ions, N=22000, L=200, "
            "lB=2.5, alpha=0.004 (real_cutoff 52, 3574 k-vectors), WCA sigma=2.227, "
            "moves 0.5 ion-translate/0.1 COM/0.3 pivot/0.1 reptation")
MOVE_PROB = [0.5, 0.1, 0.3, 0.0, 0.1]


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------- mc bench lib
def mcbench_lib():
    path = os.path.join(REPO, "plum_b200", "libplum_mcbench.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run __graft_entry__.build()")
    L = C.CDLL(path)
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    L.pb_create.restype = C.c_void_p
    L.pb_create.argtypes = [C.c_void_p, C.c_int, ip, dp, dp, C.c_double, C.c_double, C.c_double, dp, C.c_int, C.c_uint]
    L.pb_destroy.argtypes = [C.c_void_p]
    L.pb_run.argtypes = [C.c_void_p, C.c_int, ip, ip, dp, dp, bp, dp, bp, C.c_int, ip, dp, dp, dp, ip, C.c_int]
    return L


def alg_flops_per_move(sysm, mols):
    """flops = 2*evals*F_pair + 2*n_moved_q*K*16 + 12*K; bytes = 40*N + 40*K + 48*n_moved."""
    first = sysm.mol_first
    N = sysm.n
    length = (first[mols + 1] - first[mols]).astype(np.float64)
    cs = np.concatenate([[0.0], np.cumsum(sysm.q != 0)])
    nq = cs[first[mols + 1]] - cs[first[mols]]
    evals = length * (N - length) + 0.5 * length * (length - 1)
    flops = 2 * evals * F_PAIR + 2 * nq * K_FULL * 16 + 12 * K_FULL
    bytes_ = 40.0 * N + 40.0 * K_FULL + 48.0 * length
    return evals, flops, bytes_


# ----------------------------------------------------------------------- CPU sample
def cpu_sample(args):
    """Times the reference algorithm (pairwise reciprocal form, map-free port) on ion and chain moves."""
    seed, n_ion, n_chain = args
    from oracle.oracle_py import Oracle
    from plum_b200 import synth
    r, sysm, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    o = Oracle(params, repl_mode=0)
    o.upload(sysm.xyz, sysm.q, types.ids(sysm.symbol), sysm.mol_first)
    # The reference evaluates only the NEW pair energies of a trial (the old ones come from its N^2
    # std::map caches), so the port is timed in that mode: same arithmetic, same pair count, none of
    # the map overhead (measured on a 1320-bead cut of this system the real plum_ref is 2.7x slower
    # than this) — the most favourable stand-in for the reference.
    o.set_timing_new_only(True)
    rng = np.random.default_rng(seed)
    chains = [m for m in range(sysm.n_mol) if sysm.mol_first[m + 1] - sysm.mol_first[m] > 1]
    ions = [m for m in range(sysm.n_mol) if sysm.mol_first[m + 1] - sysm.mol_first[m] == 1]
    t_ion, t_chain = [], []
    for _ in range(n_ion):
        m = int(rng.choice(ions))
        f = int(sysm.mol_first[m])
        v = rng.normal(size=3)
        t0 = time.perf_counter()
        o.delta_e(m, sysm.xyz[f:f + 1] + 6.0 * v / np.linalg.norm(v), [1])
        o.commit(False)
        t_ion.append(time.perf_counter() - t0)
    for _ in range(n_chain):
        m = int(rng.choice(chains))
        f, l = int(sysm.mol_first[m]), int(sysm.mol_first[m + 1])
        t0 = time.perf_counter()
        o.delta_e(m, sysm.xyz[f:l] + rng.uniform(-1, 1, 3), np.ones(l - f, dtype=np.uint8))
        o.commit(False)
        t_chain.append(time.perf_counter() - t0)
    return t_ion, t_chain


def mix_rate(t_ion, t_chain, p_ion):
    return 1.0 / (p_ion * float(np.mean(t_ion)) + (1.0 - p_ion) * float(np.mean(t_chain)))


# ------------------------------------------------------------------- reference arm
def run_reference(a):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    cores = max(1, min(os.cpu_count() or 1, 64))
    p_ion = MOVE_PROB[0]
    budget = 150.0 / max(1, a.steps + a.warmup)           # seconds of CPU work per step and worker
    # one cheap calibration of both move kinds (also the warm-up's first step)
    t_ion0, t_chain0 = cpu_sample((1, 1, 1))
    n_ion = max(1, min(4, int(0.25 * budget / max(t_ion0[0], 1e-3))))
    n_chain = 1 if budget >= 1.2 * t_chain0[0] else 0
    t_chain_all = list(t_chain0)
    step_rates, t0_all = [], time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        for s in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            res = pool.map(cpu_sample, [(1000 * s + w, n_ion, n_chain) for w in range(cores)])
            wall = time.perf_counter() - t0
            ti = [x for r in res for x in r[0]]
            tc = [x for r in res for x in r[1]]
            t_chain_all += tc
            if s >= a.warmup:
                # `cores` independent replicas advance concurrently; per-replica rate from the mix of
                # measured per-move costs (chain-move cost from this step, or the running mean)
                rate = cores * mix_rate(ti, tc if tc else t_chain_all, p_ion)
                step_rates.append((rate, wall))
    value = float(np.mean([r for r, _ in step_rates]))
    ms_per_step = 1e3 * MOVES_PER_STEP * cores / value
    sample = (f"per step and core: {n_ion} ion move(s) + {n_chain} chain move(s) (100 beads flagged) of the same "
              f"22000-bead system evaluated with the reference's pairwise algorithm (oracle port, map-free, new-configuration "
              f"energies only like the reference); "
              f"moves/s = cores / (0.5 t_ion + 0.5 t_chain); plum_ref itself cannot hold N=22000 "
              f"(70-115 GB of std::map nodes, SURVEY.md §0.8)")
    line = {
        "impl": "reference", "metric": "MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "moves_per_step": MOVES_PER_STEP, "replicas": cores},
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_chain_value": value / cores},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0_all,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from plum_b200 import synth
    from plum_b200._abi import PgDelta  # noqa: F401
    from plum_b200.engine import Engine

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    M, K, W = a.moves_per_step, a.steps, a.warmup
    r, sysm, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    ids = types.ids(sysm.symbol)
    eng = Engine(params, device=local_rank, capacity_beads=sysm.n)
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    eng.init_energy()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)

    # ---------------- e2e leg: host-driven Metropolis loop through the C ABI (native caller)
    L = mcbench_lib()
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    xyz0 = np.ascontiguousarray(sysm.xyz)
    box = np.array(sysm.box, dtype=np.float64)
    prob = np.array(MOVE_PROB, dtype=np.float64)
    mf = np.ascontiguousarray(sysm.mol_first, dtype=np.int32)
    ctx = L.pb_create(eng.h, sysm.n_mol, mf.ctypes.data_as(ip), xyz0.ctypes.data_as(dp), box.ctypes.data_as(dp),
                      C.c_double(r.beta), C.c_double(r.move_size), C.c_double(r.rigid_bond), prob.ctypes.data_as(dp),
                      0, C.c_uint(12345 + 1000 * rank))
    n_total = (W + K) * M
    cap_beads = n_total * 100
    rec_mol = np.zeros(n_total, dtype=np.int32)
    rec_off = np.zeros(n_total, dtype=np.int32)
    rec_u = np.zeros(n_total)
    rec_dE = np.zeros(n_total)
    rec_acc = np.zeros(n_total, dtype=np.uint8)
    rec_trial = np.zeros((cap_beads, 3))
    rec_moved = np.zeros(cap_beads, dtype=np.uint8)
    used_total = 0
    e2e_wall, e2e_inner = [], []

    def run_step(s):
        nonlocal used_total
        used = C.c_int32()
        wall = C.c_double()
        ev = C.c_double()
        fl = C.c_double()
        nacc = C.c_int32()
        sl = slice(s * M, (s + 1) * M)
        rc = L.pb_run(ctx, M, rec_mol[sl].ctypes.data_as(ip), rec_off[sl].ctypes.data_as(ip),
                      rec_u[sl].ctypes.data_as(dp), rec_dE[sl].ctypes.data_as(dp), rec_acc[sl].ctypes.data_as(bp),
                      rec_trial[used_total:].ctypes.data_as(dp), rec_moved[used_total:].ctypes.data_as(bp),
                      cap_beads - used_total, C.byref(used), C.byref(wall), C.byref(ev), C.byref(fl), C.byref(nacc),
                      K_FULL)
        if rc != 0:
            raise RuntimeError(f"pb_run failed ({rc}): {eng.L.pg_last_error(eng.h).decode()}")
        rec_off[sl] += used_total
        used_total += used.value
        return wall.value

    for s in range(W):
        run_step(s)
        flush_l2()
    launches0 = eng.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    t0 = time.perf_counter()
    for s in range(W, W + K):
        e2e_inner.append(run_step(s))
        flush_l2()
    barrier()
    t_e2e = time.perf_counter() - t0
    e2e_launches = eng.launch_count() - launches0
    L.pb_destroy(ctx)
    t_e2e = max_over_ranks(t_e2e)
    final_e2e = eng.totals()

    # ---------------- value leg: the same sequence, device-resident replay
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    eng.init_energy()
    eng.replay_upload(rec_mol, rec_off, rec_u, rec_trial[:used_total], rec_moved[:used_total])
    for s in range(W):
        eng.replay_run(s * M, M)
        flush_l2()
    launches0 = eng.launch_count()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    dE_replay = []
    for s in range(W, W + K):
        dE, acc, ms = eng.replay_run(s * M, M)
        dev_ms += ms
        dE_replay.append(dE)
        flush_l2()
    barrier()
    t_wall_replay = time.perf_counter() - t0
    gpu_launches = eng.launch_count() - launches0
    clock_info = clocks.stop()
    dev_s = max_over_ranks(dev_ms * 1e-3)
    replay_matches = bool(np.array_equal(np.concatenate(dE_replay), rec_dE[W * M:]) and eng.totals() == final_e2e)

    # ---------------- roofline of the dominant kernel (k_delta), timed alone, same proposals
    timed = np.arange(W * M, (W + K) * M)
    evals, flops, bytes_ = alg_flops_per_move(sysm, rec_mol[timed])
    eng.replay_time_delta(W * M, min(M, 256))   # warm
    kd_ms = eng.replay_time_delta(W * M, K * M)
    fp64_peak_gflops = eng.measure_fp64_peak()
    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # traffic: dram bytes per k_move launch from the committed `ncu --set full` capture of this command
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "r01_ncu_summary.json")) as f:
            prof = json.load(f)["k_move_full"]
        vals = []
        for p_ in prof:
            rd, wr = p_["dram__bytes_read.sum"].split(), p_["dram__bytes_write.sum"].split()
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals.append(float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]])
        traffic = float(np.mean(vals))
    except Exception:
        traffic = None
    achieved_tflops = float(flops.sum()) / (kd_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_gflops / 1e3, "unit": "TFLOP/s",
        "frac": achieved_tflops / (fp64_peak_gflops / 1e3), "traffic": traffic,
        "traffic_note": "dram__bytes_read+write per k_move launch, ncu --set full (cold cache), profiles/r01_ncu_summary.json",
        "kernel": "k_move", "launches": int(K * M), "avg_launch_us": kd_ms * 1e3 / (K * M),
        "peak_source": "FP64 FMA microbenchmark measured in this run (pg_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flops_per_launch": float(flops.mean()),
        "bytes_view": {"bound": "hbm", "achieved": float(bytes_.sum()) / (kd_ms * 1e-3) / 1e9, "peak": hbm_peak,
                       "unit": "GB/s", "frac": float(bytes_.sum()) / (kd_ms * 1e-3) / 1e9 / hbm_peak,
                       "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                       "note": "the 0.9 MB working set is L2-resident by design; HBM is the conservative denominator"},
    }

    # ---------------- totals over ranks
    moves_total = sum_over_ranks(float(K * M))
    evals_total = sum_over_ranks(float(evals.sum()))
    value = moves_total / dev_s
    e2e_value = moves_total / t_e2e
    # bytes crossing PCIe per step: staged group block H2D (trial xyz + q + type + moved) and the result mailbox D2H
    lens = (sysm.mol_first[rec_mol[timed] + 1] - sysm.mol_first[rec_mol[timed]]).astype(np.float64)
    h2d = float((lens * (24 + 8 + 4 + 1)).sum() / K)
    d2h = float(M * 112)

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        t_ion, t_chain = cpu_sample((7, 3, 2))
        p_ion = float(np.mean(lens == 1))
        v = mix_rate(t_ion, t_chain, p_ion)
        cpu = {"value": v, "unit": "moves/s", "cores": 1, "kind": "port",
               "sample": f"3 ion moves ({np.mean(t_ion):.3f} s each) + 2 chain moves of 100 flagged beads "
                         f"({np.mean(t_chain):.3f} s each) on the same 22000-bead system, reference pairwise algorithm "
                         f"(oracle port, map-free, new-configuration energies only like the reference), mixed with the realised ion fraction {p_ion:.3f}; plum_ref cannot "
                         f"hold N=22000 (SURVEY.md §0.8)"}

    if rank == 0:
        line = {
            "metric": "MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "moves_per_step": M, "replicas_per_gpu": 1,
                       "l2": "flushed between steps (256 MiB fill, inside the timed region); within a step the "
                             "0.9 MB working set stays L2-resident by design (north_star)"},
            "pair_dE_evals_per_s": evals_total / dev_s,
            "e2e": {"value": e2e_value, "unit": "moves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e * 1e3 / K, "pair_dE_evals_per_s": evals_total / t_e2e,
                    "caller": "native C++ Metropolis loop over pg_delta_e/pg_commit (plum_b200/host/mc_bench.cc)",
                    "launches": int(e2e_launches)},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info,
            "replay_matches_e2e": replay_matches, "wall_ms_per_step_replay": t_wall_replay * 1e3 / K,
            "accept_ratio": float(rec_acc[W * M:].mean()),
        }
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--moves-per-step", type=int, default=MOVES_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
