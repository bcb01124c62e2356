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

# more replica streams than the default 8 hardware work queues would alias onto shared queues and
# serialise falsely; must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# keep stdout to the one JSON line: NCCL's version / debug lines go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

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
    L.pb_set_records.argtypes = [C.c_void_p, ip, ip, dp, dp, bp, dp, bp, C.c_int]
    L.pb_set_records.restype = None
    L.pb_stats.argtypes = [C.c_void_p, ip, dp, dp, ip]
    L.pb_stats.restype = None
    L.pb_run_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp]
    L.pb_run_mc.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, dp]
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


def plum_ref_cut(n_chains=12, steps=40):
    """The REAL reference binary (oracle/_ref/plum_ref: /root/reference/src compiled unmodified apart from the seed /
    trace hooks, oracle/build_ref.py) on one core, on a cut of S it can hold: the same box, alpha and move mix (hence the
    same cutoffs and K = 3574), 12 chains instead of 200 => N = 1320.  Its per-move cost is n_moved * N pair evaluations
    of (27 images + 3574 cos), so t = a * sum(n_moved) * N is fitted on the run and carried to N = 22 000 — reported as
    an EXTRAPOLATION next to the measured rate (profiles/r01d_cpu_scaling_plum_ref.jsonl holds N = 1320 / 2750 / 5500:
    the fit is linear in N)."""
    import tempfile
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import replay
    from plum_b200 import synth
    if not replay.have_plum_ref():
        return {"unavailable": "oracle/_ref/plum_ref not built (python oracle/build_ref.py where /root/reference exists)"}
    sysm = synth.make_system(n_chains=n_chains, chain_len=100, charged_every=10)
    with tempfile.TemporaryDirectory(prefix="plum_ref_cut_") as d:
        synth.write_inputs(d, sysm, n_steps=steps, alpha=0.004, spring=False)
        t0 = time.perf_counter()
        replay.run_plum_ref(d, 0, 1, xyz=False)                 # start-up + energy initialisation only
        t_init = time.perf_counter() - t0
        t0 = time.perf_counter()
        lines = replay.run_plum_ref(d, steps, 1, xyz=False)
        t_run = max(time.perf_counter() - t0 - t_init, 1e-9)
    T = [ln.split() for ln in lines if ln.startswith("T ")]
    n_ion = sum(1 for t in T if t[2] == "0")
    n_chain = len(T) - n_ion
    a_fit = t_run / ((n_ion + 100.0 * n_chain) * sysm.n)
    p_ion = MOVE_PROB[0]
    t_move = a_fit * 22000 * (p_ion * 1 + (1 - p_ion) * 100)
    return {"kind": "reference", "cores": 1, "N": int(sysm.n), "moves": len(T), "ion_moves": n_ion, "chain_moves": n_chain,
            "init_s": t_init, "moves_s": t_run, "measured_moves_per_s_at_this_N": len(T) / t_run,
            "extrapolated_moves_per_s_at_N22000": 1.0 / t_move,
            "note": "plum_ref itself, one core, same box / alpha / K / move mix as the workload but 12 chains; the N = 22000 "
                    "figure is EXTRAPOLATED (cost per move is linear in n_moved * N); plum_ref cannot hold N = 22000"}


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
    try:
        ref_cut = None if a.no_plum_ref else plum_ref_cut()
    except Exception as e:   # noqa: BLE001
        ref_cut = {"error": repr(e)}
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
                         "single_chain_value": value / cores, "plum_ref_cut": ref_cut},
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
    # fixed per GPU whatever N is (weak scaling); each replica has one host thread in the e2e leg
    R_auto = a.replicas_per_gpu if a.replicas_per_gpu > 0 else 30
    r, sysm, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    ids = types.ids(sysm.symbol)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)
        torch.cuda.synchronize()

    L = mcbench_lib()
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    xyz0 = np.ascontiguousarray(sysm.xyz)
    box = np.array(sysm.box, dtype=np.float64)
    prob = np.array(MOVE_PROB, dtype=np.float64)
    mf = np.ascontiguousarray(sysm.mol_first, dtype=np.int32)
    n_total = (W + K) * M

    def measure(R):
        class Replica:
            """One independent Markov chain: its own engine (stream, resident state) and random stream."""

            def __init__(self, idx):
                self.eng = Engine(params, device=local_rank, capacity_beads=sysm.n)
                self.eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
                self.eng.init_energy()
                self.stream = torch.cuda.ExternalStream(self.eng.stream(), device=torch.device("cuda", local_rank))
                self.seed = 12345 + 1000 * rank + 17 * idx
                self.ctx = self.new_ctx()
                self.cap_beads = n_total * 100
                self.rec_mol = np.zeros(n_total, dtype=np.int32)
                self.rec_off = np.zeros(n_total, dtype=np.int32)
                self.rec_u = np.zeros(n_total)
                self.rec_dE = np.zeros(n_total)
                self.rec_acc = np.zeros(n_total, dtype=np.uint8)
                self.rec_trial = np.zeros((self.cap_beads, 3))
                self.rec_moved = np.zeros(self.cap_beads, dtype=np.uint8)
                self.used_total = 0
                self.replay_dE = []
                self.replay_ms = []
                self.kd_ms = 0.0

            def new_ctx(self):
                return L.pb_create(self.eng.h, sysm.n_mol, mf.ctypes.data_as(ip), xyz0.ctypes.data_as(dp),
                                   box.ctypes.data_as(dp), C.c_double(r.beta), C.c_double(r.move_size),
                                   C.c_double(r.rigid_bond), prob.ctypes.data_as(dp), 0, C.c_uint(self.seed))

            def reset_for_mc(self):
                """Back to the initial configuration and the initial random stream: the batched leg must walk the
                very same Markov chain as the per-move leg."""
                self.eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
                self.eng.init_energy()
                self.ctx = self.new_ctx()
                self.mc_mol = np.zeros(n_total, dtype=np.int32)
                self.mc_off = np.zeros(n_total, dtype=np.int32)
                self.mc_u = np.zeros(n_total)
                self.mc_dE = np.zeros(n_total)
                self.mc_acc = np.zeros(n_total, dtype=np.uint8)

            def set_mc_records(self, s):
                sl = slice(s * M, (s + 1) * M)
                L.pb_set_records(self.ctx, self.mc_mol[sl].ctypes.data_as(ip), self.mc_off[sl].ctypes.data_as(ip),
                                 self.mc_u[sl].ctypes.data_as(dp), self.mc_dE[sl].ctypes.data_as(dp),
                                 self.mc_acc[sl].ctypes.data_as(bp), None, None, 0)

            def set_records(self, s):
                sl = slice(s * M, (s + 1) * M)
                L.pb_set_records(self.ctx, self.rec_mol[sl].ctypes.data_as(ip), self.rec_off[sl].ctypes.data_as(ip),
                                 self.rec_u[sl].ctypes.data_as(dp), self.rec_dE[sl].ctypes.data_as(dp),
                                 self.rec_acc[sl].ctypes.data_as(bp), self.rec_trial[self.used_total:].ctypes.data_as(dp),
                                 self.rec_moved[self.used_total:].ctypes.data_as(bp), self.cap_beads - self.used_total)

            def collect(self, s):
                used = C.c_int32()
                L.pb_stats(self.ctx, C.byref(used), None, None, None)
                self.rec_off[s * M:(s + 1) * M] += self.used_total
                self.used_total += used.value

            def reset_for_replay(self):
                L.pb_destroy(self.ctx)
                self.final_e2e = self.eng.totals()
                self.eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
                self.eng.init_energy()
                self.eng.replay_upload(self.rec_mol, self.rec_off, self.rec_u, self.rec_trial[:self.used_total],
                                       self.rec_moved[:self.used_total])

            def replay_step(self, s, keep):
                dE, acc, ms = self.eng.replay_run(s * M, M)
                if keep:
                    self.replay_dE.append(dE)
                    self.replay_ms.append(ms)

            def time_delta(self):
                self.kd_ms = self.eng.replay_time_delta(W * M, K * M)

        reps = [Replica(i) for i in range(R)]

        def run_all(fn):
            """Run fn(replica) for every replica concurrently (ctypes releases the GIL inside the C ABI)."""
            if R == 1:
                fn(reps[0])
                return
            errs = []

            def wrap(rp):
                try:
                    fn(rp)
                except Exception as e:   # noqa: BLE001
                    errs.append(e)
            ths = [threading.Thread(target=wrap, args=(rp,)) for rp in reps]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            if errs:
                raise errs[0]

        def device_ms(fn):
            """Device time of `fn` run on every replica concurrently: every replica stream first waits on a start
            event, the end event is recorded after all of them have joined — robust against replicas that do
            not overlap (queueing, launch skew), unlike a max over per-replica event times."""
            cur = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            for rp in reps:
                rp.stream.wait_event(e0)
            run_all(fn)
            for rp in reps:
                ej = torch.cuda.Event()
                ej.record(rp.stream)
                cur.wait_event(ej)
            e1.record(cur)
            e1.synchronize()
            return e0.elapsed_time(e1)

        # ---------------- e2e leg: host-driven Metropolis loops through the C ABI (native callers).
        # T host threads, each driving its share of the replicas through pg_delta_e_begin / _poll so that
        # their round trips overlap; every chain stays strictly sequential.
        T = max(1, min(R, a.host_threads if a.host_threads > 0 else max(1, (os.cpu_count() or 2) // world - 2)))
        groups = [reps[i::T] for i in range(T)]

        def e2e_step(s):
            errs = []

            def drive(group):
                try:
                    for rp in group:
                        rp.set_records(s)
                    arr = (C.c_void_p * len(group))(*[rp.ctx for rp in group])
                    wall = C.c_double()
                    rc = L.pb_run_multi(arr, len(group), M, C.byref(wall))
                    if rc != 0:
                        msgs = [rp.eng.L.pg_last_error(rp.eng.h).decode() for rp in group]
                        raise RuntimeError(f"pb_run_multi failed ({rc}): {msgs}")
                    for rp in group:
                        rp.collect(s)
                except Exception as e:   # noqa: BLE001
                    errs.append(e)
            if T == 1:
                drive(groups[0])
            else:
                ths = [threading.Thread(target=drive, args=(g_,)) for g_ in groups]
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()
            if errs:
                raise errs[0]

        for s in range(W):
            e2e_step(s)
            flush_l2()
        launches0 = sum(rp.eng.launch_count() for rp in reps)
        clocks = ClockSampler(local_rank)
        clocks.start()
        barrier()
        t0 = time.perf_counter()
        for s in range(W, W + K):
            e2e_step(s)
            flush_l2()
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e_launches = sum(rp.eng.launch_count() for rp in reps) - launches0
        t_e2e = max_over_ranks(t_e2e)

        # ---------------- value leg: the same sequences, device-resident replay (one CUDA graph per step)
        run_all(lambda rp: rp.reset_for_replay())
        for s in range(W):
            run_all(lambda rp: rp.replay_step(s, False))
            flush_l2()
        for s in range(W, W + K):   # instantiate the timed steps' graphs outside the timed region
            for rp in reps:
                rp.eng.replay_prepare(s * M, M, True)
        launches0 = sum(rp.eng.launch_count() for rp in reps)
        barrier()
        t0 = time.perf_counter()
        t_flush = 0.0
        step_ms = []
        for s in range(W, W + K):
            step_ms.append(device_ms(lambda rp: rp.replay_step(s, True)))
            tf = time.perf_counter()
            flush_l2()
            t_flush += time.perf_counter() - tf
        barrier()
        t_wall_replay = time.perf_counter() - t0
        gpu_launches = sum(rp.eng.launch_count() for rp in reps) - launches0
        clock_info = clocks.stop()
        dev_ms = float(sum(step_ms))   # fork/join CUDA events around all replicas of a step
        dev_s = max_over_ranks(dev_ms * 1e-3)
        replay_matches = all(bool(np.array_equal(np.concatenate(rp.replay_dE), rp.rec_dE[W * M:]) and
                                  rp.eng.totals() == rp.final_e2e) for rp in reps)


        # ---------------- roofline of the dominant kernel (k_move), timed alone, same proposals
        timed = np.arange(W * M, (W + K) * M)
        evals = flops = bytes_ = 0.0
        for rp in reps:
            e_, f_, b_ = alg_flops_per_move(sysm, rp.rec_mol[timed])
            evals += float(e_.sum()); flops += float(f_.sum()); bytes_ += float(b_.sum())
            rp.eng.replay_time_delta(W * M, min(M, 256))   # warm + instantiate
            rp.eng.replay_prepare(W * M, K * M, False)
        torch.cuda.synchronize()
        kd_ms = device_ms(lambda rp: rp.time_delta())
        fp64_peak_gflops = reps[0].eng.measure_fp64_peak()
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
            import glob
            prof_file = sorted(glob.glob(os.path.join(REPO, "profiles", "r*_ncu_summary.json")))[-1]
            with open(prof_file) as f:
                prof = json.load(f)["k_move_full"]
            vals = []
            for p_ in prof:
                rd, wr = p_["dram__bytes_read.sum"].split(), p_["dram__bytes_write.sum"].split()
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                vals.append(float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]])
            traffic = float(np.mean(vals))
        except Exception:
            traffic = None
        n_launch = R * K * M
        achieved_tflops = flops / (kd_ms * 1e-3) / 1e12
        roofline = {
            "bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak_gflops / 1e3, "unit": "TFLOP/s",
            "frac": achieved_tflops / (fp64_peak_gflops / 1e3), "traffic": traffic,
            "traffic_note": "dram__bytes_read+write per k_move launch, ncu --set full (cold cache), profiles/" + (os.path.basename(prof_file) if traffic is not None else "-"),
            "kernel": "k_move", "launches": int(n_launch), "avg_launch_us": kd_ms * 1e3 / (K * M),
            "avg_launch_note": f"CUDA-event time (fork/join over the {R} replica stream(s)) of {K * M} back-to-back k_move launches "
                               f"per replica / {K * M}; achieved = algorithmic flops of all replicas / that time",
            "peak_source": "FP64 FMA microbenchmark measured in this run (pg_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
            "algorithmic_flops_per_launch": flops / n_launch,
            "bytes_view": {"bound": "hbm", "achieved": bytes_ / (kd_ms * 1e-3) / 1e9, "peak": hbm_peak,
                           "unit": "GB/s", "frac": bytes_ / (kd_ms * 1e-3) / 1e9 / hbm_peak,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                           "note": "the 0.9 MB working set per replica is L2-resident by design; HBM is the conservative denominator"},
        }

        # ---------------- batched leg: the same chains once more, now with device-side proposals (pg_mc_*): the host
        # only draws the random stream ahead (native caller, plum_b200/host/mc_propose.h) and uploads one descriptor
        # block per batch; trial coordinates, dE, the Metropolis test and the commit never leave the device.
        mc_error = None
        try:
            run_all(lambda rp: rp.reset_for_mc())

            def mc_step(s):
                errs = []

                def drive(group):
                    try:
                        for rp in group:
                            rp.set_mc_records(s)
                        arr = (C.c_void_p * len(group))(*[rp.ctx for rp in group])
                        wall = C.c_double()
                        rc = L.pb_run_mc(arr, len(group), M, a.mc_batch, C.byref(wall))
                        if rc != 0:
                            msgs = [rp.eng.L.pg_last_error(rp.eng.h).decode() for rp in group]
                            raise RuntimeError(f"pb_run_mc failed ({rc}): {msgs}")
                    except Exception as e:   # noqa: BLE001
                        errs.append(e)
                if T == 1:
                    drive(groups[0])
                else:
                    ths = [threading.Thread(target=drive, args=(g_,)) for g_ in groups]
                    for t in ths:
                        t.start()
                    for t in ths:
                        t.join()
                if errs:
                    raise errs[0]

            for s in range(W):
                mc_step(s)
                flush_l2()
        except Exception as e:   # noqa: BLE001 — keep the other legs' numbers
            mc_error = repr(e)
        launches0 = sum(rp.eng.launch_count() for rp in reps)
        barrier()   # collectives stay outside the try blocks: a rank that failed still takes part
        t0 = time.perf_counter()
        try:
            if mc_error is None:
                for s in range(W, W + K):
                    mc_step(s)
                    flush_l2()
        except Exception as e:   # noqa: BLE001
            mc_error = repr(e)
        t_mc = (time.perf_counter() - t0) if mc_error is None else float("inf")
        barrier()
        t_mc = max_over_ranks(t_mc)
        mc_launches, mc_matches, mc_close = 0, False, False
        try:
            if mc_error is not None:
                raise RuntimeError(mc_error)
            mc_launches = sum(rp.eng.launch_count() for rp in reps) - launches0
            mc_matches = all(bool(np.array_equal(rp.mc_mol, rp.rec_mol) and np.array_equal(rp.mc_acc, rp.rec_acc) and
                                  np.array_equal(rp.mc_dE, rp.rec_dE) and rp.eng.totals() == rp.final_e2e) for rp in reps)
            mc_close = all(bool(np.array_equal(rp.mc_mol, rp.rec_mol) and np.array_equal(rp.mc_acc, rp.rec_acc) and
                                np.all(np.abs(rp.mc_dE - rp.rec_dE) <= 1e-10 * np.maximum(1.0, np.abs(rp.rec_dE)))) for rp in reps)
            for rp in reps:
                L.pb_destroy(rp.ctx)
        except Exception as e:   # noqa: BLE001
            mc_error = mc_error or repr(e)

        # ---------------- totals over ranks
        moves_total = sum_over_ranks(float(R * K * M))
        evals_total = sum_over_ranks(evals)
        value = moves_total / dev_s
        e2e_value = moves_total / t_e2e
        # bytes crossing PCIe per step: staged group block H2D (trial xyz + q + type + index + moved) and the mailbox D2H
        lens = np.concatenate([(sysm.mol_first[rp.rec_mol[timed] + 1] - sysm.mol_first[rp.rec_mol[timed]]) for rp in reps]).astype(np.float64)
        h2d = float((lens * (24 + 8 + 4 + 4 + 1)).sum() / K)
        d2h = float(R * M * 128)
        rec_acc_all = np.concatenate([rp.rec_acc[W * M:] for rp in reps])
        # batched path: 72 B descriptor per move + 21 B per bead of the moved molecule (charge, type, index, flag)
        # + 32 B per pivot step H2D; 9 B per move (dE, accept bit) + the stop word per batch D2H
        chain_moves = lens > 1
        p_piv = MOVE_PROB[2] / max(MOVE_PROB[1] + MOVE_PROB[2] + MOVE_PROB[4], 1e-12)   # pivots among the chain moves (rows are not recorded)
        mc_h2d = float((72.0 * lens.size + 21.0 * lens.sum() + p_piv * 32.0 * (lens[chain_moves] - 1).sum()) / K)
        mc_d2h = float(R * M * 9 + R * (M // max(a.mc_batch, 1) + 1) * 4)

        for rp in reps:
            rp.eng.close()
        return dict(R=R, T=T, value=value, e2e_value=e2e_value, dev_s=dev_s, t_e2e=t_e2e, evals_total=evals_total,
                    e2e_launches=e2e_launches, gpu_launches=gpu_launches, roofline=roofline, clock_info=clock_info,
                    replay_matches=replay_matches, wall_replay=(t_wall_replay - t_flush), h2d=h2d, d2h=d2h,
                    accept=float(rec_acc_all.mean()), p_ion=float(np.mean(lens == 1)),
                    mc_value=moves_total / t_mc, t_mc=t_mc, mc_launches=mc_launches, mc_matches=mc_matches, mc_close=mc_close,
                    mc_h2d=mc_h2d, mc_d2h=mc_d2h, mc_error=mc_error)

    single = measure(1) if (R_auto > 1 and not a.no_single) else None
    res = measure(R_auto)
    R = res["R"]
    value, e2e_value, dev_s, t_e2e, evals_total = res["value"], res["e2e_value"], res["dev_s"], res["t_e2e"], res["evals_total"]
    e2e_launches, gpu_launches, roofline, clock_info = res["e2e_launches"], res["gpu_launches"], res["roofline"], res["clock_info"]
    replay_matches, t_wall_replay, t_flush, h2d, d2h = res["replay_matches"], res["wall_replay"], 0.0, res["h2d"], res["d2h"]
    # What the kernel EXECUTES next to what the reference's algorithm would (`achieved` prices every pair at the reference's
    # 36 flop; the kernel culls most of them): executed FP64 flop per launch = dadd + dmul + 2 dfma thread instructions
    # (smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on) of the `ncu --set full` capture of this command,
    # gpurun_out/prof_r01d_kmove.ncu-rep, summarised in profiles/r01_summary.md: 100-bead chain move 5.80e6, ion move 1.21e6.
    try:
        exe_per_launch = res["p_ion"] * 1.21e6 + (1.0 - res["p_ion"]) * 5.80e6
        exe_tflops = roofline["achieved"] * exe_per_launch / roofline["algorithmic_flops_per_launch"]
        roofline["executed_view"] = {
            "achieved": exe_tflops, "peak": roofline["peak"], "unit": "TFLOP/s", "frac": exe_tflops / roofline["peak"],
            "executed_flops_per_launch": exe_per_launch,
            "note": "executed FP64 flop per launch from ncu (chain move 5.80e6, ion move 1.21e6, mixed with this run's ion "
                    "fraction) over the same CUDA-event time: the kernel is bound by CTA-slot time and instruction issue, "
                    "not by the FP64 pipe; `frac` above exceeds 1 because the algorithmic count includes the culled pairs"}
    except Exception as e:   # noqa: BLE001
        roofline["executed_view"] = {"error": repr(e)}

    # ---------------- k-sharded full S(k) recompute: every rank holds the positions, fills its slice of
    # the k list with k_sk_slice, NCCL all-gathers the slices; time per rank with the collective broken out, for S and for
    # the stress variant S-full (where there is enough work per rank to be worth sharding).  Not part of `value`.
    recompute = None
    if not a.no_recompute:
        recompute = {}
        from plum_b200 import sharded
        for name, loader in (("S", synth.load), ("S_full", synth.load_full)):
            eng, setup_err = None, None
            try:
                _, s2, ty2, par2 = loader(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
                eng = Engine(par2, device=local_rank, capacity_beads=s2.n)
                eng.upload(s2.xyz, s2.q, ty2.ids(s2.symbol), s2.mol_first)
                t_init = eng.init_energy()
                l0 = eng.launch_count()
            except Exception as e:   # noqa: BLE001
                setup_err = repr(e)
            # a rank that failed to set up must not leave the others waiting in the all-gather
            if max_over_ranks(1.0 if setup_err else 0.0) > 0.0:
                recompute[name] = {"error": setup_err or "set-up failed on another rank"}
                if eng is not None:
                    eng.close()
                continue
            try:
                recompute[name] = dict(n_charged=int(np.count_nonzero(s2.q)),
                                       **sharded.time_sharded_recompute(eng, rank, world, t_init["recip"]))
                recompute[name]["launches"] = int(eng.launch_count() - l0)
            except Exception as e:   # noqa: BLE001
                recompute[name] = {"error": repr(e)}
            finally:
                if eng is not None:
                    eng.close()

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        t_ion, t_chain = cpu_sample((7, 6, 2))
        p_ion = res["p_ion"]
        v = mix_rate(t_ion, t_chain, p_ion)
        cpu = {"value": v, "unit": "moves/s", "cores": 1, "kind": "port",
               "sample": f"6 ion moves ({np.mean(t_ion):.3f} s each) + 2 chain moves of 100 flagged beads "
                         f"({np.mean(t_chain):.3f} s each) on the same 22000-bead system, reference pairwise algorithm "
                         f"(oracle port, map-free, new-configuration energies only like the reference), mixed with the realised ion fraction {p_ion:.3f}; plum_ref cannot "
                         f"hold N=22000 (SURVEY.md §0.8)"}

    e2e_per_move = {"value": e2e_value, "unit": "moves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e * 1e3 / K, "pair_dE_evals_per_s": evals_total / t_e2e,
                    "caller": f"native C++ Metropolis loops over pg_delta_e_begin/_poll/pg_commit (plum_b200/host/mc_bench.cc): {res['T']} host thread(s) per GPU driving {R} replicas; trial coordinates built on the host",
                    "launches": int(e2e_launches)}
    e2e_batched = {"value": res["mc_value"], "unit": "moves/s", "h2d_bytes_per_step": res["mc_h2d"], "d2h_bytes_per_step": res["mc_d2h"],
                   "ms_per_step": res["t_mc"] * 1e3 / K, "pair_dE_evals_per_s": evals_total / res["t_mc"],
                   "caller": f"native C++ caller over pg_mc_upload/_begin/_end (plum_b200/host/mc_bench.cc), batches of {a.mc_batch} steps: the host draws the "
                             f"std::mt19937 stream ahead in the reference's order and uploads descriptors; proposals (k_propose), dE (k_move), Metropolis test and "
                             f"commit stay on the device; {res['T']} host thread(s) per GPU driving {R} replicas; wall clock incl. generation, uploads and result downloads",
                   "launches": int(res["mc_launches"]),
                   "same_chain_as_per_move": {"bit_identical": res["mc_matches"], "within_1e-10": res["mc_close"]}, "error": res["mc_error"]}
    e2e_best = e2e_batched if (res["mc_close"] and res["mc_value"] > e2e_value) else e2e_per_move
    if rank == 0:
        line = {
            "metric": "MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "moves_per_step": M, "replicas_per_gpu": R,
                       "l2": "flushed between steps (256 MiB fill, inside the timed region); within a step the "
                             "0.9 MB working set stays L2-resident by design (north_star)"},
            "pair_dE_evals_per_s": evals_total / dev_s,
            "e2e": e2e_best,
            "e2e_per_move": e2e_per_move, "e2e_batched": e2e_batched,
            "gpu_launches": int(gpu_launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info,
            "replay_matches_e2e": replay_matches, "wall_ms_per_step_replay": (t_wall_replay - t_flush) * 1e3 / K,
            "accept_ratio": res["accept"],
            "sharded_recompute": recompute,
            "single_replica": None if single is None else {
                "value": single["value"], "e2e": max(single["e2e_value"], single["mc_value"] if single["mc_close"] else 0.0),
                "e2e_per_move": single["e2e_value"], "e2e_batched": single["mc_value"], "batched_same_chain": single["mc_close"], "unit": "moves/s", "roofline_frac": single["roofline"]["frac"],
                "avg_launch_us": single["roofline"]["avg_launch_us"], "replay_matches_e2e": single["replay_matches"],
                "note": "one Markov chain on the GPU (latency-bound): the rate a single Plum run sees"},
        }
        print(json.dumps(line))
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
    ap.add_argument("--no-single", action="store_true", help="skip the extra single-replica measurement")
    ap.add_argument("--no-recompute", action="store_true", help="skip the k-sharded full S(k) recompute timing")
    ap.add_argument("--no-plum-ref", action="store_true", help="reference arm: skip the real plum_ref binary's leg (1320-bead cut)")
    ap.add_argument("--replicas-per-gpu", type=int, default=0,
                    help="independent Markov chains per GPU, each with its own engine/stream; 0 = 30 (fixed per GPU: weak scaling; one stream each, below the 32 hardware queues)")
    ap.add_argument("--mc-batch", type=int, default=256, help="steps per uploaded batch in the batched (device-side proposal) leg")
    ap.add_argument("--host-threads", type=int, default=0,
                    help="host threads per GPU driving the replicas in the e2e leg (0 = host cores per GPU - 2, at most one per replica)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
