#!/usr/bin/env python3
"""bench.py — MC moves/s of the per-trial-move energy path on the synthetic 22k-bead polyelectrolyte
(BASELINE.json configs[4], SURVEY.md §8(d) "S").

  python bench.py --gpus N --steps K --warmup W            this repo (CUDA engine via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU algorithm on the host cores

One *step* = MOVES_PER_STEP sequential Metropolis steps (move mix 0.5 ion translation / 0.1 COM / 0.3 pivot /
0.1 reptation) of EVERY replica of the GPU.  Replicas are independent Markov chains (own engine, own random stream);
there is no communication: scaling "weak", `value` is the aggregate over all ranks.

  value   THROUGHPUT mode: `replicas_per_gpu` chains per GPU (default two per SM, one CTA each: the 448-thread build of k_chain at 2 CTAs per SM), all of them in ONE launch
          of the device-resident chain kernel k_chain per step (pg_chain_run_multi); generator state, coordinates, S(k),
          cell grid already resident; CUDA-event time of the launches.  The rate ONE Markov chain sees is in
          `single_chain` (a 16-CTA cluster per chain), next to `value`, for both pivot modes.
  e2e     the same chains through the reference-facing call sequence with HOST buffers, one host thread per GPU: per
          batch the std::mt19937 state goes down, one launch, and the step log, the generator state and the accepted
          coordinates (what ForceField::TranslationalBatch hands back to the driver's beads) come back; wall clock.
  per_move   the round-1 path for comparison: pg_delta_e / pg_commit with host-built trial coordinates (k_move).
  roofline   dominant kernel k_chain.  `frac` = EXECUTED FP64 flop (dfma x2 + dadd + dmul thread instructions of the tracked
          ncu capture profiles/r02_ncu_summary.json, per step) x measured steps/s / the FP64 FMA peak measured in this
          run.  Also: issue-slot fraction (warp instructions/s over SMs x 4 x clock), an L2 view against an L2 streaming
          peak measured in this run, the "necessary work" count (pair configurations inside a cutoff, counted by the
          kernel) and the SURVEY §8(d) algorithmic figure, labelled as such.
  cpu_baseline   the CPU oracle port (the reference's pairwise algorithm, oracle/plum_oracle.c) on 1 host core, bounded sample.
  examples   bin/plum_gpu (the unchanged reference driver over the façade) on the four reference examples, 10^4 steps.
"""
import argparse
import ctypes as C
import glob
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import threading
import time

# keep stdout to the one JSON line: NCCL's version / debug lines go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
EXAMPLES = ["bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"]
EXAMPLE_STEPS = 10000          # the reference binary: 30-70 s per example on one core
EXAMPLE_STEPS_GPU = 100000     # bin/plum_gpu: long enough to stand out from the process start-up


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
    vpp = C.POINTER(C.c_void_p)
    L.pb_create.restype = C.c_void_p
    L.pb_create.argtypes = [C.c_void_p, C.c_int, ip, dp, dp, C.c_double, C.c_double, C.c_double, dp, C.c_int, C.c_uint]
    L.pb_destroy.argtypes = [C.c_void_p]
    L.pb_set_records.argtypes = [C.c_void_p, ip, ip, dp, dp, bp, dp, bp, C.c_int]
    L.pb_set_records.restype = None
    L.pb_stats.argtypes = [C.c_void_p, ip, dp, dp, ip]
    L.pb_stats.restype = None
    L.pb_run_multi.argtypes = [vpp, C.c_int, C.c_int, dp]
    L.pb_run_chain.argtypes = [vpp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    L.pb_chain_seed.argtypes = [vpp, C.c_int, C.c_int]
    L.pb_chain_resident.argtypes = [vpp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.pb_set_pivot_mode.argtypes = [C.c_void_p, C.c_int]
    L.pb_set_pivot_mode.restype = None
    return L


def alg_flops_per_move(n, first, q, mols):
    """flops = 2*evals*F_pair + 2*n_moved_q*K*16 + 12*K; bytes = 40*N + 40*K + 48*n_moved."""
    length = (first[mols + 1] - first[mols]).astype(np.float64)
    cs = np.concatenate([[0.0], np.cumsum(q != 0)])
    nq = cs[first[mols + 1]] - cs[first[mols]]
    evals = length * (n - length) + 0.5 * length * (length - 1)
    flops = 2 * evals * F_PAIR + 2 * nq * K_FULL * 16 + 12 * K_FULL
    return evals, flops, nq


# ----------------------------------------------------------------------- CPU sample
def cpu_sample(args):
    """Times the reference algorithm (pairwise reciprocal form, map-free port) on ion and chain moves of S."""
    seed, n_ion, n_chain = args
    from oracle.oracle_py import Oracle
    from plum_b200 import synth
    r, sysm, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    o = Oracle(params, repl_mode=0)
    o.upload(sysm.xyz, sysm.q, types.ids(sysm.symbol), sysm.mol_first)
    # The reference evaluates only the NEW pair energies of a trial (the old ones come from its N^2 std::map caches), so
    # the port is timed in that mode: same arithmetic, same pair count, none of the map overhead (on a 1320-bead cut the
    # real plum_ref is 2.7x slower than this) — the most favourable stand-in for the reference.
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


# ------------------------------------------------------------------ the real reference binary
def _replay():
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import replay
    return replay


def plum_ref_cut(args):
    """oracle/_ref/plum_ref ITSELF on a cut of S it can hold (same box, alpha, move mix, K = 3574; fewer chains), one core:
    (N, moves, ion moves, chain moves, init seconds, move seconds)."""
    n_chains, steps, core = args
    replay = _replay()
    from plum_b200 import synth
    if core is not None:
        try:
            os.sched_setaffinity(0, {core})
        except Exception:
            pass
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
    return {"N": int(sysm.n), "moves": len(T), "ion_moves": n_ion, "chain_moves": len(T) - n_ion, "init_s": t_init, "moves_s": t_run}


def run_example(args):
    """One reference example through a driver binary (plum_ref or bin/plum_gpu): wall seconds of `steps` steps with the
    sampling / output frequencies pushed beyond the run (only the move path is timed), minus the start-up of a 0-step run."""
    name, binary, steps, core, extra_env = args
    replay = _replay()
    if core is not None:
        try:
            os.sched_setaffinity(0, {core})
        except Exception:
            pass
    d = os.path.join(replay.GOLDEN, "examples", name)
    off = {"s1_equilibrium_steps": 1000000000, "s1_sampling_frequency": 1000000000, "s1_sampling_print_frequency": 1000000000,
           "s1_trajectory_print_frequency": 1000000000}
    # start-up (process, CUDA context, energy initialisation) is taken out with a short run of the same binary: the
    # difference of a `base`-step run and a `base + steps`-step run is the time of `steps` steps
    base = steps // 10
    t_init = float("inf")
    for _ in range(2):   # (the first start of a CUDA process on a box is slower: take the faster of two)
        t0 = time.perf_counter()
        replay.run_plum_ref(d, base, 1, xyz=False, binary=binary, overrides=off, extra_env=extra_env, trace=False)
        t_init = min(t_init, time.perf_counter() - t0)
    err = []
    prof_env = dict(extra_env or {})
    if binary == replay.PLUM_GPU:
        prof_env["PLUM_B200_PROFILE"] = "1"     # wall time inside each ABI entry point / GC attempt, printed at exit
    t0 = time.perf_counter()
    replay.run_plum_ref(d, base + steps, 1, xyz=False, binary=binary, overrides=off, extra_env=prof_env, stderr_to=err, trace=False)
    t_all = time.perf_counter() - t0
    # which steps were translational / grand-canonical: a traced run of the first `base` steps (not timed)
    lines = replay.run_plum_ref(d, base, 1, xyz=False, binary=binary, overrides=off, extra_env=extra_env)
    n_gc = sum(1 for ln in lines if ln.startswith("G ")) * steps // max(base, 1)
    n_t = sum(1 for ln in lines if ln.startswith("T ")) * steps // max(base, 1)
    t_run = max(t_all - t_init, 1e-9)
    sites = {}
    for ln in "".join(err).split("\n"):
        if ln.startswith("plum_b200 profile: "):
            t = ln.split()
            sites[t[2]] = {"calls": int(t[4]), "total_s": float(t[6]), "us_per_call": float(t[8])}
    out = {"example": name, "steps": steps, "translational_steps": n_t, "gc_steps": n_gc, "startup_s": t_init, "run_s": t_run,
           "steps_per_s": steps / t_run}
    if "run_wall" in sites:
        # bin/plum_gpu times its own run loop (end of the energy initialisation to exit): CUDA start-up varies by seconds
        # from process to process, which the differencing above cannot take out of a 2 s run
        w = sites.pop("run_wall")["total_s"]
        out.update({"steps": base + steps, "run_s": w, "steps_per_s": (base + steps) / w, "startup_s": t_all - w,
                    "run_s_by_differencing": t_run,
                    "timed": "wall time of the driver's run loop measured inside the binary (PLUM_B200_PROFILE), all steps of the long run"})
        n_gc = n_gc * (base + steps) // steps
        n_t = n_t * (base + steps) // steps
        out["translational_steps"], out["gc_steps"] = n_t, n_gc
    if sites:
        gc_s = sum(sites[k]["total_s"] for k in ("CBMCFChainInsertion", "CBMCFChainDeletion") if k in sites)
        gc_n = sum(sites[k]["calls"] for k in ("CBMCFChainInsertion", "CBMCFChainDeletion") if k in sites)
        out["facade_sites"] = sites      # of the (base + steps)-step run
        if gc_n:
            out["gc_attempts_per_s"] = gc_n / gc_s
            out["gc_us_per_attempt"] = 1e6 * gc_s / gc_n
        out["timing_note"] = "PLUM_TRACE off in the timed runs (a trace line reads the four totals every step); step kinds counted in a separate traced run"
    return out


def run_example_own_frequencies(name, steps):
    """bin/plum_gpu on one example with the example's OWN sampling / print frequencies (SURVEY.md §8d asks for both):
    Sample(), the pressure samplers and the writers run as shipped; timed by the binary's own run-loop clock."""
    replay = _replay()
    err = []
    replay.run_plum_ref(os.path.join(replay.GOLDEN, "examples", name), steps, 1, xyz=False, binary=replay.PLUM_GPU,
                        extra_env={"PLUM_B200_PROFILE": "1"}, stderr_to=err, trace=False)
    for ln in "".join(err).split("\n"):
        if ln.startswith("plum_b200 profile: run_wall "):
            w = float(ln.split()[6])
            return {"steps": steps, "run_s": w, "steps_per_s": steps / w}
    return {"error": "no run_wall line"}


def examples_block(binary, cores, extra_env=None, parallel=True, steps=EXAMPLE_STEPS):
    replay = _replay()
    if not os.path.exists(binary):
        return {"unavailable": f"{os.path.relpath(binary, REPO)} not built"}
    jobs = [(name, binary, steps, (cores[i % len(cores)] if cores else None), extra_env) for i, name in enumerate(EXAMPLES)]
    try:
        if parallel:
            with mp.get_context("fork").Pool(len(jobs)) as pool:
                res = pool.map(run_example, jobs)
        else:
            res = [run_example(j) for j in jobs]
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)}
    return {r["example"]: {k: v for k, v in r.items() if k != "example"} for r in res}


# ------------------------------------------------------------------- reference arm
def run_reference(a):
    """The reference's own CPU implementation of the path on this box's host cores.  plum_ref cannot hold S (70-115 GB of
    std::map nodes), so the line's `value` is what was actually MEASURED per step: every host core evaluates
    one ion move and one chain move (the workload's 50/50 mix) of the 22 000-bead system with the oracle port of the
    reference's pairwise algorithm; value = moves evaluated / wall time of the step.  Beside it: plum_ref itself on three
    cuts of S (1320 / 2750 / 5500 beads) with the per-core extrapolation to N = 22 000 (labelled), and plum_ref on the four
    reference examples (10^4 steps each, one core each)."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    t0_all = time.perf_counter()
    cores_avail = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    cores = max(1, min(len(cores_avail), 64))
    p_ion = MOVE_PROB[0]
    step_walls, t_ion_all, t_chain_all = [], [], []
    with mp.get_context("fork").Pool(cores) as pool:
        for s in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            res = pool.map(cpu_sample, [(1000 * s + w, 1, 1) for w in range(cores)])
            wall = time.perf_counter() - t0
            if s >= a.warmup:
                step_walls.append(wall)
                t_ion_all += [x for r_ in res for x in r_[0]]
                t_chain_all += [x for r_ in res for x in r_[1]]
    moves_per_step = 2 * cores
    ms_per_step = 1e3 * float(np.mean(step_walls)) if step_walls else float("nan")
    value = moves_per_step / (ms_per_step * 1e-3)
    per_core = mix_rate(t_ion_all, t_chain_all, p_ion) if t_ion_all else float("nan")
    ref_cuts, ex = None, None
    replay = _replay()
    if not a.no_plum_ref and replay.have_plum_ref():
        try:
            jobs = [(12, 40, cores_avail[0 % cores]), (25, 30, cores_avail[1 % cores]), (50, 24, cores_avail[2 % cores])]
            with mp.get_context("fork").Pool(len(jobs)) as pool:
                cuts = pool.map(plum_ref_cut, jobs)
            # cost per move is linear in n_moved * N (pair evaluations of 27 images + 3574 cos): fit t = a * sum(n_moved) * N
            num = sum(c["moves_s"] * (c["ion_moves"] + 100.0 * c["chain_moves"]) * c["N"] for c in cuts)
            den = sum(((c["ion_moves"] + 100.0 * c["chain_moves"]) * c["N"]) ** 2 for c in cuts)
            a_fit = num / den
            t_move = a_fit * 22000 * (p_ion * 1 + (1 - p_ion) * 100)
            ref_cuts = {"kind": "reference", "cores_per_run": 1, "cuts": cuts,
                        "fit_seconds_per_moved_bead_and_partner": a_fit,
                        "extrapolated_moves_per_s_per_core_at_N22000": 1.0 / t_move,
                        "note": "oracle/_ref/plum_ref itself on cuts of S with the same box / alpha / K / move mix; the N = 22000 "
                                "figure is EXTRAPOLATED (least-squares fit linear in n_moved x N over the cuts) and is per core; the "
                                "N = 11000 cut (29 GB of map nodes, 4 min of initialisation) is in profiles/r02_cpu_scaling_plum_ref.jsonl"}
        except Exception as e:   # noqa: BLE001
            ref_cuts = {"error": repr(e)}
        if not a.no_examples:
            ex = examples_block(replay.PLUM_REF, cores_avail)
    sample = (f"per step, each of {cores} host cores evaluates 1 ion move + 1 chain move (100 beads flagged) of the same 22000-bead "
              f"system with the reference's pairwise algorithm (oracle port, map-free, new-configuration energies only like the "
              f"reference): {moves_per_step} moves per step, value = moves / measured wall time of the step")
    line = {
        "impl": "reference", "metric": "MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "moves_per_step": moves_per_step, "replicas": cores},
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": cores, "kind": "port", "sample": sample,
                         "per_core_moves_per_s": per_core, "plum_ref_cuts": ref_cuts},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "examples": ex, "gpu_launches": 0, "wall_s": time.perf_counter() - t0_all,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------- our arm
def load_profile_summary():
    """Executed-work counters of k_chain from the tracked ncu capture (profiles/r*_ncu_summary.json, newest round)."""
    files = sorted(glob.glob(os.path.join(REPO, "profiles", "r*_ncu_summary.json")))
    for f in reversed(files):
        try:
            with open(f) as fh:
                j = json.load(fh)
            if "k_chain_fleet" in j:     # the whole fleet (two chains per SM) captured as ONE launch: device-wide counters
                return os.path.basename(f), dict(j["k_chain_fleet"], section="k_chain_fleet")
            if "k_chain" in j:
                return os.path.basename(f), dict(j["k_chain"], section="k_chain")
        except Exception:
            continue
    return None, None


def run_ours(a):
    import torch
    import torch.distributed as dist
    from plum_b200 import synth
    from plum_b200.engine import Engine

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        # (NCCL's default watchdog time: rank 0 alone runs the CPU baseline legs, the others wait in the next collective)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=600))

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
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    R_auto = a.replicas_per_gpu if a.replicas_per_gpu > 0 else 2 * n_sm   # two chains per SM: the 448-thread build of k_chain
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

    class Fleet:
        """R independent Markov chains on this GPU: one engine (resident state) and one std::mt19937 each."""

        def __init__(self, R, record=True):
            self.R = R
            self.engs, self.ctxs, self.rec = [], [], []
            for i in range(R):
                e = Engine(params, device=local_rank, capacity_beads=sysm.n)
                self.engs.append(e)
                self.rec.append(dict(mol=np.zeros(n_total, dtype=np.int32), off=np.zeros(n_total, dtype=np.int32), u=np.zeros(n_total),
                                     dE=np.zeros(n_total), acc=np.zeros(n_total, dtype=np.uint8)) if record else None)
            self.reset()

        def seed(self, i):
            return 12345 + 1000 * rank + 17 * i

        def reset(self, pivot_mode=0):
            """Back to the initial configuration and the initial random streams."""
            for c in self.ctxs:
                L.pb_destroy(c)
            self.ctxs = []
            for i, e in enumerate(self.engs):
                e.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
                e.init_energy()
                c = L.pb_create(e.h, sysm.n_mol, mf.ctypes.data_as(ip), xyz0.ctypes.data_as(dp), box.ctypes.data_as(dp),
                                C.c_double(r.beta), C.c_double(r.move_size), C.c_double(r.rigid_bond), prob.ctypes.data_as(dp), 0,
                                C.c_uint(self.seed(i)))
                L.pb_set_pivot_mode(c, pivot_mode)
                self.ctxs.append(c)
            self.arr = (C.c_void_p * self.R)(*self.ctxs)

        def set_records(self, s, which="rec"):
            sl = slice(s * M, (s + 1) * M)
            for c, rc in zip(self.ctxs, getattr(self, which)):
                if rc is None:
                    L.pb_set_records(c, None, None, None, None, None, None, None, 0)
                else:
                    L.pb_set_records(c, rc["mol"][sl].ctypes.data_as(ip), rc["off"][sl].ctypes.data_as(ip), rc["u"][sl].ctypes.data_as(dp),
                                     rc["dE"][sl].ctypes.data_as(dp), rc["acc"][sl].ctypes.data_as(bp), None, None, 0)

        def check(self, rc, what):
            if rc != 0:
                msgs = sorted({e.L.pg_last_error(e.h).decode() for e in self.engs})
                raise RuntimeError(f"{what} failed ({rc}): {msgs}")

        def e2e_step(self, s, cluster, batch):
            self.set_records(s)
            wall = C.c_double()
            self.check(L.pb_run_chain(self.arr, self.R, M, batch, cluster, 0, 1, C.byref(wall)), "pb_run_chain")
            return wall.value

        def resident_seed(self, cluster):
            self.check(L.pb_chain_seed(self.arr, self.R, cluster), "pb_chain_seed")

        def resident_step(self, s, book):
            if book:
                self.set_records(s, "rec2")
            ms = C.c_float()
            self.check(L.pb_chain_resident(self.arr, self.R, M, 1 if book else 0, C.byref(ms)), "pb_chain_resident")
            return ms.value

        def launches(self):
            return sum(e.launch_count() for e in self.engs)

        def close(self):
            for c in self.ctxs:
                L.pb_destroy(c)
            for e in self.engs:
                e.close()

    def chain_legs(R, cluster, pivot_mode, batch, clocks=None):
        """e2e and device-resident legs of R chains with `cluster` CTAs each."""
        fl = Fleet(R)
        fl.reset(pivot_mode)
        for s in range(W):
            fl.e2e_step(s, cluster, batch)
            flush_l2()
        if clocks:
            clocks.start()
        barrier()
        t0 = time.perf_counter()
        t_flush = 0.0
        for s in range(W, W + K):
            fl.e2e_step(s, cluster, batch)
            tf = time.perf_counter()
            flush_l2()
            t_flush += time.perf_counter() - tf
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0 - t_flush)
        final_e2e = [e.totals() for e in fl.engs]
        # device-resident: same seeds, same chains
        fl.rec2 = [dict(mol=np.zeros(n_total, dtype=np.int32), off=np.zeros(n_total, dtype=np.int32), u=np.zeros(n_total),
                        dE=np.zeros(n_total), acc=np.zeros(n_total, dtype=np.uint8)) for _ in range(R)]
        fl.reset(pivot_mode)
        fl.resident_seed(cluster)
        for s in range(W):
            fl.resident_step(s, True)
            flush_l2()
        for e in fl.engs:
            e.chain_counters(reset=True)
        l0 = fl.launches()
        barrier()
        step_ms = []
        t0 = time.perf_counter()
        for s in range(W, W + K):
            step_ms.append(fl.resident_step(s, True))
            flush_l2()
        barrier()
        wall_res = time.perf_counter() - t0
        launches = fl.launches() - l0
        dev_s = max_over_ranks(float(sum(step_ms)) * 1e-3)
        cnt = np.array([e.chain_counters() for e in fl.engs], dtype=np.float64).sum(axis=0)
        same = all(bool(np.array_equal(a_["mol"], b_["mol"]) and np.array_equal(a_["acc"], b_["acc"]) and np.array_equal(a_["dE"], b_["dE"]))
                   for a_, b_ in zip(fl.rec, fl.rec2)) and all(e.totals() == f_ for e, f_ in zip(fl.engs, final_e2e))
        timed = slice(W * M, (W + K) * M)
        mols = np.concatenate([rc["mol"][timed] for rc in fl.rec])
        evals, flops, nq = alg_flops_per_move(sysm.n, sysm.mol_first, sysm.q, mols)
        lens = (sysm.mol_first[mols + 1] - sysm.mol_first[mols])
        acc = float(np.concatenate([rc["acc"][timed] for rc in fl.rec]).mean())
        out = dict(R=R, cluster=cluster, pivot_mode=pivot_mode, dev_s=dev_s, t_e2e=t_e2e, launches=launches, same=same,
                   evals=float(evals.sum()), alg_flops=float(flops.sum()), counters=cnt.tolist(), accept=acc,
                   p_ion=float(np.mean(lens == 1)), wall_res=wall_res, rec0=fl.rec[0], step_ms=step_ms,
                   batches_per_step=(M + batch - 1) // batch)
        out["peaks"] = None
        if clocks is not None:
            out["peaks"] = (fl.engs[0].measure_fp64_peak(), fl.engs[0].measure_l2_peak())
        fl.close()
        return out

    def per_move_legs(R):
        """The round-1 path for comparison: host-built trials through pg_delta_e_begin/_poll/pg_commit (k_move), one thread."""
        fl = Fleet(R)
        wall = C.c_double()
        for s in range(W):
            fl.set_records(s)
            fl.check(L.pb_run_multi(fl.arr, R, M, C.byref(wall)), "pb_run_multi")
        barrier()
        t0 = time.perf_counter()
        for s in range(W, W + K):
            fl.set_records(s)
            fl.check(L.pb_run_multi(fl.arr, R, M, C.byref(wall)), "pb_run_multi")
        barrier()
        t = max_over_ranks(time.perf_counter() - t0)
        rec0 = fl.rec[0]
        fl.close()
        return dict(t=t, rec0=rec0)

    # ---------------- throughput mode (the headline) and one chain per GPU
    clocks = ClockSampler(local_rank)
    main = chain_legs(R_auto, a.cluster, a.pivot_mode, a.batch, clocks)
    clock_info = clocks.stop()
    R = main["R"]
    moves_total = sum_over_ranks(float(R * K * M))
    value = moves_total / main["dev_s"]
    e2e_value = moves_total / main["t_e2e"]

    single = None
    if not a.no_single:
        single = {}
        for pm in (0, 2, 1):
            sc = chain_legs(1, a.single_cluster, pm, a.batch)
            single[f"pivot_mode_{pm}"] = {
                "value": float(K * M) / sc["dev_s"], "e2e": float(K * M) / sc["t_e2e"], "us_per_step": sc["dev_s"] * 1e6 / (K * M),
                "resident_matches_e2e": sc["same"], "accept_ratio": sc["accept"]}
            if pm == 0:
                sc0 = sc
        pmv = per_move_legs(1)
        n_cmp = min(len(sc0["rec0"]["mol"]), len(pmv["rec0"]["mol"]))
        a_, b_ = sc0["rec0"], pmv["rec0"]
        big = b_["dE"][:n_cmp] >= 1e8
        same_chain = bool(np.array_equal(a_["mol"][:n_cmp], b_["mol"][:n_cmp]) and np.array_equal(a_["acc"][:n_cmp], b_["acc"][:n_cmp]) and
                          np.array_equal(a_["dE"][:n_cmp] >= 1e8, big))
        ddE = np.abs(a_["dE"][:n_cmp][~big] - b_["dE"][:n_cmp][~big]) / np.maximum(1.0, np.abs(b_["dE"][:n_cmp][~big]))
        single["per_move_path"] = {"e2e": float(K * M) / pmv["t"], "unit": "moves/s",
                                   "note": "round-1 path: pg_delta_e / pg_commit round trip per move, host-built trial coordinates (k_move)"}
        single["chain_vs_per_move_same_chain"] = {"moves_compared": int(n_cmp), "molecule_and_accept_identical": same_chain,
                                                  "max_ddE_over_max1_dE": float(ddE.max()) if ddE.size else 0.0}
        single["unit"] = "moves/s"
        single["cluster_ctas"] = a.single_cluster
        single["note"] = ("ONE Markov chain per GPU (what a single Plum run sees), a cluster of CTAs sharing the chain; pivot_mode 0 builds "
                          "pivot arms in the reference's operation order (trial coordinates bit-identical to the reference's, one "
                          "dependent step per bead), pivot_mode 1 as a prefix sum (coordinates equal to ~1e-13), pivot_mode 2 evaluates the energy "
                          "change from the prefix-sum arms while two warps build the exact ones, which are what gets committed "
                          "(bit-identical coordinates, dE equal to ~1e-13 relative); the throughput legs run config.pivot_mode")

    # ---------------- roofline of the dominant kernel (k_chain): executed work from the tracked ncu capture
    peaks = {}
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    fp64_peak_gflops, l2_peak_gbs = main["peaks"]
    steps_per_s = R * K * M / main["dev_s"]             # this rank's chains
    prof_file, prof = load_profile_summary()
    n_in, n_eval, n_acc = main["counters"]
    kernel_us = main["dev_s"] * 1e6 / max(main["launches"], 1)
    roofline = {"bound": "fp64", "kernel": "k_chain", "unit": "TFLOP/s", "peak": fp64_peak_gflops / 1e3,
                "peak_source": "FP64 FMA microbenchmark measured in this run (pg_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                "launches": int(main["launches"]), "avg_launch_us": kernel_us,
                "avg_launch_note": f"one launch = {M} steps of each of the {R} chains of this GPU (CUDA events around the launch)"}
    if prof:
        fl_step = prof["executed_fp64_flop_per_step"]
        wi_step = prof["warp_instructions_per_step"]
        l2_step = prof.get("l2_bytes_per_step")
        ach = fl_step * steps_per_s / 1e12
        clk = (clock_info.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        roofline.update({
            "achieved": ach, "frac": ach / (fp64_peak_gflops / 1e3),
            "achieved_note": "EXECUTED FP64 flop per step (2 x dfma + dadd + dmul thread instructions of the tracked ncu capture) x the "
                             "steps/s measured here with CUDA events",
            "executed_fp64_flop_per_step": fl_step, "profile": "profiles/" + prof_file, "profile_section": prof.get("section"),
            "device_wide_pct_under_ncu": prof.get("device_wide_pct"),
            "traffic": prof.get("dram_bytes_per_step"),
            "traffic_note": "dram__bytes_read + dram__bytes_write per step from the same capture (one chain, cold caches): the working "
                            "set is L2-resident by design",
            "issue_slots": {"warp_instructions_per_step": wi_step, "achieved_per_s": wi_step * steps_per_s,
                            "peak_per_s": n_sm * 4 * clk, "frac": wi_step * steps_per_s / (n_sm * 4 * clk),
                            "note": "warp instructions/s over SMs x 4 schedulers x the SM clock sampled during the timed region"},
            "l2_view": None if not l2_step else {"achieved": l2_step * steps_per_s / 1e9, "peak": l2_peak_gbs, "unit": "GB/s",
                                                 "frac": l2_step * steps_per_s / 1e9 / l2_peak_gbs,
                                                 "peak_source": "L2 streaming-read microbenchmark measured in this run (pg_measure_l2_peak)"},
        })
    else:
        roofline.update({"achieved": None, "frac": None, "traffic": None,
                         "note": "profiles/r*_ncu_summary.json with a k_chain section not found: executed counters unavailable"})
    # what had to be computed: pair configurations inside a cutoff (counted by the kernel) + the reciprocal-space update
    nec_flops = n_in * F_PAIR + main["alg_flops"] - 2 * main["evals"] * F_PAIR
    roofline["necessary_work_view"] = {
        "pair_configurations_in_range_per_step": n_in / max(n_eval, 1), "flops_per_step": nec_flops / max(n_eval, 1),
        "achieved": nec_flops / main["dev_s"] / 1e12, "frac": nec_flops / main["dev_s"] / 1e12 / (fp64_peak_gflops / 1e3),
        "note": "SURVEY §8(d) prices (36 flop per pair configuration, 2 n_q K 16 + 12 K for the reciprocal update) applied only to the pair "
                "configurations that lie inside a cutoff, as counted by the kernel itself"}
    roofline["algorithmic_view"] = {
        "flops_per_step": main["alg_flops"] / (R * K * M), "achieved": main["alg_flops"] / main["dev_s"] / 1e12,
        "speedup_equivalent": main["alg_flops"] / main["dev_s"] / 1e12 / (fp64_peak_gflops / 1e3),
        "note": "SURVEY §8(d) formula: EVERY moved-bead x partner pair priced at 36 flop although all but ~0.1 % lie outside both cutoffs "
                "and are never evaluated (cell grid, charged list) — not a utilisation figure, only how many reference-style pair "
                "evaluations per second the path stands for"}
    roofline["hbm_view"] = {"peak": float(peaks.get("hbm_gbs", 6650.0)),
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                            "note": "not the bound: DRAM is touched on first use only"}

    # ---------------- examples through the unchanged driver
    examples = None
    if rank == 0 and world == 1 and not a.no_examples:   # (multi-GPU runs: the other ranks would sit in a collective meanwhile)
        replay = _replay()
        examples = examples_block(replay.PLUM_GPU, None, parallel=False, steps=EXAMPLE_STEPS_GPU)
        for name in EXAMPLES:     # the same examples with their own sampling / output frequencies
            try:
                if isinstance(examples.get(name), dict):
                    examples[name]["own_frequencies"] = run_example_own_frequencies(name, EXAMPLE_STEPS_GPU)
            except Exception as e:   # noqa: BLE001
                examples[name]["own_frequencies"] = {"error": repr(e)}

    # ---------------- k-sharded full S(k) recompute
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
            if max_over_ranks(1.0 if setup_err else 0.0) > 0.0:
                recompute[name] = {"error": setup_err or "set-up failed on another rank"}
                if eng is not None:
                    eng.close()
                continue
            try:
                sharded.attach_peers(eng, rank, world)
                recompute[name] = dict(n_charged=int(np.count_nonzero(s2.q)),
                                       **sharded.time_fused_recompute(eng, rank, world, t_init["recip"]))
                recompute[name]["launches"] = int(eng.launch_count() - l0)
                if world > 1:
                    eng.sk_detach()
                    recompute[name]["nccl_allgather_baseline"] = sharded.time_sharded_recompute(eng, rank, world, t_init["recip"], reps=10)
            except Exception as e:   # noqa: BLE001
                recompute[name] = {"error": repr(e)}
            finally:
                if eng is not None:
                    eng.close()

    # ---------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        t_ion, t_chain = cpu_sample((7, 6, 2))
        v = mix_rate(t_ion, t_chain, main["p_ion"])
        cpu = {"value": v, "unit": "moves/s", "cores": 1, "kind": "port",
               "sample": f"6 ion moves ({np.mean(t_ion):.3f} s each) + 2 chain moves of 100 flagged beads ({np.mean(t_chain):.3f} s each) on "
                         f"the same 22000-bead system, reference pairwise algorithm (oracle port, map-free, new-configuration energies only "
                         f"like the reference), mixed with the realised ion fraction {main['p_ion']:.3f}; plum_ref cannot hold N=22000 "
                         f"(SURVEY.md §0.8)"}

    # bytes crossing PCIe per step in the e2e leg (per GPU): generator state + launch arguments down; step log, generator
    # state (padded to 640 words), status words and the accepted coordinates (3 doubles per bead, gathered on the device)
    # up — per batch and replica, in one copy
    nb = main["batches_per_step"]
    h2d = float(R * nb * (625 * 4 + 768))
    d2h = float(R * (16 * M + nb * (640 * 4 + 16 + ((24 * sysm.n + 15) // 16) * 16)))
    evals_total = sum_over_ranks(main["evals"])     # a collective: every rank takes part
    if rank == 0:
        line = {
            "metric": "MC moves/sec", "value": value, "unit": "moves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["dev_s"] * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "moves_per_step": M, "replicas_per_gpu": R, "cluster_ctas_per_chain": main["cluster"],
                       "pivot_mode": main["pivot_mode"],
                       "mode": f"throughput: {R} independent Markov chains per GPU, all in one k_chain launch per step; the rate of ONE chain is in single_chain",
                       "l2": "flushed between steps (256 MiB fill); within a step each chain's working set stays L2-resident by design (north_star)"},
            "pair_dE_evals_per_s": evals_total / main["dev_s"],
            "e2e": {"value": e2e_value, "unit": "moves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": main["t_e2e"] * 1e3 / K,
                    "caller": f"native C++ caller (plum_b200/host/mc_bench.cc pb_run_chain), ONE host thread per GPU driving {R} chains: per batch of "
                              f"{a.batch} steps pg_chain_set_rng (std::mt19937 state down), pg_chain_run_multi (one launch), pg_chain_steps (log up), "
                              f"pg_chain_get_rng, pg_download_positions (accepted coordinates up, as ForceField::TranslationalBatch does); = pg_chain_run_multi_io: "
                              f"a pack kernel gathers logs / generators / coordinates of all chains, one copy brings them up, up to 8 worker "
                              f"threads inside the call hand them to the caller's arrays"},
            "gpu_launches": int(main["launches"]), "roofline": roofline, "cpu_baseline": cpu, "clocks": clock_info,
            "resident_matches_e2e": main["same"], "accept_ratio": main["accept"],
            "wall_ms_per_step_resident": main["wall_res"] * 1e3 / K,
            "single_chain": single, "examples": examples, "sharded_recompute": recompute,
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
    ap.add_argument("--no-single", action="store_true", help="skip the single-chain measurements")
    ap.add_argument("--no-recompute", action="store_true", help="skip the k-sharded full S(k) recompute timing")
    ap.add_argument("--no-examples", action="store_true", help="skip the four reference examples")
    ap.add_argument("--no-plum-ref", action="store_true", help="reference arm: skip the real plum_ref binary's legs")
    ap.add_argument("--replicas-per-gpu", type=int, default=0, help="independent Markov chains per GPU; 0 = two per SM")
    ap.add_argument("--cluster", type=int, default=1, help="CTAs per chain in throughput mode")
    ap.add_argument("--single-cluster", type=int, default=16, help="CTAs per chain in the single-chain measurement")
    ap.add_argument("--pivot-mode", type=int, default=0, help="throughput legs: 0 pivot arms in the reference's operation order (fastest bit-identical mode for a fleet: the other CTA of the SM hides the arm); 1 prefix sums; 2 exact arms committed, energies from prefix-sum arms (the single-chain default)")
    ap.add_argument("--batch", type=int, default=MOVES_PER_STEP, help="steps per host round trip in the e2e leg")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
