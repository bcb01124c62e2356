#!/usr/bin/env python3
"""Generates the committed golden fixtures from the REAL reference.

Runs only where /root/reference and oracle/_ref/plum_ref exist (the build
container).  For each example of BASELINE.json's configs it
  1. copies the three input files (run.in, input_crd.dat, input_top.dat) next to
     this script — they are the reference's example *inputs*, needed on the GPU
     box where /root/reference does not exist;
  2. runs plum_ref (unmodified reference arithmetic + seed/trace hooks, see
     oracle/build_ref.py) with PLUM_SEED and writes
       short/<example>_seed<k>.trace.gz   full replayable trace (coordinates of every
                                          trial, every CBMC BeadsEnergy call) of the
                                          first SHORT_STEPS steps;
       long/<example>_seed<k>.npz         (k = 1; 2 and 3 with --long-seed k) the first 10^5 steps: accept/reject bits and
                                          move kinds for every step, dE for the first
                                          2000 steps and every 25th step after, the
                                          four running totals every 1000 steps.
Usage: python tests/golden/make_golden.py [--long-from DIR]
  --long-from DIR : reuse existing <DIR>/<example>/trace_seed1.txt instead of re-running
"""
import argparse
import gzip
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import replay  # noqa: E402

EXAMPLES = ["bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"]
SHORT_STEPS = {"bulk_nvt": 250, "confined_nvt": 500, "bulk_muvt": 400, "confined_muvt": 400}
SHORT_SEEDS = [1, 2]
LONG_STEPS = 100000
REF_EXAMPLES = "/root/reference/examples"


def compact_long(lines, path, n_steps):
    """Arrays are indexed by step-1; steps that attempted no move (empty system) keep kind 3."""
    kind = np.full(n_steps, 3, dtype=np.uint8)      # 0 translational, 1 GC insert, 2 GC delete, 3 none
    accept = np.zeros(n_steps, dtype=np.uint8)
    mtype = np.full(n_steps, -1, dtype=np.int8)
    mol = np.full(n_steps, -1, dtype=np.int32)
    dE_idx, dE_val = [], []
    tot_idx, tot_val = [], []
    init = None
    for ln in lines:
        f = ln.split()
        if not f:
            continue
        if f[0] == "I":
            init = [replay.hx(x) for x in f[2:6]]
            continue
        if f[0] not in ("T", "G"):
            continue
        step = int(f[1])
        i = step - 1
        if f[0] == "T":
            kind[i] = 0
            mtype[i] = int(f[2])
            mol[i] = int(f[3])
            accept[i] = int(f[5])
            val = replay.hx(f[4])                      # dE
            tots = [replay.hx(x) for x in f[6:10]]
        else:
            kind[i] = 1 if f[2] == "I" else 2
            mol[i] = int(f[3])                         # insertion 0/1, deletion: molecule id or -1
            accept[i] = 1 if ((f[2] == "I" and int(f[3]) == 1) or (f[2] == "D" and int(f[3]) >= 0)) else 0
            val = replay.hx(f[4])                      # Rosenbluth weight at the acceptance test (-1: not reached)
            tots = [replay.hx(x) for x in f[6:10]]
        if step <= 2000 or step % 25 == 0:
            dE_idx.append(step)
            dE_val.append(val)
        if step % 500 == 0:
            tot_idx.append(step)
            tot_val.append(tots)
    np.savez_compressed(
        path, init=np.array(init), kind=kind, accept=np.packbits(accept), n_steps=n_steps, move_type=mtype, mol=mol,
        dE_step=np.array(dE_idx, dtype=np.int32), dE=np.array(dE_val), tot_step=np.array(tot_idx, dtype=np.int32),
        tot=np.array(tot_val))


def extras():
    """Fixtures of the two samplers that sit next to the per-move path:
       short/bulk_nvt_mu_seed1.{trace.gz,stat.dat}            230 steps with s1_calc_chem_pot 1 (Widom sampler
                                                              on the trajectory, mu column of output_stat.dat);
       short/confined_nvt_pressure_seed1.{trace.gz,stat.dat}  3000 steps with coordinates: the wall-force
                                                              pressure columns of output_stat.dat at steps
                                                              1000/2000/3000 (CalcPressureForceLJELSlit)."""
    jobs = [("bulk_nvt", "bulk_nvt_mu_seed1", 230, False, {"s1_calc_chem_pot": 1}),
            ("confined_nvt", "confined_nvt_pressure_seed1", 3000, True, {})]
    for ex, name, steps, xyz, over in jobs:
        lines, files = replay.run_plum_ref(os.path.join(REF_EXAMPLES, ex), steps, 1, xyz=xyz, overrides=over,
                                           want_files=("output_stat.dat",))
        with gzip.open(os.path.join(HERE, "short", name + ".trace.gz"), "wt") as f:
            f.write("\n".join(lines))
        with open(os.path.join(HERE, "short", name + ".stat.dat"), "w") as f:
            f.write(files["output_stat.dat"])
        print("extra", name, len(lines), "lines")


def synth_crank():
    """examples/synth_crank + short/synth_crank_seed{1,2}.trace.gz: the small cut of synth_spring with rigid bonds
    and crankshaft moves on (s1_prob_p_crankshaft 0.3 — zero in all four of the reference's examples), 400 steps
    with trial coordinates.  The rotation is Eigen's AngleAxisd in the reference; plum_ref is built against the
    stand-in of plum_b200/host/eigen_standin, which follows Eigen's operation order."""
    from plum_b200 import synth
    sysm = synth.make_system(n_chains=4, chain_len=24, charged_every=6)
    dst = os.path.join(HERE, "examples", "synth_crank")
    synth.write_inputs(dst, sysm, n_steps=400, alpha=0.004, move_prob=(0.3, 0.1, 0.2, 0.3, 0.1))
    for seed in SHORT_SEEDS:
        lines = replay.run_plum_ref(dst, 400, seed, xyz=True)
        with gzip.open(os.path.join(HERE, "short", f"synth_crank_seed{seed}.trace.gz"), "wt") as f:
            f.write("\n".join(lines))
        print("short synth_crank", seed, len(lines), "lines")


def vol_pressure():
    """short/{bulk_nvt,synth_spring}_volp_seed1.trace.gz: 330 steps with coordinates and sampling every 10 steps, so
    that ForceField::CalcPressureVolScalingHSELSlit (pressure.cc:187-387) runs at steps 200 and 300; the V lines
    carry its accumulators (oracle/build_ref.py pressure hook).  bulk_nvt: rigid bonds; synth_spring: the Spring
    bond potential (per-bead scaling branch, bond term)."""
    over = {"s1_sampling_frequency": 10, "s1_equilibrium_steps": 100}
    for ex_dir, name in ((os.path.join(HERE, "examples", "bulk_nvt"), "bulk_nvt_volp_seed1"),
                         (os.path.join(HERE, "examples", "synth_spring"), "synth_spring_volp_seed1")):
        lines = replay.run_plum_ref(ex_dir, 330, 1, xyz=True, overrides=over)
        assert sum(1 for ln in lines if ln.startswith("V ")) == 2
        with gzip.open(os.path.join(HERE, "short", name + ".trace.gz"), "wt") as f:
            f.write("\n".join(lines))
        print("vol_pressure", name, len(lines), "lines")


def synth_spring():
    """examples/synth_spring + short/synth_spring_seed{1,2}.trace.gz: a small cut of the synthetic polyelectrolyte
    (plum_b200/synth.py: 4 chains x 24 beads + counter-ions, same box / alpha / move mix as the benchmark system S)
    with the Spring bond potential on, so that Pivot and RandomReptation vary the bond length (the extra draw of
    src/molecules/molecule.cc:184-187, :278-281) — the reference ships no spring example."""
    from plum_b200 import synth
    sysm = synth.make_system(n_chains=4, chain_len=24, charged_every=6)
    dst = os.path.join(HERE, "examples", "synth_spring")
    synth.write_inputs(dst, sysm, n_steps=400, alpha=0.004, spring=True)
    for seed in SHORT_SEEDS:
        lines = replay.run_plum_ref(dst, 400, seed, xyz=True)
        with gzip.open(os.path.join(HERE, "short", f"synth_spring_seed{seed}.trace.gz"), "wt") as f:
            f.write("\n".join(lines))
        print("short synth_spring", seed, len(lines), "lines")


def compact_cut(lines, path, n_trials_kept):
    """One npz per seed: every step's move kind / molecule / dE / accept bit / four running totals, and the trial
    coordinates the reference built (X lines) for the first n_trials_kept steps."""
    kind, mol, dE, acc, tot, t_off, t_xyz = [], [], [], [], [], [0], []
    init = None
    for i, ln in enumerate(lines):
        f = ln.split()
        if not f:
            continue
        if f[0] == "I":
            init = [replay.hx(x) for x in f[2:6]]
        if f[0] != "T":
            continue
        kind.append(int(f[2])); mol.append(int(f[3])); dE.append(replay.hx(f[4])); acc.append(int(f[5]))
        tot.append([replay.hx(x) for x in f[6:10]])
        if len(kind) <= n_trials_kept:
            x = lines[i + 1].split()
            assert x[0] == "X"
            n = int(x[1])
            t_xyz.append(np.array([[replay.hx(x[3 + 4 * k + a]) for a in range(3)] for k in range(n)]))
            t_off.append(t_off[-1] + n)
    np.savez_compressed(path, init=np.array(init), kind=np.array(kind, dtype=np.int8), mol=np.array(mol, dtype=np.int32),
                        dE=np.array(dE), accept=np.array(acc, dtype=np.uint8), tot=np.array(tot),
                        trial_off=np.array(t_off, dtype=np.int32), trial_xyz=np.concatenate(t_xyz))


def synth_cut(steps=2000, seeds=(1, 2), n_trials_kept=300):
    """examples/synth_cut + long/synth_cut_seed{1,2}.npz: the 1320-bead cut of the benchmark system S (plum_b200/synth.py:
    12 chains x 100 beads, every 10th bead charged, + 120 counter-ions; SAME box L = 200, alpha = 0.004 — hence the same
    cutoffs and K = 3574 — and the same move mix), the largest cut of S the reference walks in minutes.  It spans several
    partner tiles / cells / charged-list chunks of the single-image kernels (k_move<true>, k_chain), which the 112-bead
    synth_spring fixture does not."""
    from plum_b200 import synth
    sysm = synth.make_system(n_chains=12, chain_len=100, charged_every=10)
    dst = os.path.join(HERE, "examples", "synth_cut")
    synth.write_inputs(dst, sysm, n_steps=steps, alpha=0.004, spring=False)
    for seed in seeds:
        lines = replay.run_plum_ref(dst, steps, seed, xyz=True)
        compact_cut(lines, os.path.join(HERE, "long", f"synth_cut_seed{seed}.npz"), n_trials_kept)
        print("synth_cut seed", seed, len(lines), "lines", flush=True)


def long_seed(seed, examples=None):
    """long/<example>_seed<seed>.{npz,stat.dat} for one more seed (the inputs are the committed copies)."""
    for ex in (examples or EXAMPLES):
        lines, files = replay.run_plum_ref(os.path.join(HERE, "examples", ex), LONG_STEPS, seed, xyz=False,
                                           want_files=("output_stat.dat",))
        compact_long(lines, os.path.join(HERE, "long", f"{ex}_seed{seed}.npz"), LONG_STEPS)
        with open(os.path.join(HERE, "long", f"{ex}_seed{seed}.stat.dat"), "w") as f:
            f.write(files["output_stat.dat"])
        print("long", ex, "seed", seed, len(lines), "lines", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--long-seed", type=int, default=0, help="only long/<example>_seed<k>.{npz,stat.dat} for this extra seed")
    ap.add_argument("--examples", default="", help="comma-separated subset for --long-seed")
    ap.add_argument("--long-from", default=None)
    ap.add_argument("--skip-long", action="store_true")
    ap.add_argument("--extras-only", action="store_true", help="only the sampler fixtures (extras())")
    ap.add_argument("--crank-only", action="store_true", help="only the crankshaft fixture (synth_crank())")
    ap.add_argument("--volp-only", action="store_true", help="only the volume-perturbation pressure fixtures (vol_pressure())")
    ap.add_argument("--synth-only", action="store_true", help="only the spring-bond fixture (synth_spring())")
    ap.add_argument("--cut-seed", type=int, default=0, help="only the 1320-bead cut of S (synth_cut()) for this seed")
    ap.add_argument("--cut-steps", type=int, default=2000)
    a = ap.parse_args()
    if not replay.have_plum_ref():
        raise SystemExit("oracle/_ref/plum_ref missing: run python oracle/build_ref.py")
    if a.long_seed:
        long_seed(a.long_seed, [e for e in a.examples.split(",") if e] or None)
        return
    if a.extras_only:
        os.makedirs(os.path.join(HERE, "short"), exist_ok=True)
        extras()
        return
    if a.synth_only:
        synth_spring()
        return
    if a.crank_only:
        synth_crank()
        return
    if a.volp_only:
        vol_pressure()
        return
    if a.cut_seed:
        synth_cut(a.cut_steps, (a.cut_seed,))
        return
    os.makedirs(os.path.join(HERE, "short"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "long"), exist_ok=True)
    for ex in EXAMPLES:
        src = os.path.join(REF_EXAMPLES, ex)
        dst = os.path.join(HERE, "examples", ex)
        os.makedirs(dst, exist_ok=True)
        for fn in ("run.in", "input_crd.dat", "input_top.dat"):
            shutil.copyfile(os.path.join(src, fn), os.path.join(dst, fn))
        for seed in SHORT_SEEDS:
            lines = replay.run_plum_ref(src, SHORT_STEPS[ex], seed, xyz=True)
            with gzip.open(os.path.join(HERE, "short", f"{ex}_seed{seed}.trace.gz"), "wt") as f:
                f.write("\n".join(lines))
            print("short", ex, seed, len(lines), "lines")
        if a.skip_long:
            continue
        if a.long_from:
            with open(os.path.join(a.long_from, ex, "trace_seed1.txt")) as f:
                lines = f.read().split("\n")
        else:
            lines = replay.run_plum_ref(src, LONG_STEPS, 1, xyz=False)
        compact_long(lines, os.path.join(HERE, "long", f"{ex}_seed1.npz"), LONG_STEPS)
        print("long", ex, len(lines), "lines")
    extras()
    synth_spring()
    synth_crank()
    vol_pressure()
    synth_cut()


if __name__ == "__main__":
    main()
