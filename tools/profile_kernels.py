#!/usr/bin/env python3
"""Exercises every kernel of the path once or twice so that one ncu pass can capture them all:
   S (22 000 beads): init (k_tot_pairs, k_sk_slice, k_sk_reduce, k_tot_final), k_move<true> ion + chain,
                     k_trials (30 trials), k_delta insert / delete (8-bead chain + 8 ions), k_commit, k_propose (two 24-step batches)
                     k_chain (40 steps, cluster 1 and 8), k_sk_block / k_sk_finish / k_sk_energy (full S(k) recompute), k_vol_* (pressure sample)
   bulk_nvt / confined_nvt (reference examples): init (k_tot_pairs on a few hundred beads), k_move<false> ion + chain (multi-image path),
                     k_wall_force, k_vol_* on bulk_nvt.
Run:  ncu --set full --clock-control none --import-source on -k regex:'k_' -o gpurun_out/prof_<tag>_all python tools/profile_kernels.py   (tools/round_evidence.sh does it and keeps the raw-metric CSV)
"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import replay
from plum_b200 import synth
from plum_b200.engine import Engine

rng = np.random.default_rng(0)
r, s, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
eng = Engine(params, device=0, capacity_beads=s.n + 64)
ids = types.ids(s.symbol)
eng.upload(s.xyz, s.q, ids, s.mol_first)
for _ in range(2):
    eng.init_energy()
ions = [m for m in range(s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] == 1]
chains = [m for m in range(s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] > 1]
for rep in range(2):
    for pool in (ions, chains):
        m = int(rng.choice(pool)); f, l = int(s.mol_first[m]), int(s.mol_first[m + 1])
        eng.delta_e(m, s.xyz[f:l] + rng.normal(scale=0.3, size=(l - f, 3)), np.ones(l - f, dtype=np.uint8))
        eng.commit(False)
# k_propose (device-side proposals): two batches of 24 steps through pg_mc_* (bead, COM, pivot, reptation kinds)
from plum_b200 import mcgen
g = mcgen.Generator.for_run(r, s.mol_first, 4)
for rep in range(2):
    descs, rows, n_rows = [], [], 0
    while len(descs) < 24:
        kind, d, rv = g.next()
        if kind < 0:
            continue
        if kind == 2:
            d.rv_offset = n_rows
        descs.append(d); rows.append(rv); n_rows += rv.shape[0]
    eng.mc_upload(descs, np.concatenate(rows) if n_rows else None)
    eng.mc_run(0, len(descs))
tP = types.ids(["P"])[0]
cl = 4
cx = np.zeros((2 * cl, 3)); cq = np.zeros(2 * cl); ct = np.full(2 * cl, tP, dtype=np.int32)
for i in range(cl):
    cx[i] = [100 + 2.5 * i, 100, 100]; cq[i] = -1.0
    cx[cl + i] = cx[i] + 1.9; cq[cl + i] = 1.0
d = rng.normal(size=(30, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
b1 = cx[cl - 1] + 2.5 * d
for rep in range(2):
    eng.trial_energies(b1, b1 + 1.7, tP, -1.0, tP, 1.0, 1, cx, cq, ct, cl)
new_xyz = np.concatenate([cx[:cl], b1[:4], cx[cl:], b1[:4] + 1.7])
for rep in range(2):
    eng.insert_molecules([8] + [1] * 8, new_xyz, np.array([-1.0] * 8 + [1.0] * 8), np.full(16, tP, dtype=np.int32))
    eng.delete_molecules(s.n_mol, s.n_mol + 8)
eng.totals()
# round 2: the device-resident chain (k_chain, 40 steps, one CTA and a cluster of 8), the full S(k) recompute
# (k_sk_block / k_sk_finish / k_sk_energy) and one volume-perturbation pressure sample (k_vol_*)
bl, vary = mcgen.bond_settings(r)
for cluster in (1, 8):
    eng.chain_configure(r.phantom, r.move_size, r.move_prob, bl, vary_bond=vary, cluster=cluster)
    eng.chain_seed(3)
    for rep in range(2):
        eng.chain_run(40)
for rep in range(2):
    eng.recompute_sk()
    eng.vol_scaling_sample(0)
eng.close()

for name in ("bulk_nvt", "confined_nvt"):
    r, s, types, params = replay.load_golden(name)
    eng = Engine(params, device=0, capacity_beads=s.n + 64)
    eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first); eng.init_energy()
    ions = [m for m in range(r.phantom, s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] == 1]
    chains = [m for m in range(r.phantom, s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] > 1]
    for rep in range(2):
        for pool in (ions, chains):
            m = int(rng.choice(pool)); f, l = int(s.mol_first[m]), int(s.mol_first[m + 1])
            eng.delta_e(m, s.xyz[f:l] + rng.normal(scale=0.3, size=(l - f, 3)), np.ones(l - f, dtype=np.uint8))
            eng.commit(False)
    if r.phantom:
        for rep in range(2):
            eng.wall_force(r.phantom)
    else:
        for rep in range(2):
            eng.vol_scaling_sample(0)
    eng.close()
print("profile_kernels done")
