"""Synthetic large polyelectrolyte system — BASELINE.json configs[4], SURVEY.md §8(d) "S".

  200 chains x 100 beads, symbol P, sigma 2.22724679535, epsilon 0.1, WCA (cutoff -1), rigid
  bond 2.5; every 10th monomer q = -1 (2000 charged monomers) + 2000 P(+1) counter-ions
  => N = 22 000 beads (4000 charged); cubic L = 200; npbc = 3; lB = 2.5; alpha = 0.004
  => real_cutoff 52 (< L/2: central image only), repl_cell 10, 3574 k vectors (1787 half space);
  move mix 0.5 bead-translate / 0.1 COM / 0.3 pivot / 0.0 crankshaft / 0.1 reptation.
  Chains are self-avoiding random walks (step 2.5, overlaps < 2.0 rejected) from uniform random
  starts, ions uniform with the same overlap test; generator seed 20261017.

`write_inputs` emits run.in / input_crd.dat / input_top.dat in the reference's own formats
(src/simulation/simulation.cc:582-646) so the same system can be fed to Plum's driver.
"""
from __future__ import annotations

import os

import numpy as np

from . import runin

SEED = 20261017
SIGMA = 2.22724679535


def make_run_in(n_steps: int = 1000, alpha: float = 0.004, spring: bool = False,
                move_prob=(0.5, 0.1, 0.3, 0.0, 0.1)) -> str:
    lines = [
        "s1_input_coordinate_file        input_crd.dat",
        "s1_input_topology_file          input_top.dat",
        "s1_prefix_of_output_file_names  output",
        f"s1_total_simulation_steps       {n_steps}",
        "s1_equilibrium_steps            1000000000",
        "s1_sampling_frequency           1000000000",
        "s1_sampling_print_frequency     1000000000",
        "s1_trajectory_print_frequency   1000000000",
        "s1_PBC_dimensions               3",
        "s1_beta_(1/kBT)                 1",
        "s1_MC_move_size_in_unit_length  2",
        "s1_calc_chem_pot                0",
        "s1_number_of_surface_sites      0",
        "s1_number_of_surf_counterions   0",
        "s1_number_of_grafted_chains     0",
        "s1_number_of_grafted_counterion 0",
        f"s1_prob_b_bead_translation      {move_prob[0]}",
        f"s1_prob_p_COM_translation       {move_prob[1]}",
        f"s1_prob_p_pivot                 {move_prob[2]}",
        f"s1_prob_p_crankshaft            {move_prob[3]}",
        f"s1_prob_p_random_reptation      {move_prob[4]}",
        "s1_vp_bin_resolution_in_ul      10",
        "s1_bead_size_virial_pressure    2.5",
        "s2_use_short_range_potential    1",
        "s2_use_Ewald_potential          1",
        f"s2_use_bond_potential           {1 if spring else 0}",
        "s2_use_rigid_bond               1",
        "s2_use_angle_potential          0",
        "s2_use_dihedral_potential       0",
        "s2_use_external_potential       0",
        "s2_use_grand_canonical_MC_move  0",
        "s3_short_range_potential_type   TruncatedLJ",
        "s3_LJ_cutoff                    -1",
        "s3_bead_type_1                  P",
        f"s3_sigma_for_bead_type_1        {SIGMA}",
        "s3_epsilon_for_bead_type_1      0.1",
        "s3_end_flag                     end",
        "s3_Ewald_potential_type         Coul",
        "s3_Bjerrum_length               2.5",
        "s3_dielectric_constant          80",
        f"s3_Ewald_alpha                  {alpha}",
        "s3_use_dipole_correction        0",
    ]
    if spring:
        lines += ["s3_bond_potential_type         Spring", "s3_spring_constant              30",
                  "s3_equilibrium_bond_length      2.5"]
    lines += ["s3_rigid_bond_length           2.5"]
    return "\n".join(lines) + "\n"


class _Grid:
    """Cell grid for the overlap test while growing the configuration."""

    def __init__(self, L, cell):
        self.L = L
        self.n = max(1, int(L // cell))
        self.w = L / self.n
        self.cells = {}

    def _key(self, p):
        return tuple(int(np.floor((c % self.L) / self.w)) % self.n for c in p)

    def ok(self, p, rmin2):
        k = self._key(p)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    kk = ((k[0] + dx) % self.n, (k[1] + dy) % self.n, (k[2] + dz) % self.n)
                    for o in self.cells.get(kk, ()):
                        d = p - o
                        d -= self.L * np.round(d / self.L)
                        if d @ d < rmin2:
                            return False
        return True

    def add(self, p):
        self.cells.setdefault(self._key(p), []).append(p.copy())


def make_system(n_chains: int = 200, chain_len: int = 100, charged_every: int = 10, L: float = 200.0,
                seed: int = SEED, bond: float = 2.5, rmin: float = 2.0) -> runin.System:
    rng = np.random.default_rng(seed)
    grid = _Grid(L, max(rmin, 2.5))
    xyz, q, first = [], [], [0]
    for _ in range(n_chains):
        while True:
            chain = []
            p = rng.uniform(0, L, 3)
            if not grid.ok(p, rmin * rmin):
                continue
            chain.append(p)
            stuck = False
            while len(chain) < chain_len:
                for _attempt in range(200):
                    v = rng.normal(size=3)
                    c = chain[-1] + bond * v / np.linalg.norm(v)
                    if grid.ok(c, rmin * rmin) and all(((c - o) @ (c - o)) >= rmin * rmin for o in chain[:-1]):
                        chain.append(c)
                        break
                else:
                    stuck = True
                    break
            if not stuck:
                break
        for b, p in enumerate(chain):
            grid.add(p)
            xyz.append(p)
            q.append(-1.0 if (b % charged_every == 0) else 0.0)
        first.append(len(q))
    n_ions = int(round(-sum(q)))
    for _ in range(n_ions):
        while True:
            p = rng.uniform(0, L, 3)
            if grid.ok(p, rmin * rmin):
                break
        grid.add(p)
        xyz.append(p)
        q.append(1.0)
        first.append(len(q))
    n = len(q)
    return runin.System(np.array(xyz, dtype=np.float64).reshape(n, 3), np.array(q, dtype=np.float64), ["P"] * n,
                        np.array(first, dtype=np.int32), [L, L, L])


def write_inputs(directory: str, sysm: runin.System, n_steps: int = 1000, alpha: float = 0.004, spring: bool = False,
                 move_prob=(0.5, 0.1, 0.3, 0.0, 0.1)):
    os.makedirs(directory, exist_ok=True)
    with open(os.path.join(directory, "run.in"), "w") as f:
        f.write(make_run_in(n_steps, alpha, spring, move_prob))
    with open(os.path.join(directory, "input_top.dat"), "w") as f:
        f.write(f"TotNoOfBeads: {sysm.n}\nTotNoOfMolec: {sysm.n_mol}\n")
        f.write(f"Box_Length_X: {sysm.box[0]!r}\nBox_Length_Y: {sysm.box[1]!r}\nBox_Length_Z: {sysm.box[2]!r}\n")
        f.write("Bonds: -1\nAngles: -1\nDihedrals: -1\n")
    with open(os.path.join(directory, "input_crd.dat"), "w") as f:
        f.write(f"{sysm.n}\n \n")
        for m in range(sysm.n_mol):
            for i in range(sysm.mol_first[m], sysm.mol_first[m + 1]):
                x, y, z = sysm.xyz[i]
                f.write(f"{m} {sysm.symbol[i]} {x:.17g} {y:.17g} {z:.17g} {sysm.q[i]:g}\n")


def load(n_chains: int = 200, chain_len: int = 100, alpha: float = 0.004, cache_dir: str = None,
         charged_every: int = 10, L: float = 200.0):
    """(RunIn, System, TypeTable, params) of the synthetic system; the configuration is cached as .npz.
    Defaults: S of SURVEY.md §8(d).  `load_full` gives the stress variant S-full."""
    r = runin.parse_run_in(make_run_in(alpha=alpha))
    key = f"synth_S_{n_chains}x{chain_len}_seed{SEED}.npz" if (charged_every == 10 and L == 200.0) else \
        f"synth_S_{n_chains}x{chain_len}_q{charged_every}_L{L:g}_seed{SEED}.npz"
    path = os.path.join(cache_dir, key) if cache_dir else None
    if path and os.path.exists(path):
        z = np.load(path)
        sysm = runin.System(z["xyz"], z["q"], ["P"] * int(z["q"].shape[0]), z["mol_first"], [L] * 3)
    else:
        sysm = make_system(n_chains, chain_len, charged_every=charged_every, L=L)
        if path:
            os.makedirs(cache_dir, exist_ok=True)
            tmp = f"{path}.{os.getpid()}.tmp.npz"   # atomic: several ranks may generate at once
            np.savez_compressed(tmp, xyz=sysm.xyz, q=sysm.q, mol_first=sysm.mol_first)
            os.replace(tmp, path)
    types = runin.TypeTable(r, sysm.symbol)
    return r, sysm, types, runin.params_dict(r, sysm.box, types)


def load_full(cache_dir: str = None):
    """Stress variant S-full: every monomer charged => N = 40 000 (all charged), L = 250,
    alpha = 0.003 => real_cutoff 60, repl_cell 10, K = 4138."""
    return load(200, 100, alpha=0.003, cache_dir=cache_dir, charged_every=1, L=250.0)
