"""Replay a trace written by the real reference (oracle/_ref/plum_ref, format in
oracle/build_ref.py) through an engine-like object — the CPU oracle or the CUDA
engine — and report the worst disagreement.  Test helper, not product code."""
from __future__ import annotations

import dataclasses
import os
import subprocess
import tempfile
from typing import List, Optional

import numpy as np

from plum_b200 import runin

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUM_REF = os.path.join(REPO, "oracle", "_ref", "plum_ref")
GOLDEN = os.path.join(REPO, "tests", "golden")
VLE = 1.0e8


def hx(s: str) -> float:
    return float.fromhex(s)


def have_plum_ref() -> bool:
    return os.path.exists(PLUM_REF) and os.access(PLUM_REF, os.X_OK)


PLUM_GPU = os.path.join(REPO, "bin", "plum_gpu")


def have_plum_gpu() -> bool:
    return os.path.exists(PLUM_GPU) and os.access(PLUM_GPU, os.X_OK)


def run_plum_ref(example_dir: str, steps: int, seed: int, xyz: bool = True, binary: str = None,
                 overrides: dict = None, want_files=(), extra_env: dict = None, stderr_to: list = None,
                 trace: bool = True) -> List[str]:
    """Run a driver binary (default: the reference, oracle/_ref/plum_ref; or bin/plum_gpu) on an example
    (inputs copied to a temp dir) and return its trace lines.  `overrides` replaces run.in values by key."""
    binary = binary or PLUM_REF
    overrides = dict(overrides or {})
    overrides["s1_total_simulation_steps"] = steps
    with tempfile.TemporaryDirectory(prefix="plum_ref_run_") as tmp:
        with open(os.path.join(example_dir, "run.in")) as f:
            text = f.read()
        out = []
        for ln in text.split("\n"):
            key = ln.split()[0] if ln.split() else ""
            if key in overrides:
                ln = f"{key} {overrides[key]}"
            out.append(ln)
        with open(os.path.join(tmp, "run.in"), "w") as f:
            f.write("\n".join(out))
        for fn in ("input_crd.dat", "input_top.dat"):
            with open(os.path.join(example_dir, fn)) as fi, open(os.path.join(tmp, fn), "w") as fo:
                fo.write(fi.read())
        env = dict(os.environ, PLUM_SEED=str(seed))
        if trace:
            env["PLUM_TRACE"] = os.path.join(tmp, "trace.txt")
        else:
            env.pop("PLUM_TRACE", None)      # timing runs: no per-step trace line (it reads the four totals every step)
        if xyz:
            env["PLUM_TRACE_XYZ"] = "1"
        env.update(extra_env or {})
        with open(os.path.join(tmp, "run.in")) as fin, open(os.path.join(tmp, "run.log"), "w") as fout:
            if stderr_to is None:
                subprocess.check_call([binary], stdin=fin, stdout=fout, cwd=tmp, env=env)
            else:     # the caller wants the binary's stderr (PLUM_B200_PROFILE=1 summary lines)
                p = subprocess.run([binary], stdin=fin, stdout=fout, stderr=subprocess.PIPE, cwd=tmp, env=env, text=True, check=True)
                stderr_to.append(p.stderr)
        lines = []
        if trace:
            with open(os.path.join(tmp, "trace.txt")) as f:
                lines = f.read().split("\n")
        if want_files:
            extra = {}
            for fn in want_files:
                with open(os.path.join(tmp, fn)) as f:
                    extra[fn] = f.read()
            return lines, extra
        return lines


@dataclasses.dataclass
class ReplayReport:
    n_moves: int = 0
    n_accept: int = 0
    n_gc: int = 0
    n_beads_energy: int = 0
    max_rel_dE: float = 0.0         # |dE - ref| / max(|ref|, scale)
    max_abs_dE: float = 0.0
    max_rel_tot: float = 0.0
    max_rel_beads: float = 0.0
    sentinel_mismatch: int = 0      # dE >= 1e8 on one side only
    worst: Optional[str] = None


def rel(a: float, b: float, floor: float = 1.0) -> float:
    return abs(a - b) / max(abs(b), floor)


def replay(engine, r: runin.RunIn, sysm: runin.System, types: runin.TypeTable, lines: List[str],
           max_steps: Optional[int] = None, check_totals_every: int = 1, beads_energy=True,
           on_step=None) -> ReplayReport:
    """`engine` needs: upload, init_energy, delta_e, commit, totals, insert_molecules,
    delete_molecules, and (optionally) beads_energy."""
    rep = ReplayReport()
    engine.upload(sysm.xyz, sysm.q, types.ids(sysm.symbol), sysm.mol_first)
    tot = engine.init_energy()
    mol_first = sysm.mol_first.copy()
    q = sysm.q.copy()
    gc_type = types.index.get(r.gc_bead_symbol, 0)
    pending = None
    step_count = 0
    for ln in lines:
        if not ln:
            continue
        f = ln.split()
        tag = f[0]
        if tag == "I":
            ref = dict(pair=hx(f[2]), ewald=hx(f[3]), bond=hx(f[4]), ext=hx(f[5]))
            for k in ref:
                e = rel(tot[k], ref[k])
                if e > rep.max_rel_tot:
                    rep.max_rel_tot = e
                    rep.worst = f"init {k}: {tot[k]!r} vs {ref[k]!r}"
        elif tag == "T":
            pending = f
        elif tag == "X":
            assert pending is not None
            t = pending
            pending = None
            step, mol, dE_ref, acc = int(t[1]), int(t[3]), hx(t[4]), int(t[5])
            if max_steps is not None and step > max_steps:
                break
            n = int(f[1])
            vals = f[2:]
            moved = np.array([int(vals[4 * i]) for i in range(n)], dtype=np.uint8)
            xyz = np.array([[hx(vals[4 * i + 1]), hx(vals[4 * i + 2]), hx(vals[4 * i + 3])] for i in range(n)])
            d = engine.delta_e(mol, xyz, moved)
            rep.n_moves += 1
            rep.n_accept += acc
            if (d["dE"] >= VLE) != (dE_ref >= VLE):
                rep.sentinel_mismatch += 1
                rep.worst = f"step {step}: sentinel mismatch {d['dE']!r} vs {dE_ref!r}"
            elif dE_ref < VLE:
                e = rel(d["dE"], dE_ref)
                rep.max_abs_dE = max(rep.max_abs_dE, abs(d["dE"] - dE_ref))
                if e > rep.max_rel_dE:
                    rep.max_rel_dE = e
                    rep.worst = f"step {step} mol {mol}: dE {d['dE']!r} vs {dE_ref!r}"
            engine.commit(bool(acc))
            step_count += 1
            if on_step is not None:
                on_step(step)   # the configuration Simulation::Sample sees after this step
            if check_totals_every and step_count % check_totals_every == 0:
                tot = engine.totals()
                ref = dict(pair=hx(t[6]), ewald=hx(t[7]), bond=hx(t[8]), ext=hx(t[9]))
                for k in ref:
                    e = rel(tot[k], ref[k])
                    if e > rep.max_rel_tot:
                        rep.max_rel_tot = e
                        rep.worst = f"step {step} total {k}: {tot[k]!r} vs {ref[k]!r}"
        elif tag == "B" and beads_energy and hasattr(engine, "beads_energy"):
            cur_len, delete_id = int(f[1]), int(f[2])
            b1 = [hx(f[3]), hx(f[4]), hx(f[5])]
            q1 = hx(f[6])
            b2 = [hx(f[7]), hx(f[8]), hx(f[9])]
            q2 = hx(f[10])
            e_ref = hx(f[13])
            nch = int(f[14])
            ch = np.array([hx(v) for v in f[15:15 + 4 * nch]]).reshape(-1, 4) if nch else np.zeros((0, 4))
            use2 = int(r.gc_bead_charge != 0)
            skip_first, skip_last = -1, -1
            if delete_id >= 0:
                skip_first = delete_id
                skip_last = delete_id + (r.gc_chain_len if r.gc_bead_charge != 0 else 0)
            e, pe, ee = engine.beads_energy(b1, gc_type, q1, b2, gc_type, q2, use2, ch[:, :3], ch[:, 3],
                                            np.full(max(nch, 1), gc_type, dtype=np.int32), cur_len,
                                            skip_first, skip_last)
            rep.n_beads_energy += 1
            if (e >= VLE) != (e_ref >= VLE):
                rep.sentinel_mismatch += 1
                rep.worst = f"BeadsEnergy sentinel mismatch {e!r} vs {e_ref!r}"
            elif e_ref < VLE:
                er = rel(e, e_ref)
                if er > rep.max_rel_beads:
                    rep.max_rel_beads = er
                    rep.worst = f"BeadsEnergy {e!r} vs {e_ref!r} (len {cur_len}, del {delete_id})"
        elif tag == "A":
            nb = int(f[1])
            vals = f[2:]
            mols, syms, qs, xyz = [], [], [], []
            for i in range(nb):
                mols.append(int(vals[6 * i]))
                syms.append(vals[6 * i + 1])
                qs.append(hx(vals[6 * i + 2]))
                xyz.append([hx(vals[6 * i + 3]), hx(vals[6 * i + 4]), hx(vals[6 * i + 5])])
            lens = []
            for m in mols:
                if not lens or m != last_m:
                    lens.append(0)
                lens[-1] += 1
                last_m = m
            engine.insert_molecules(lens, np.array(xyz), np.array(qs), types.ids(syms))
            q = np.concatenate([q, np.array(qs)])
            mol_first = np.concatenate([mol_first, mol_first[-1] + np.cumsum(lens)]).astype(np.int32)
        elif tag == "G":
            step = int(f[1])
            if max_steps is not None and step > max_steps:
                break
            rep.n_gc += 1
            kind, result = f[2], int(f[3])
            if kind == "D" and result >= 0:
                b0, b1_ = mol_first[result], mol_first[result + 1]
                ions = int(sum(abs(round(x)) for x in q[b0:b1_]))
                ml = result + ions
                nrem = int(mol_first[ml + 1] - mol_first[result])
                engine.delete_molecules(result, ml)
                q = np.concatenate([q[:b0], q[b0 + nrem:]])
                mol_first = np.concatenate([mol_first[:result + 1], mol_first[ml + 2:] - nrem]).astype(np.int32)
            tot = engine.totals()
            ref = dict(pair=hx(f[6]), ewald=hx(f[7]), bond=hx(f[8]), ext=hx(f[9]))
            for k in ref:
                e = rel(tot[k], ref[k])
                if e > rep.max_rel_tot:
                    rep.max_rel_tot = e
                    rep.worst = f"GC step {step} total {k}: {tot[k]!r} vs {ref[k]!r}"
    return rep


# ---------------------------------------------------------------- golden data
def golden_example_dir(name: str) -> str:
    return os.path.join(GOLDEN, "examples", name)


def load_golden(name: str):
    """(RunIn, System, TypeTable, params dict) of a committed example input."""
    r, s = runin.load_example(golden_example_dir(name))
    types = runin.TypeTable(r, s.symbol)
    return r, s, types, runin.params_dict(r, s.box, types)


def golden_short_trace(name: str, seed: int) -> List[str]:
    import gzip
    with gzip.open(os.path.join(GOLDEN, "short", f"{name}_seed{seed}.trace.gz"), "rt") as f:
        return f.read().split("\n")


def golden_pressure_fixture():
    """confined_nvt, seed 1, 3000 steps: (trace lines, {step: [six pressure columns]}) from the reference's
    own output_stat.dat (tests/golden/make_golden.py extras())."""
    import gzip
    with gzip.open(os.path.join(GOLDEN, "short", "confined_nvt_pressure_seed1.trace.gz"), "rt") as f:
        lines = f.read().split("\n")
    cols = {}
    with open(os.path.join(GOLDEN, "short", "confined_nvt_pressure_seed1.stat.dat")) as f:
        rows = [ln.split() for ln in f.read().split("\n") if ln]
    hdr = rows[0]
    i0 = hdr.index("<Pzz_LJ_ion>")
    for row in rows[1:]:
        cols[int(row[0])] = [float(x) for x in row[i0:i0 + 6]]
    return lines, cols


def golden_vol_pressure_fixture(name):
    """short/<name>_volp_seed1.trace.gz (make_golden.py vol_pressure()): (trace lines, {step: V-line values}) —
    the reference's accumulators after each CalcPressureVolScalingHSELSlit call: p_tensor[0..5], p_tensor_el[16],
    p_tensor_hs[16]."""
    import gzip
    with gzip.open(os.path.join(GOLDEN, "short", f"{name}_volp_seed1.trace.gz"), "rt") as f:
        lines = f.read().split("\n")
    ref, step = {}, 0
    for ln in lines:
        t = ln.split()
        if not t:
            continue
        if t[0] == "T":
            step = int(t[1])
        elif t[0] == "V":
            v = [hx(x) for x in t[2:]]
            ref[step] = {"n": int(t[1]), "p": np.array(v[:6]), "el": np.array(v[6:22]), "hs": np.array(v[22:38])}
    return [ln for ln in lines if not ln.startswith("V ")], ref


class VolScalingAverager:
    """Host-side bookkeeping of ForceField::CalcPressureVolScalingHSELSlit (pressure.cc:340-384) around an engine
    that returns the dU terms of one virtual stretch (mirrors plum_b200/host/force_field.cc)."""

    def __init__(self, box, beta, dz=1e-5):
        self.box, self.beta, self.dz = [float(x) for x in box], float(beta), float(dz)
        self.vp_z = 0
        self.p = np.zeros(6)
        self.el = np.zeros(16)
        self.hs = np.zeros(16)

    def add(self, smp):
        import math
        b, dz = self.box, self.dz
        self.vp_z += 1
        self.el += smp["el"]; self.hs += smp["hs"]
        self.p[5] += smp["bond"]; self.p[4] += smp["dipole"]
        vol = b[0] * b[1] * b[2]
        self.p[2] += smp["n_free"] / vol
        self.p[3] += math.pow(1.0 + dz / b[2], smp["n_free"]) * math.exp(-self.beta * smp["dU"])
        tot = 0.0
        for i in range(4):
            for j in range(i, 4):
                tot += self.el[i * 4 + j] + self.hs[i * 4 + j]
        tot += self.p[4] + self.p[5]
        tot /= (b[0] * b[1] * dz)
        self.p[0] = (self.beta * self.p[2] - tot) / self.vp_z
        self.p[1] = self.beta / (b[0] * b[1] * dz) * math.log(self.p[3] / self.vp_z)


class WallForceAverager:
    """Host-side bookkeeping of ForceField::CalcPressureForceLJELSlit (pressure.cc:404-484) around an
    engine that returns the six force sums of one configuration: wall-wall terms only from the first
    sample, running averages per unit area."""

    def __init__(self, box):
        self.area = box[0] * box[1]
        self.vp_z = 0
        self.cum = [0.0] * 6

    def add(self, f6):
        self.vp_z += 1
        for k in (0, 1, 3, 4):
            self.cum[k] += f6[k]
        if self.vp_z == 1:
            self.cum[2] += f6[2]
            self.cum[5] += f6[5]

    def p_tensor(self):
        a = self.area
        return [self.cum[0] / (self.vp_z * a), self.cum[1] / (self.vp_z * a), self.cum[2] / a,
                self.cum[3] / (self.vp_z * a), self.cum[4] / (self.vp_z * a), self.cum[5] / a]


def golden_long(name: str, seed: int = 1):
    return np.load(os.path.join(GOLDEN, "long", f"{name}_seed{seed}.npz"))
