"""GPU parity tests: the CUDA engine, called through the C ABI, against (a) traces written
by the real reference (tests/golden/short) and (b) the CPU oracle on seeded random systems.

Bar: FP64, |dE - ref| <= 1e-10 * max(1, |ref|) (BASELINE.json north_star: 1e-10 relative);
the >= 1e8 sentinel classification must agree exactly (it decides RNG consumption,
src/simulation/simulation.cc:324-332)."""
import numpy as np
import pytest

import replay
from plum_b200 import runin

pytestmark = pytest.mark.gpu

TOL = 1e-10
EXAMPLES = ["bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"]


def _engine(params, cap):
    from plum_b200.engine import Engine
    return Engine(params, device=0, capacity_beads=max(cap, 1))


def _oracle(params, mode=1):
    from oracle.oracle_py import Oracle
    return Oracle(params, repl_mode=mode)


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("seed", [1, 2])
def test_engine_replays_reference_trace(name, seed):
    r, s, types, params = replay.load_golden(name)
    eng = _engine(params, s.n + 64)
    rep = replay.replay(eng, r, s, types, replay.golden_short_trace(name, seed))
    assert rep.n_moves > 100
    assert rep.sentinel_mismatch == 0, rep.worst
    assert rep.max_rel_dE < TOL, rep.worst
    assert rep.max_rel_tot < TOL, rep.worst
    assert rep.max_rel_beads < TOL, rep.worst
    assert eng.launch_count() > rep.n_moves
    eng.close()


def test_ewald_setup_matches_oracle():
    for name in EXAMPLES:
        _, s, _, params = replay.load_golden(name)
        eng, orc = _engine(params, s.n), _oracle(params)
        a, b = eng.ewald_info(), orc.ewald_info()
        assert list(a.ewald_box) == list(b.ewald_box)
        assert (a.real_cutoff, a.repl_cutoff, a.box_vol) == (b.real_cutoff, b.repl_cutoff, b.box_vol)
        assert list(a.real_cell) == list(b.real_cell) and list(a.repl_cell) == list(b.repl_cell)
        assert (a.n_k, a.n_k_half) == (b.n_k, b.n_k_half)
        eng.close()


def test_structure_factor_matches_oracle():
    r, s, types, params = replay.load_golden("bulk_nvt")
    eng, orc = _engine(params, s.n), _oracle(params)
    ids = types.ids(s.symbol)
    eng.upload(s.xyz, s.q, ids, s.mol_first)
    orc.upload(s.xyz, s.q, ids, s.mol_first)
    eng.init_energy()
    sg, so = eng.sk_download(), orc.sk_half()
    assert sg.shape == so.shape
    assert np.max(np.abs(sg - so)) < 1e-10 * max(1.0, np.max(np.abs(so)))
    eng.close()


def _random_system(rng, n_chain, chain_len, n_ion, box, charged_every=1, slab=False, rmin=1.6):
    """Random chains (bond 2.5) + ions with a minimum separation: without it random beads overlap so
    strongly that single pair energies reach 1e7-1e8 kT and every dE is a difference of terms whose
    own rounding (ulp(1e8) = 1.5e-8) exceeds the 1e-10 bar — for the oracle and the reference alike."""
    pts = []

    def ok(p):
        if not pts:
            return True
        d = np.array(pts) - p
        for a in range(3 if not slab else 2):
            d[:, a] -= box[a] * np.round(d[:, a] / box[a])
        return float(np.min(np.einsum("ij,ij->i", d, d))) >= rmin * rmin

    xyz, q, sym, first = [], [], [], [0]
    for _ in range(n_chain):
        while True:
            p = rng.uniform(0.1, 0.9, 3) * box
            if ok(p):
                break
        for b in range(chain_len):
            pts.append(p.copy())
            xyz.append(p.copy())
            q.append(-1.0 if b % charged_every == 0 else 0.0)
            sym.append("P")
            for _try in range(200):
                step = rng.normal(size=3)
                c = p + 2.5 * step / np.linalg.norm(step)
                if slab:
                    c[2] = min(max(c[2], 1.2), box[2] - 1.2)
                if ok(c):
                    break
            p = c
        first.append(len(q))
    for _ in range(n_ion):
        while True:
            p = rng.uniform(0.05, 0.95, 3) * box
            if ok(p):
                break
        pts.append(p.copy())
        xyz.append(p)
        q.append(1.0)
        sym.append("C")
        first.append(len(q))
    return runin.System(np.array(xyz), np.array(q), sym, np.array(first, dtype=np.int32), list(box))


def _params(box, npbc=3, alpha=0.01, dipole=0, ext=0, bond=0, lj_cutoff=-1.0, pair="TruncatedLJ", ext_name="TruncatedLJWall",
            wall_cut=-1.0, graft=False):
    r = runin.RunIn()
    r.npbc = npbc
    r.beta = 1.0
    r.use_pair = 1
    r.pair_name = pair
    r.lj_cutoff = lj_cutoff
    r.lj_sigma = {"P": 2.2, "C": 1.8}
    r.lj_epsilon = {"P": 0.1, "C": 0.3}
    r.hs_radius = {"P": 1.0, "C": 0.8}
    r.use_ewald = 1
    r.ewald_name = "Coul"
    r.lB = 2.5
    r.alpha = alpha
    r.dipole_correction = dipole
    if bond:
        r.use_bond = 1
        r.bond_name = "Spring"
        r.bond_k = 30.0
        r.bond_r0 = 2.5
    if ext:
        r.use_ext = 1
        r.ext_name = ext_name
        r.wall_cut = wall_cut
        r.wall_sigma = {"P": 1.0, "C": 0.9, "L": 1.0, "R": 1.0}
        r.wall_epsilon = {"P": 0.1, "C": 0.2, "L": 0.15, "R": 0.15}
        r.well_width = 2.0
        r.well_depth = -0.7
    if graft:
        r.lj_sigma.update({"L": 2.2, "R": 2.2})
        r.lj_epsilon.update({"L": 0.1, "R": 0.1})
    types = runin.TypeTable(r, ["P", "C"] + (["L", "R"] if graft else []))
    return r, types, runin.params_dict(r, box, types)


CASES = [
    dict(box=[30.0, 30.0, 30.0], alpha=0.02, n_chain=6, chain_len=12, n_ion=40),                       # multi-image real space
    dict(box=[60.0, 60.0, 60.0], alpha=0.05, n_chain=5, chain_len=40, n_ion=60, charged_every=4),      # single image, long chains
    dict(box=[25.0, 25.0, 40.0], alpha=0.02, n_chain=4, chain_len=10, n_ion=30, npbc=2, dipole=1, ext=1, slab=True),
    dict(box=[40.0, 40.0, 40.0], alpha=0.03, n_chain=4, chain_len=70, n_ion=20, bond=1),               # > 2 group chunks, springs
    dict(box=[35.0, 35.0, 35.0], alpha=0.02, n_chain=5, chain_len=9, n_ion=33, lj_cutoff=6.0),         # attractive LJ
    dict(box=[35.0, 35.0, 35.0], alpha=0.02, n_chain=5, chain_len=9, n_ion=33, pair="HardSphere"),
    dict(box=[25.0, 25.0, 40.0], alpha=0.02, n_chain=4, chain_len=10, n_ion=30, npbc=2, dipole=1, ext=1, slab=True,
         ext_name="HardWall"),
    dict(box=[25.0, 25.0, 40.0], alpha=0.02, n_chain=4, chain_len=10, n_ion=30, npbc=2, dipole=1, ext=1, slab=True,
         ext_name="WellWall"),
    dict(box=[25.0, 25.0, 40.0], alpha=0.02, n_chain=4, chain_len=10, n_ion=30, npbc=2, dipole=1, ext=1, slab=True,
         wall_cut=3.0),                                                                                   # full 9-3 LJ wall
    dict(box=[25.0, 25.0, 40.0], alpha=0.02, n_chain=4, chain_len=10, n_ion=30, npbc=2, dipole=1, ext=1, slab=True,
         graft=True),                                                                                     # L / R tethered beads (FENE branch)
    dict(box=[30.0, 30.0, 30.0], alpha=0.02, n_chain=3, chain_len=300, n_ion=40, charged_every=3),      # group > one staging chunk, many charged
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_random_moves_match_oracle(case):
    c = dict(CASES[case])
    rng = np.random.default_rng(100 + case)
    box = c.pop("box")
    sysm = _random_system(rng, c.pop("n_chain"), c.pop("chain_len"), c.pop("n_ion"), np.array(box),
                          c.pop("charged_every", 1), c.pop("slab", False))
    r, types, params = _params(box, **c)
    if c.get("graft"):
        for m in range(4):
            f = int(sysm.mol_first[m])
            sysm.symbol[f] = "L" if m % 2 == 0 else "R"
            sysm.xyz[f, 2] = 1.3 if m % 2 == 0 else box[2] - 1.3
    eng, orc = _engine(params, sysm.n), _oracle(params)
    ids = types.ids(sysm.symbol)
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    orc.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    tg, to = eng.init_energy(), orc.init_energy()
    for k in ("pair", "ewald", "bond", "ext", "real", "recip", "self", "dipole"):
        if params["pair_kind"] == 2 and k == "pair":
            continue  # random hard spheres overlap: sums of 1e8 sentinels, compared via n_overlap below
        assert abs(tg[k] - to[k]) <= TOL * max(1.0, abs(to[k])), (k, tg[k], to[k])
    n_acc = 0
    for it in range(60):
        mol = int(rng.integers(0, sysm.n_mol))
        f, l = int(sysm.mol_first[mol]), int(sysm.mol_first[mol + 1])
        cur = orc.positions()[f:l]
        kind = it % 4
        moved = np.ones(l - f, dtype=np.uint8)
        if kind == 0 or l - f == 1:
            trial = cur + rng.normal(scale=0.4, size=3)          # rigid translation
        elif kind == 1:
            trial = cur + rng.normal(scale=0.3, size=cur.shape)   # every bead displaced
        elif kind == 2:
            trial = cur.copy()                                    # crankshaft-like: a sub-range moves
            a, b = sorted(rng.integers(0, l - f, 2))
            moved[:] = 0
            moved[a:b + 1] = 1
            trial[a:b + 1] += rng.normal(scale=0.3, size=(b + 1 - a, 3))
        else:
            trial = cur + rng.normal(scale=3.0, size=3) * np.array([1, 1, 4.0])  # large: may leave a slab
        dg, do = eng.delta_e(mol, trial, moved), orc.delta_e(mol, trial, moved)
        assert (dg["dE"] >= 1e8) == (do["dE"] >= 1e8), (it, dg, do)
        assert dg["stage"] == do["stage"], (it, dg, do)
        if do["dE"] < 1e8:
            for k in ("dE", "pair", "ext", "ewald", "bond", "real", "recip"):
                assert abs(dg[k] - do[k]) <= TOL * max(1.0, abs(do[k])), (it, k, dg, do)
        else:
            assert dg["n_overlap"] == do["n_overlap"] or do["stage"] == 2
        acc = bool(do["dE"] < 1e8 and rng.random() < np.exp(-min(50.0, max(-50.0, do["dE"]))))
        n_acc += acc
        eng.commit(acc)
        orc.commit(acc)
    assert n_acc > 3
    tg, to = eng.totals(), orc.totals()
    for k in ("pair", "ewald", "bond", "ext"):
        if params["pair_kind"] == 2 and k == "pair":
            continue
        assert abs(tg[k] - to[k]) <= TOL * max(1.0, abs(to[k])), (k, tg[k], to[k])
    assert np.max(np.abs(eng.positions() - orc.positions())) == 0.0
    # drift check: accumulated totals vs a fresh full recompute on the device
    # (the slab dipole term is excluded: the reference's running total lags it by one accepted
    #  move on purpose — SURVEY.md §0.5 — so only the fresh value is "true")
    fresh = eng.recompute_totals()
    fresh["ewald"] -= fresh["dipole"]
    tg["ewald"] -= tg["dipole"]
    for k in ("ewald", "bond", "ext"):
        assert abs(fresh[k] - tg[k]) <= 1e-9 * max(1.0, abs(tg[k])), (k, fresh[k], tg[k])
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [0, 1, 2, 4, 5, 6, 9])
def test_batched_growth_steps_match_oracle(case):
    """pg_trial_energies (one k_trials launch per growth step) against the oracle's BeadsEnergy
    (cbmc.cc:5-151) trial by trial: inline and staged inputs, more trials than one launch holds,
    skipped molecules (deletion), partial chains with and without the counter-ion, trial beads on top
    of resident beads (1e8 sentinel) and outside a slab."""
    c = dict(CASES[case])
    rng = np.random.default_rng(500 + case)
    box = c.pop("box")
    sysm = _random_system(rng, c.pop("n_chain"), c.pop("chain_len"), c.pop("n_ion"), np.array(box),
                          c.pop("charged_every", 1), c.pop("slab", False))
    r, types, params = _params(box, **c)
    eng, orc = _engine(params, sysm.n), _oracle(params)
    ids = types.ids(sysm.symbol)
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    orc.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    eng.init_energy(); orc.init_energy()
    tP, tC = types.ids(["P"])[0], types.ids(["C"])[0]
    b = np.array(box)
    #        n_trials, current_len, use_bead2, skip range
    plans = [(1, 0, 1, None), (7, 3, 1, None), (30, 8, 1, None), (30, 5, 0, None), (70, 2, 1, None),
             (12, 12, 1, None), (9, 20, 0, (1, 2)), (30, 4, 1, (0, 0)), (33, 0, 0, (2, 3))]
    for nt, cl, use2, skip in plans:
        n_chain = cl * (2 if use2 else 1)
        # a partial chain: a random walk of monomers (bond 2.5) and, with use2, one ion near each
        cx = np.zeros((max(n_chain, 1), 3)); cq = np.zeros(max(n_chain, 1)); ct = np.zeros(max(n_chain, 1), dtype=np.int32)
        p = rng.uniform(0.3, 0.7, 3) * b
        for i in range(cl):
            cx[i] = p; cq[i] = -1.0; ct[i] = tP
            step = rng.normal(size=3); p = p + 2.5 * step / np.linalg.norm(step)
            if use2:
                cx[cl + i] = cx[i] + rng.normal(scale=2.0, size=3); cq[cl + i] = 1.0; ct[cl + i] = tC
        anchor = cx[cl - 1] if cl else rng.uniform(0.3, 0.7, 3) * b
        d = rng.normal(size=(nt, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
        b1 = anchor + 2.5 * d
        b2 = b1 + rng.normal(scale=2.0, size=(nt, 3))
        if nt >= 7:
            b1[3] = sysm.xyz[5] + 1e-3          # overlap with a resident bead: LJ huge / hard-sphere sentinel
            b2[4] = sysm.xyz[-1]                 # exactly on top of an ion: r = 0
            b1[5, 2] = -0.5                      # below the lower wall of a slab
        sf, sl = skip if skip else (-1, -1)
        eg, pg_, wg = eng.trial_energies(b1, b2, tP, -1.0, tC, 1.0, use2, cx[:max(n_chain, 1)], cq, ct, cl, sf, sl)
        for t in range(nt):
            eo, po, wo = orc.beads_energy(b1[t], tP, -1.0, b2[t], tC, 1.0, use2, cx, cq, ct, cl, sf, sl)
            tag = (case, nt, cl, use2, skip, t)
            if eo >= 1e8:
                assert eg[t] >= 1e8, tag
                continue
            assert abs(eg[t] - eo) <= TOL * max(1.0, abs(eo)), tag + (eg[t], eo)
            assert abs(pg_[t] - po) <= TOL * max(1.0, abs(po)), tag + (pg_[t], po)
            assert abs(wg[t] - wo) <= TOL * max(1.0, abs(wo)), tag + (wg[t], wo)
    eng.close()


def test_edge_cases():
    """r == 0 overlap, bead exactly on a cutoff, q == 0 partners, single bead system, empty system."""
    box = [30.0, 30.0, 30.0]
    r, types, params = _params(box, alpha=0.02)
    # two ions + one neutral bead
    xyz = np.array([[1.0, 1.0, 1.0], [4.0, 1.0, 1.0], [8.0, 8.0, 8.0]])
    q = np.array([1.0, -1.0, 0.0])
    sysm = runin.System(xyz, q, ["C", "C", "P"], np.array([0, 1, 2, 3], dtype=np.int32), box)
    eng, orc = _engine(params, 8), _oracle(params)
    ids = types.ids(sysm.symbol)
    for e in (eng, orc):
        e.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
        e.init_energy()
    # exact overlap: r <= 0 -> 1e8 sentinel (potential_truncated_lj.cc:59-64)
    dg, do = eng.delta_e(0, xyz[1:2], [1]), orc.delta_e(0, xyz[1:2], [1])
    assert dg["dE"] >= 1e8 and do["dE"] >= 1e8 and dg["stage"] == do["stage"] == 1
    assert dg["n_overlap"] == do["n_overlap"] == 1
    eng.commit(False); orc.commit(False)
    # exactly at the WCA cutoff r = k216*sigma
    sig = params["lj_sigma"][types.index["C"]]
    trial = xyz[1:2] - np.array([[1.12246204831 * sig, 0, 0]])
    dg, do = eng.delta_e(0, trial, [1]), orc.delta_e(0, trial, [1])
    assert abs(dg["dE"] - do["dE"]) <= TOL * max(1.0, abs(do["dE"]))
    eng.commit(True); orc.commit(True)
    # neutral bead moves: Ewald terms vanish identically
    dg, do = eng.delta_e(2, [[9.0, 8.5, 8.0]], [1]), orc.delta_e(2, [[9.0, 8.5, 8.0]], [1])
    assert dg["ewald"] == 0.0 and do["ewald"] == 0.0
    assert abs(dg["dE"] - do["dE"]) <= TOL
    eng.commit(True); orc.commit(True)
    eng.close()
    # empty system: totals are zero, nothing to move
    eng = _engine(params, 4)
    eng.upload(np.zeros((0, 3)), np.zeros(0), np.zeros(0, dtype=np.int32), np.array([0], dtype=np.int32))
    t = eng.init_energy()
    assert all(v == 0.0 for v in t.values())
    eng.close()


def test_insert_delete_round_trip():
    """Insert a chain + ions, then delete it again: totals and S(k) return to where they were."""
    rng = np.random.default_rng(5)
    box = [30.0, 30.0, 30.0]
    r, types, params = _params(box, alpha=0.02)
    sysm = _random_system(rng, 3, 8, 24, np.array(box))
    eng, orc = _engine(params, sysm.n), _oracle(params)
    ids = types.ids(sysm.symbol)
    for e in (eng, orc):
        e.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    t0 = eng.init_energy()
    orc.init_energy()
    sk0 = eng.sk_download()
    new = _random_system(rng, 1, 8, 8, np.array(box))
    lens = np.diff(new.mol_first)
    ag = eng.insert_molecules(lens, new.xyz, new.q, types.ids(new.symbol))
    ao = orc.insert_molecules(lens, new.xyz, new.q, types.ids(new.symbol))
    for k in ("pair", "ewald", "real", "recip"):
        assert abs(ag[k] - ao[k]) <= TOL * max(1.0, abs(ao[k])), (k, ag, ao)
    assert eng.n == orc.n == sysm.n + 16
    fresh = eng.recompute_totals()
    tg = eng.totals()
    for k in ("pair", "ewald"):
        assert abs(fresh[k] - tg[k]) <= 1e-9 * max(1.0, abs(tg[k])), (k, fresh, tg)
    # a move after insertion still agrees
    mol = sysm.n_mol
    cur = orc.positions()[sysm.n:sysm.n + 8]
    dg = eng.delta_e(mol, cur + 0.2, np.ones(8, dtype=np.uint8))
    do = orc.delta_e(mol, cur + 0.2, np.ones(8, dtype=np.uint8))
    assert abs(dg["dE"] - do["dE"]) <= TOL * max(1.0, abs(do["dE"]))
    eng.commit(False); orc.commit(False)
    rg = eng.delete_molecules(sysm.n_mol, sysm.n_mol + 8)
    ro = orc.delete_molecules(sysm.n_mol, sysm.n_mol + 8)
    for k in ("pair", "ewald", "real", "recip"):
        assert abs(rg[k] - ro[k]) <= TOL * max(1.0, abs(ro[k])), (k, rg, ro)
    t1 = eng.totals()
    for k in ("pair", "ewald"):
        assert abs(t1[k] - t0[k]) <= 1e-9 * max(1.0, abs(t0[k]))
    assert np.max(np.abs(eng.sk_download() - sk0)) < 1e-9
    assert np.array_equal(eng.positions(), sysm.xyz)
    # delete from the middle (compaction): remove the first chain, compare a move with the oracle
    eng.delete_molecules(0, 0)
    orc.delete_molecules(0, 0)
    assert np.array_equal(eng.positions(), orc.positions())
    cur = orc.positions()[0:8]
    dg = eng.delta_e(0, cur - 0.3, np.ones(8, dtype=np.uint8))
    do = orc.delta_e(0, cur - 0.3, np.ones(8, dtype=np.uint8))
    assert abs(dg["dE"] - do["dE"]) <= TOL * max(1.0, abs(do["dE"]))
    eng.close()


def test_replay_on_device_matches_host_driven_path():
    """The device-resident replay (bench `value` leg) reproduces the host-driven sequence bit for bit."""
    rng = np.random.default_rng(9)
    box = [40.0, 40.0, 40.0]
    r, types, params = _params(box, alpha=0.03)
    sysm = _random_system(rng, 6, 20, 40, np.array(box), charged_every=2)
    ids = types.ids(sysm.symbol)
    eng = _engine(params, sysm.n)
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    eng.init_energy()
    mols, offs, us, xyzs, mvs, dEs, accs = [], [], [], [], [], [], []
    off = 0
    pos = sysm.xyz.copy()
    for it in range(40):
        mol = int(rng.integers(0, sysm.n_mol))
        f, l = int(sysm.mol_first[mol]), int(sysm.mol_first[mol + 1])
        trial = pos[f:l] + rng.normal(scale=0.3, size=(l - f, 3))
        d = eng.delta_e(mol, trial, np.ones(l - f, dtype=np.uint8))
        u = rng.random()
        acc = bool(d["dE"] < 1e8 and u < np.exp(-min(700.0, max(-700.0, d["dE"]))))
        eng.commit(acc)
        if acc:
            pos[f:l] = trial
        mols.append(mol); offs.append(off); us.append(u); xyzs.append(trial); mvs.append(np.ones(l - f, dtype=np.uint8))
        dEs.append(d["dE"]); accs.append(acc)
        off += l - f
    final_host = eng.totals()
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    eng.init_energy()
    eng.replay_upload(mols, offs, us, np.concatenate(xyzs), np.concatenate(mvs))
    dE, acc, ms = eng.replay_run(0, len(mols))
    assert np.array_equal(dE, np.array(dEs))
    assert np.array_equal(acc.astype(bool), np.array(accs))
    assert eng.totals() == final_host
    assert np.array_equal(eng.positions(), pos)
    assert ms > 0
    eng.close()


@pytest.mark.parametrize("name", ["confined_nvt", "confined_muvt"])
def test_wall_force_matches_oracle_on_examples(name):
    """pg_wall_force (CalcPressureForceLJELSlit, pressure.cc:404-469) on the slab examples: the six force
    sums against the oracle (itself pinned to the reference's printed columns, test_oracle_vs_reference)
    at the start and after replayed reference moves."""
    r, s, types, params = replay.load_golden(name)
    eng, orc = _engine(params, s.n + 64), _oracle(params)
    lines = replay.golden_short_trace(name, 1)
    seen = []

    def check(step):
        if step % 40 == 0:
            g, o = eng.wall_force(r.phantom), orc.wall_force(r.phantom)
            scale = max(1.0, float(np.max(np.abs(o))))
            assert np.max(np.abs(g - o)) <= TOL * scale, (step, g, o)
            seen.append(step)

    class Both:
        """Drives engine and oracle through the same trace."""
        def upload(self, *a): eng.upload(*a); orc.upload(*a)
        def init_energy(self): orc.init_energy(); return eng.init_energy()
        def delta_e(self, *a): orc.delta_e(*a); return eng.delta_e(*a)
        def commit(self, acc): orc.commit(acc); eng.commit(acc)
        def totals(self): return eng.totals()
        def insert_molecules(self, *a): orc.insert_molecules(*a); return eng.insert_molecules(*a)
        def delete_molecules(self, *a): orc.delete_molecules(*a); return eng.delete_molecules(*a)

    replay.replay(Both(), r, s, types, lines, max_steps=200, check_totals_every=0, beads_energy=False, on_step=check)
    assert len(seen) >= 3
    eng.close()


@pytest.mark.parametrize("case", [2, 6, 7, 8, 9])
def test_wall_force_plate_term_matches_oracle(case):
    """No wall sites (phantom = 0): only 0.5 * BeadForceOnWall is left — repulsive, full 9-3 (wall_cut),
    grafted L/R (FENE) branches, and the force-free hard / well walls."""
    c = dict(CASES[case])
    rng = np.random.default_rng(900 + case)
    box = c.pop("box")
    sysm = _random_system(rng, c.pop("n_chain"), c.pop("chain_len"), c.pop("n_ion"), np.array(box),
                          c.pop("charged_every", 1), c.pop("slab", False))
    r, types, params = _params(box, **c)
    if c.get("graft"):
        for m in range(4):
            f = int(sysm.mol_first[m])
            sysm.symbol[f] = "L" if m % 2 == 0 else "R"
            sysm.xyz[f, 2] = 1.3 if m % 2 == 0 else box[2] - 1.3
    sysm.xyz[-3:, 2] = [0.6, box[2] - 0.7, 1.0]     # ions inside the repulsive range of both plates
    eng, orc = _engine(params, sysm.n), _oracle(params)
    ids = types.ids(sysm.symbol)
    eng.upload(sysm.xyz, sysm.q, ids, sysm.mol_first); orc.upload(sysm.xyz, sysm.q, ids, sysm.mol_first)
    eng.init_energy(); orc.init_energy()
    g, o = eng.wall_force(0), orc.wall_force(0)
    assert np.max(np.abs(g - o)) <= TOL * max(1.0, float(np.max(np.abs(o)))), (g, o)
    if params["ext_kind"] == 1:
        assert abs(o[0]) > 0
    else:
        assert not o.any()
    # four ions promoted to wall sites (they are single-bead molecules): the site-site terms, wall-wall included
    perm_first = list(range(sysm.n_mol - 4, sysm.n_mol)) + list(range(sysm.n_mol - 4))
    xyz = np.concatenate([sysm.xyz[sysm.mol_first[m]:sysm.mol_first[m + 1]] for m in perm_first])
    q = np.concatenate([sysm.q[sysm.mol_first[m]:sysm.mol_first[m + 1]] for m in perm_first])
    sym = [sysm.symbol[b] for m in perm_first for b in range(sysm.mol_first[m], sysm.mol_first[m + 1])]
    lens = [int(sysm.mol_first[m + 1] - sysm.mol_first[m]) for m in perm_first]
    mf = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    xyz[:2, 2] = 0.0; xyz[2:4, 2] = box[2]
    ids2 = types.ids(sym)
    eng.upload(xyz, q, ids2, mf); orc.upload(xyz, q, ids2, mf)
    eng.init_energy(); orc.init_energy()
    g, o = eng.wall_force(4), orc.wall_force(4)
    assert np.max(np.abs(g - o)) <= TOL * max(1.0, float(np.max(np.abs(o)))), (g, o)
    assert abs(o[3]) > 0 and abs(o[5]) > 0
    eng.close()


def _vol_close(g, o, scale):
    """dU tables of one virtual stretch: every entry is a sum of (stretched - current) energies, so the bar is
    1e-10 of the energy being differenced (TOL * scale); entries the oracle leaves at zero must be zero."""
    for k in ("el", "hs"):
        assert np.max(np.abs(g[k] - o[k])) <= TOL * scale, (k, g[k], o[k])
        assert not g[k][o[k] == 0.0].any(), (k, g[k], o[k])
    for k in ("bond", "dipole", "dU"):
        assert abs(g[k] - o[k]) <= TOL * max(scale, abs(o[k]) * 1e-2), (k, g[k], o[k])
    assert g["n_free"] == o["n_free"]


@pytest.mark.parametrize("name", ["bulk_nvt", "synth_spring"])
def test_vol_scaling_sample_matches_reference_accumulators(name):
    """pg_vol_scaling_sample (CalcPressureVolScalingHSELSlit, pressure.cc:187-387) on the reference's own trace: the
    accumulators the reference holds after its samples at steps 200 and 300 (V lines of the fixture) from the GPU
    samples, and every sample against the oracle's pairwise evaluation."""
    r, s, types, params = replay.load_golden(name)
    lines, ref = replay.golden_vol_pressure_fixture(name)
    eng, orc = _engine(params, s.n + 64), _oracle(params)
    avg = replay.VolScalingAverager(s.box, r.beta)
    seen = []

    def check(step):
        if step in ref:
            g, o = eng.vol_scaling_sample(r.phantom), orc.vol_scaling_sample(r.phantom)
            tot = orc.totals()
            scale = max(1.0, abs(tot["ewald"]), abs(tot["pair"]), abs(tot["bond"]))
            _vol_close(g, o, scale)
            avg.add(g)
            w = ref[step]
            assert np.max(np.abs(avg.el - w["el"])) <= TOL * scale and np.max(np.abs(avg.hs - w["hs"])) <= TOL * scale
            assert np.max(np.abs(avg.p[4:6] - w["p"][4:6])) <= TOL * scale
            for k in (0, 1):
                assert abs(avg.p[k] - w["p"][k]) <= 1e-6 * abs(w["p"][k]), (step, k, avg.p[k], w["p"][k])
            print(f"{name} step {step}: P_yethiraj {avg.p[0]:.9e} (ref {w['p'][0]:.9e}), rel err "
                  f"{abs(avg.p[0] - w['p'][0]) / abs(w['p'][0]):.2e}; max |d el| / scale "
                  f"{np.max(np.abs(avg.el - w['el'])) / scale:.2e}")
            seen.append(step)

    class Both:
        def upload(self, *a): eng.upload(*a); orc.upload(*a)
        def init_energy(self): orc.init_energy(); return eng.init_energy()
        def delta_e(self, *a): orc.delta_e(*a); return eng.delta_e(*a)
        def commit(self, acc): orc.commit(acc); eng.commit(acc)
        def totals(self): return eng.totals()

    replay.replay(Both(), r, s, types, lines, check_totals_every=0, beads_energy=False, on_step=check)
    assert seen == [200, 300]
    eng.close()


@pytest.mark.parametrize("case", [0, 1, 2, 3, 4, 5, 8, 9])
def test_vol_scaling_sample_matches_oracle_on_random_systems(case):
    """Multi-image and single-image real space, slab + dipole term + walls (the wall branch of pressure.cc:300-309,
    never reached by the reference's own call site but part of the function), springs, attractive LJ, hard spheres
    (stored energy 0 for bonded neighbours), grafted L/R beads; then with four ions promoted to surface sites so that
    the first-plate / second-plate restriction of pressure.cc:264-265 is exercised."""
    c = dict(CASES[case])
    rng = np.random.default_rng(1200 + case)
    box = c.pop("box")
    sysm = _random_system(rng, c.pop("n_chain"), c.pop("chain_len"), c.pop("n_ion"), np.array(box),
                          c.pop("charged_every", 1), c.pop("slab", False))
    r, types, params = _params(box, **c)
    if c.get("graft"):
        for m in range(4):
            f = int(sysm.mol_first[m])
            sysm.symbol[f] = "L" if m % 2 == 0 else "R"
            sysm.xyz[f, 2] = 1.3 if m % 2 == 0 else box[2] - 1.3
    eng, orc = _engine(params, sysm.n), _oracle(params)

    def both(xyz, q, ids, mf, phantom):
        eng.upload(xyz, q, ids, mf); orc.upload(xyz, q, ids, mf)
        eng.init_energy(); to = orc.init_energy()
        g, o = eng.vol_scaling_sample(phantom), orc.vol_scaling_sample(phantom)
        pair = abs(to["pair"]) if abs(to["pair"]) < 1e7 else 1.0      # hard-sphere overlaps: 1e8 sentinels
        scale = max(1.0, abs(to["ewald"]), abs(to["real"]), abs(to["recip"]), pair, abs(to["bond"]), abs(to["ext"]) if abs(to["ext"]) < 1e7 else 1.0)
        big = np.abs(o["hs"]) >= 1e7
        if big.any():    # sentinel differences: exact multiples of 1e8
            assert np.array_equal(np.round(g["hs"][big] / 1e8), np.round(o["hs"][big] / 1e8))
            g["hs"][big] = o["hs"][big] = 0.0
            g["dU"] = o["dU"] = 0.0
        _vol_close(g, o, scale)
        return o

    o0 = both(sysm.xyz, sysm.q, types.ids(sysm.symbol), sysm.mol_first, 0)
    assert np.count_nonzero(o0["el"]) >= 3
    if params["bond_kind"]:
        assert o0["bond"] != 0.0
    if params["dipole_correction"]:
        assert o0["dipole"] != 0.0
    perm_first = list(range(sysm.n_mol - 4, sysm.n_mol)) + list(range(sysm.n_mol - 4))
    xyz = np.concatenate([sysm.xyz[sysm.mol_first[m]:sysm.mol_first[m + 1]] for m in perm_first])
    q = np.concatenate([sysm.q[sysm.mol_first[m]:sysm.mol_first[m + 1]] for m in perm_first])
    sym = [sysm.symbol[b] for m in perm_first for b in range(sysm.mol_first[m], sysm.mol_first[m + 1])]
    lens = [int(sysm.mol_first[m + 1] - sysm.mol_first[m]) for m in perm_first]
    mf = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    o4 = both(xyz, q, types.ids(sym), mf, 4)
    assert o4["el"][15] != 0.0 and o4["n_free"] == sysm.n_mol - 4      # plate-plate term, surface x mobile terms
    assert any(o4["el"][i] != 0.0 for i in (3, 7, 11))
    eng.close()


def test_full_size_synthetic_system_parity():
    """BASELINE.json configs[4] at FULL size (22 000 beads, K = 3574): the FAST k_move path and k_trials
    against the oracle on a bounded sample — init totals, three ion moves, one 100-bead chain move
    (pivot-like, every bead displaced), one accepted move followed by another (deferred commit), and a
    CBMC growth step of 6 trials; plus size-independent properties: a rejected move leaves the totals
    untouched, an accepted move followed by its inverse returns them."""
    import os
    from plum_b200 import synth
    r, s, types, params = synth.load(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    ids = types.ids(s.symbol)
    eng, orc = _engine(params, s.n), _oracle(params)
    eng.upload(s.xyz, s.q, ids, s.mol_first); orc.upload(s.xyz, s.q, ids, s.mol_first)
    tg = eng.init_energy()
    # oracle totals of 2.4e8 pairs would take minutes: check the O(N) / O(N_q K) components only
    sk_o = orc.sk_half()
    assert np.max(np.abs(eng.sk_download() - sk_o)) <= 1e-10 * max(1.0, float(np.max(np.abs(sk_o))))
    rng = np.random.default_rng(5)
    ions = [m for m in range(s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] == 1]
    chains = [m for m in range(s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] > 1]
    pos = s.xyz.copy()

    def one(mol, trial, accept):
        f, l = int(s.mol_first[mol]), int(s.mol_first[mol + 1])
        mv = np.ones(l - f, dtype=np.uint8)
        dg, do = eng.delta_e(mol, trial, mv), orc.delta_e(mol, trial, mv)
        for k in ("dE", "pair", "ewald", "real", "recip"):
            assert abs(dg[k] - do[k]) <= TOL * max(1.0, abs(do[k])), (mol, k, dg[k], do[k])
        eng.commit(accept); orc.commit(accept)
        if accept:
            pos[f:l] = trial
        return dg["dE"]

    t0 = eng.totals()
    for i in range(3):
        m = int(rng.choice(ions)); f = int(s.mol_first[m])
        v = rng.normal(size=3)
        one(m, pos[f:f + 1] + 6.0 * v / np.linalg.norm(v), accept=(i == 1))
    m = int(rng.choice(chains)); f, l = int(s.mol_first[m]), int(s.mol_first[m + 1])
    keep = pos[f:l].copy()
    dE1 = one(m, keep + rng.normal(scale=0.4, size=(l - f, 3)), accept=True)      # accepted chain move
    t1 = eng.totals()
    dE2 = one(m, keep, accept=True)                                                # ... and its inverse
    t2 = eng.totals()
    assert abs(dE1 + dE2) <= 1e-9 * max(1.0, abs(dE1))
    etot = lambda t: t["pair"] + t["ewald"] + t["bond"] + t["ext"]
    assert abs(etot(t2) - (etot(t1) - dE1)) <= 1e-9 * max(1.0, abs(etot(t1)))
    assert np.max(np.abs(eng.positions() - pos)) == 0.0
    # CBMC growth step at full size (k_trials, single-image column split = 1)
    tP = types.ids(["P"])[0]
    cl = 3
    cx = np.zeros((2 * cl, 3)); cq = np.zeros(2 * cl); ct = np.full(2 * cl, tP, dtype=np.int32)
    p = np.array([100.0, 100.0, 100.0])
    for i in range(cl):
        cx[i] = p + 2.5 * i; cq[i] = -1.0
        cx[cl + i] = cx[i] + 1.9; cq[cl + i] = 1.0
    b1 = cx[cl - 1] + 2.5 * np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0.6, 0.8, 0]])
    b2 = b1 + np.array([1.5, -1.2, 0.7])
    eg, pg_, wg = eng.trial_energies(b1, b2, tP, -1.0, tP, 1.0, 1, cx, cq, ct, cl, chains[0], chains[0])
    for t in range(len(b1)):
        eo, po, wo = orc.beads_energy(b1[t], tP, -1.0, b2[t], tP, 1.0, 1, cx, cq, ct, cl, chains[0], chains[0])
        assert abs(eg[t] - eo) <= TOL * max(1.0, abs(eo)), (t, eg[t], eo)
        assert abs(wg[t] - wo) <= TOL * max(1.0, abs(wo)), (t, wg[t], wo)
    eng.close()


def test_async_halves_interleave_two_replicas():
    """pg_delta_e_begin / pg_delta_e_poll: two engines driven by one thread with their round trips
    overlapped give exactly the numbers of the synchronous pg_delta_e, and state errors are reported."""
    from plum_b200.engine import EngineError
    rng = np.random.default_rng(11)
    box = [40.0, 40.0, 40.0]
    r, types, params = _params(box, alpha=0.03)
    sysm = _random_system(rng, 5, 14, 50, np.array(box))
    ids = types.ids(sysm.symbol)
    engs = [_engine(params, sysm.n) for _ in range(3)]
    for e in engs:
        e.upload(sysm.xyz, sysm.q, ids, sysm.mol_first); e.init_energy()
    a, b, ref = engs
    pos = [sysm.xyz.copy(), sysm.xyz.copy()]
    for it in range(40):
        props = []
        for k, e in enumerate((a, b)):
            mol = int(rng.integers(0, sysm.n_mol))
            f, l = int(sysm.mol_first[mol]), int(sysm.mol_first[mol + 1])
            trial = pos[k][f:l] + rng.normal(scale=0.4, size=(l - f, 3))
            e.delta_e_begin(mol, trial, np.ones(l - f, dtype=np.uint8))
            props.append((mol, f, l, trial))
        with pytest.raises(EngineError):
            a.delta_e_begin(0, pos[0][:int(sysm.mol_first[1])], np.ones(int(sysm.mol_first[1]), dtype=np.uint8))   # one in flight already
        out = [None, None]
        while out[0] is None or out[1] is None:
            for k, e in enumerate((a, b)):
                if out[k] is None:
                    out[k] = e.delta_e_poll()
        for k, e in enumerate((a, b)):
            mol, f, l, trial = props[k]
            # the third engine replays replica k's history synchronously
            ref.upload(pos[k], sysm.q, ids, sysm.mol_first); ref.init_energy()
            d = ref.delta_e(mol, trial, np.ones(l - f, dtype=np.uint8)); ref.commit(False)
            assert abs(out[k]["dE"] - d["dE"]) <= 1e-9 * max(1.0, abs(d["dE"])), (it, k, out[k]["dE"], d["dE"])
            acc = bool(out[k]["dE"] < 0.5)
            e.commit(acc)
            if acc:
                pos[k][f:l] = trial
    for k, e in enumerate((a, b)):
        assert np.max(np.abs(e.positions() - pos[k])) == 0.0
    for e in engs:
        e.close()


def test_k_sharded_recompute_single_rank_matches_init():
    """pg_sk_compute_slice / pg_sk_set / pg_sk_energy (the k-sharded recompute, world = 1 here):
    slices written into a torch CUDA buffer reproduce the engine's own S(k) and reciprocal energy."""
    import torch
    from plum_b200 import sharded
    r, s, types, params = replay.load_golden("bulk_muvt")
    rng = np.random.default_rng(3)
    box = [40.0, 40.0, 40.0]
    r2, types2, params2 = _params(box, alpha=0.03)
    sysm = _random_system(rng, 6, 20, 60, np.array(box))
    eng = _engine(params2, sysm.n)
    eng.upload(sysm.xyz, sysm.q, types2.ids(sysm.symbol), sysm.mol_first)
    t = eng.init_energy()
    sk0 = eng.sk_download()
    n_k = sk0.shape[0]
    torch.cuda.set_device(0)
    # emulate 3 ranks on one device: every slice is computed separately, then adopted
    parts = []
    e_sum = 0.0
    for rank in range(3):
        f, c = sharded.k_slice(n_k, rank, 3)
        buf = torch.zeros((max(c, 1), 2), dtype=torch.float64, device="cuda")
        eng.sk_compute_slice(f, c, buf.data_ptr())
        parts.append(buf[:c])
        e_sum += eng.sk_energy(f, c)
    full = torch.cat(parts).contiguous()
    assert np.max(np.abs(full.cpu().numpy() - sk0)) <= 1e-12 * max(1.0, np.max(np.abs(sk0)))
    assert abs(e_sum - t["recip"]) <= 1e-12 * max(1.0, abs(t["recip"]))
    sk, e = sharded.sharded_sk_recompute(eng, 0, 1)
    assert abs(e - t["recip"]) <= 1e-12 * max(1.0, abs(t["recip"]))
    assert np.array_equal(eng.sk_download(), sk.cpu().numpy())
    eng.close()


def test_stress_variant_s_full_parity():
    """SURVEY.md §8(d) stress variant S-full — every monomer charged: N = 40 000 beads, all of them charged,
    L = 250, alpha = 0.003 (real_cutoff 60, K = 4138).  Ewald set-up, the full S(k) (also recomputed slice by
    slice as the k-sharded path does), two ion moves, one 100-bead chain move with its inverse, against the
    oracle; the device-side proposal path runs the same chain as the per-move path at this size, too."""
    import os
    import torch
    from plum_b200 import mcgen, sharded, synth
    r, s, types, params = synth.load_full(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    assert s.n == 40000 and int((s.q != 0).sum()) == 40000
    ids = types.ids(s.symbol)
    eng, orc = _engine(params, s.n), _oracle(params)
    a, b = eng.ewald_info(), orc.ewald_info()
    assert (a.real_cutoff, a.repl_cutoff, a.n_k, a.n_k_half) == (b.real_cutoff, b.repl_cutoff, b.n_k, b.n_k_half)
    assert a.n_k == 4138 and a.real_cutoff == 60.0
    eng.upload(s.xyz, s.q, ids, s.mol_first); orc.upload(s.xyz, s.q, ids, s.mol_first)
    t = eng.init_energy()
    sk_o = orc.sk_half()
    sk_g = eng.sk_download()
    assert np.max(np.abs(sk_g - sk_o)) <= 1e-10 * max(1.0, float(np.max(np.abs(sk_o))))
    # k-sharded recompute, 4 emulated ranks: bit-identical rows, energies add up
    torch.cuda.set_device(0)
    e_sum, parts = 0.0, []
    for rank in range(4):
        f, c = sharded.k_slice(a.n_k_half, rank, 4)
        buf = torch.zeros((max(c, 1), 2), dtype=torch.float64, device="cuda")
        eng.sk_compute_slice(f, c, buf.data_ptr())
        parts.append(buf[:c])
        e_sum += eng.sk_energy(f, c)
    assert np.array_equal(torch.cat(parts).cpu().numpy(), sk_g)
    assert abs(e_sum - t["recip"]) <= 1e-12 * max(1.0, abs(t["recip"]))
    rng = np.random.default_rng(11)
    ions = [m for m in range(s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] == 1]
    chains = [m for m in range(s.n_mol) if s.mol_first[m + 1] - s.mol_first[m] > 1]

    def one(mol, trial, accept):
        mv = np.ones(trial.shape[0], dtype=np.uint8)
        dg, do = eng.delta_e(mol, trial, mv), orc.delta_e(mol, trial, mv)
        for k in ("dE", "pair", "ewald", "real", "recip"):
            assert abs(dg[k] - do[k]) <= TOL * max(1.0, abs(do[k])), (mol, k, dg[k], do[k])
        eng.commit(accept); orc.commit(accept)
        return dg["dE"]

    pos = s.xyz
    for i in range(2):
        m = int(rng.choice(ions)); f = int(s.mol_first[m])
        v = rng.normal(size=3)
        one(m, pos[f:f + 1] + 6.0 * v / np.linalg.norm(v), accept=False)
    m = int(rng.choice(chains)); f, l = int(s.mol_first[m]), int(s.mol_first[m + 1])
    dE1 = one(m, pos[f:l] + rng.normal(scale=0.3, size=(l - f, 3)), accept=True)
    dE2 = one(m, pos[f:l].copy(), accept=True)
    assert abs(dE1 + dE2) <= 1e-9 * max(1.0, abs(dE1))
    assert np.array_equal(eng.positions(), pos)
    # device-side proposals at this size: same chain as the per-move path
    eng2 = _engine(params, s.n)
    eng2.upload(s.xyz, s.q, ids, s.mol_first); eng2.init_energy()
    n = 120
    ca = mcgen.NativeChain(eng, r, s, 21, record_moves=n, record_trials=True)
    cb = mcgen.NativeChain(eng2, r, s, 21, record_moves=n)
    ca.run_per_move(n); cb.run_batched(n, 64)
    assert ca.rec_mol.tolist() == cb.rec_mol.tolist() and ca.rec_acc.tolist() == cb.rec_acc.tolist()
    assert np.array_equal(ca.positions(), cb.positions())
    eng.close(); eng2.close()


@pytest.mark.parametrize("world", [2, 4, 7])
def test_fused_k_sharded_recompute_with_emulated_ranks(world):
    """pg_recompute_sk with peers (pg_sk_attach_local: `world` engines of this process on one GPU stand in for the ranks): every
    rank computes its k slice, the reduction epilogue stores it into every peer's S(k) and handshakes with flags.  Every
    rank's S(k) must equal, bit for bit, the S(k) a single engine computes alone (the chunking of the charged list does not
    depend on the slicing), and the reciprocal energy the initialisation's."""
    r, s, types, params = replay.load_golden("synth_cut")
    ids = types.ids(s.symbol)
    engs = []
    for _ in range(world):
        e = _engine(params, s.n)
        e.upload(s.xyz, s.q, ids, s.mol_first)
        engs.append(e)
    t0 = engs[0].init_energy()
    want = engs[0].sk_download()
    for rk, e in enumerate(engs):
        e.sk_attach_local(rk, engs)
    for rep in range(3):                      # flags are sequence numbers: the collective can be repeated
        for e in engs:
            e.recompute_sk_begin()
        for e in engs:
            en, ms = e.recompute_sk_end()
            assert abs(en - t0["recip"]) <= 1e-12 * max(1.0, abs(t0["recip"]))
        for e in engs:
            assert np.array_equal(e.sk_download(), want)
    for e in engs:
        e.sk_detach()
        assert abs(e.recompute_sk() - t0["recip"]) <= 1e-12 * max(1.0, abs(t0["recip"]))
        e.close()


def test_recompute_sk_resets_drift_without_touching_totals():
    """After a stretch of accepted moves the incrementally updated S(k) equals the recomputed one to rounding; pg_recompute_sk
    replaces it and leaves the accumulated totals alone."""
    from plum_b200 import mcgen
    r, s, types, params = replay.load_golden("synth_cut")
    e = _engine(params, s.n)
    e.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    e.init_energy()
    c = mcgen.NativeChain(e, r, s, 9, record_moves=400)
    c.run_chain(400, batch=400, cluster=4)
    before, tot = e.sk_download(), e.totals()
    en = e.recompute_sk()
    after = e.sk_download()
    assert np.abs(after - before).max() <= 1e-10
    assert e.totals() == tot
    assert abs(en - e.recompute_totals()["recip"]) <= 1e-12 * max(1.0, abs(en))
    e.close()
