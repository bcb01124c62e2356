"""Pins the CPU oracle (oracle/plum_oracle.c) against the REAL reference.

The committed traces under tests/golden/short were written by the reference
itself (oracle/_ref/plum_ref = /root/reference/src compiled with seed + trace
hooks, tests/golden/make_golden.py).  Replaying them through the oracle must
reproduce every per-move dE, every CBMC BeadsEnergy value and the four running
energy totals.  Tolerance: 1e-12 relative (floor 1.0) for the pairwise
reciprocal form, which follows the reference's loop order; 1e-11 for the S(k)
form the CUDA path uses (truncated kPi makes the two forms differ by O(1e-13)).
"""
import numpy as np
import pytest

import replay
from oracle.oracle_py import Oracle

EXAMPLES = ["bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"]


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_replays_reference_trace_pairwise(name, seed):
    r, s, types, params = replay.load_golden(name)
    o = Oracle(params, repl_mode=0)
    rep = replay.replay(o, r, s, types, replay.golden_short_trace(name, seed))
    assert rep.n_moves > 100
    assert rep.sentinel_mismatch == 0, rep.worst
    assert rep.max_rel_dE < 1e-12, rep.worst
    assert rep.max_rel_tot < 1e-11, rep.worst
    assert rep.max_rel_beads < 1e-12, rep.worst
    if "muvt" in name:
        assert rep.n_gc > 10 and rep.n_beads_energy > 1000


@pytest.mark.parametrize("name", ["bulk_nvt", "confined_nvt"])
def test_oracle_structure_factor_form_matches_reference(name):
    r, s, types, params = replay.load_golden(name)
    o = Oracle(params, repl_mode=1)
    rep = replay.replay(o, r, s, types, replay.golden_short_trace(name, 1), max_steps=150)
    assert rep.sentinel_mismatch == 0, rep.worst
    assert rep.max_rel_dE < 1e-11, rep.worst
    assert rep.max_rel_tot < 1e-10, rep.worst


def test_oracle_wall_force_matches_reference_pressure_columns():
    """CalcPressureForceLJELSlit (pressure.cc:404-484): replay the reference's own 3000-step confined_nvt
    trace and sample the oracle's wall force where the reference samples (every 10 * sampling_frequency
    steps after equilibration); the three electrostatic columns of the reference's output_stat.dat must
    come out to its printed 6 significant digits.  The LJ columns are NOT compared: the reference reads an
    uninitialised variable there (potential_truncated_lj.cc:94-100, ~3.5e7 in its own output)."""
    r, s, types, params = replay.load_golden("confined_nvt")
    lines, ref_cols = replay.golden_pressure_fixture()
    o = Oracle(params, repl_mode=1)
    avg = replay.WallForceAverager(s.box)
    checked = []

    def on_step(step):
        if step > r.steps_eq and step % (r.sample_freq * 10) == 0:
            avg.add(o.wall_force(r.phantom))
            got, ref = avg.p_tensor(), ref_cols[step]
            for k in (3, 4, 5):
                assert float(f"{got[k]:.6g}") == ref[k], (step, k, got[k], ref[k])   # ostream default precision
            checked.append(step)

    replay.replay(o, r, s, types, lines, check_totals_every=0, on_step=on_step)
    assert checked == [1000, 2000, 3000]
    # plate LJ force: the site-site LJ sums vanish (wall sites have epsilon 0), what is left is BeadForceOnWall
    assert avg.cum[2] == 0.0 and abs(avg.cum[0]) > 0.0


@pytest.mark.parametrize("name", ["bulk_nvt", "synth_spring"])
def test_oracle_vol_scaling_matches_reference_accumulators(name):
    """CalcPressureVolScalingHSELSlit (pressure.cc:187-387): replay the reference's own trace and sample the oracle
    where the reference samples (steps 200 and 300 of the fixture); the per-species accumulators p_tensor_el /
    p_tensor_hs, the bond and dipole sums and both pressure estimates of the reference's V lines must come out.
    Every entry is a sum of (stretched - current) pair energies, i.e. differences ~1e-7 of the energies themselves:
    the tolerance is 1e-10 of the energy being differenced (the total Ewald / pair energy), which is what a
    relative 1e-10 on the two terms allows; the plain relative error is asserted at 1e-6."""
    r, s, types, params = replay.load_golden(name)
    lines, ref = replay.golden_vol_pressure_fixture(name)
    o = Oracle(params)
    avg = replay.VolScalingAverager(s.box, r.beta)
    checked = []

    def on_step(step):
        if step in ref:
            avg.add(o.vol_scaling_sample(r.phantom))
            tot = o.totals()
            scale = max(1.0, abs(tot["ewald"]), abs(tot["pair"]), abs(tot["bond"]))
            g = ref[step]
            assert avg.vp_z == g["n"]
            for got, want in ((avg.el, g["el"]), (avg.hs, g["hs"]), (avg.p[4:6], g["p"][4:6])):
                assert np.max(np.abs(got - want)) <= 1e-10 * scale, (step, got, want)
                nz = np.abs(want) > 0
                assert np.all(np.abs(got[nz] - want[nz]) <= 1e-6 * np.abs(want[nz])), (step, got, want)
                assert not got[~nz].any()
            assert abs(avg.p[2] - g["p"][2]) <= 1e-14 * abs(g["p"][2])
            assert abs(avg.p[3] - g["p"][3]) <= 1e-12 * abs(g["p"][3])
            for k in (0, 1):   # the two pressure estimates: dU / (A dz) amplifies the error of dU by 1 / (A dz)
                assert abs(avg.p[k] - g["p"][k]) <= 1e-6 * abs(g["p"][k]), (step, k, avg.p[k], g["p"][k])
            checked.append(step)

    replay.replay(o, r, s, types, lines, check_totals_every=0, beads_energy=False, on_step=on_step)
    assert checked == [200, 300]
    assert np.count_nonzero(avg.el) >= 2 and np.count_nonzero(avg.hs) >= 1
    if name == "synth_spring":
        assert avg.p[5] != 0.0


def test_oracle_vol_scaling_is_linear_in_the_stretch():
    """Size-independent property of the volume-perturbation sample: for a stretch this small the energy change is the
    first-order response, so doubling dz doubles every table entry (to the relative size of the stretch itself), and the
    bookkeeping fields do not depend on dz."""
    r, s, types, params = replay.load_golden("bulk_nvt")
    o = Oracle(params)
    o.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    o.init_energy()
    a, b = o.vol_scaling_sample(r.phantom, 1e-5), o.vol_scaling_sample(r.phantom, 2e-5)
    assert a["n_free"] == b["n_free"] == s.n_mol - r.phantom
    for k in ("el", "hs"):
        nz = np.abs(a[k]) > 0
        assert np.array_equal(nz, np.abs(b[k]) > 0)     # (no WCA pair is inside its cutoff in the start configuration: hs may be empty)
        assert k == "hs" or nz.any()
        assert np.all(np.abs(b[k][nz] / a[k][nz] - 2.0) <= 1e-3), (k, a[k], b[k])
    assert abs(b["dU"] / a["dU"] - 2.0) <= 1e-3
    assert abs(a["dU"] - (a["el"].sum() + a["hs"].sum() + a["bond"] + a["dipole"])) <= 1e-12 * max(1.0, abs(a["dU"]))


def test_ewald_setup_matches_reference_logs():
    """Cutoffs / k tables as echoed by the reference."""
    expect = {
        "bulk_nvt": (100.0, [2, 2, 2], [2, 2, 2], 26),
        "confined_nvt": (100.0, [3, 3, 1], [2, 2, 7], 40),
        "bulk_muvt": (100.0, [1, 1, 1], [3, 3, 3], 92),
    }
    for name, (rc, real_cell, repl_cell, nk) in expect.items():
        _, _, _, params = replay.load_golden(name)
        info = Oracle(params).ewald_info()
        assert info.real_cutoff == rc
        assert list(info.real_cell) == real_cell
        assert list(info.repl_cell) == repl_cell
        assert info.n_k == nk and info.n_k_half * 2 == nk


@pytest.mark.skipif(not replay.have_plum_ref(), reason="oracle/_ref/plum_ref not built here")
def test_fresh_reference_run_agrees_with_oracle():
    """A trace generated now by the reference binary (different seed from the fixtures)."""
    import os
    d = "/root/reference/examples/confined_nvt"
    if not os.path.isdir(d):
        d = replay.golden_example_dir("confined_nvt")
    r, s, types, params = replay.load_golden("confined_nvt")
    lines = replay.run_plum_ref(d, 200, 7)
    rep = replay.replay(Oracle(params), r, s, types, lines)
    assert rep.n_moves == 200 and rep.sentinel_mismatch == 0
    assert rep.max_rel_dE < 1e-12, rep.worst


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("seed", [1, 2])
def test_long_fixture_is_the_same_reference_trajectory_as_the_short_trace(name, seed):
    """The 10^5-step fixtures (tests/golden/long, accept bits / move kinds / sampled dE) and the replayable short
    traces were written by separate plum_ref runs with the same PLUM_SEED: the reference is deterministic for a
    fixed seed, so the long fixture must start with exactly the short trace — same move kind,
    molecule, accept bit and hex-float dE at every step.  Guards the fixtures themselves (a stale or mislabelled
    file would otherwise only show up on the GPU box)."""
    import numpy as np
    gold = replay.golden_long(name, seed)
    acc = np.unpackbits(gold["accept"])
    dE_at = dict(zip(gold["dE_step"].tolist(), gold["dE"].tolist()))
    n = 0
    for ln in replay.golden_short_trace(name, seed):
        f = ln.split()
        if not f or f[0] not in ("T", "G"):
            continue
        step = int(f[1])
        i = step - 1
        if f[0] == "T":
            assert gold["kind"][i] == 0 and gold["move_type"][i] == int(f[2]) and gold["mol"][i] == int(f[3]), ln
            assert acc[i] == int(f[5]), ln
        else:
            assert gold["kind"][i] == (1 if f[2] == "I" else 2) and gold["mol"][i] == int(f[3]), ln
        if step in dE_at:
            assert dE_at[step] == replay.hx(f[4]), ln
        n += 1
    assert n >= 200
