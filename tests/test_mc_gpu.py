"""GPU parity of the batched Markov-chain path (pg_mc_*: device-side proposals k_propose + k_move + Metropolis
test + commit without the host), through the C ABI:

  * against the REFERENCE: the chain the device runs from PLUM_SEED's random stream reproduces the committed
    plum_ref traces step for step — molecule, move kind, accept/reject bit, dE within 1e-10 — and the trial
    coordinates k_propose built are bit-identical to the ones the reference wrote (X lines);
  * against the per-move path (pg_delta_e / pg_commit with host-built trial coordinates): same seed, same chain,
    bit-identical final coordinates;
  * the stop rule: a step with dE >= 1e8 draws no acceptance variate in the reference (simulation.cc:327-332),
    so the batch ends there and the host generator rewinds (confined_nvt has 22 such steps in 500).
"""
import numpy as np
import pytest

import replay
from plum_b200 import mcgen
from plum_b200.engine import Engine

pytestmark = pytest.mark.gpu
VLE = 1.0e8


def _engine(name):
    r, s, types, params = replay.load_golden(name)
    eng = Engine(params, device=0, capacity_beads=max(s.n, 1))
    eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first)
    eng.init_energy()
    return r, s, eng


def _trace_moves(lines):
    out = []
    for i, ln in enumerate(lines):
        if ln.startswith("T "):
            t = ln.split()
            x = lines[i + 1].split()
            n = int(x[1])
            trial = np.array([[replay.hx(x[3 + 4 * k + a]) for a in range(3)] for k in range(n)])
            out.append((int(t[2]), int(t[3]), replay.hx(t[4]), int(t[5]), trial, [replay.hx(v) for v in t[6:10]]))
        elif ln.startswith("G "):
            break
    return out


@pytest.mark.parametrize("name,seed,batch", [("bulk_nvt", 1, 64), ("bulk_nvt", 2, 1000), ("confined_nvt", 1, 37),
                                             ("confined_nvt", 2, 512), ("synth_spring", 1, 128), ("synth_spring", 2, 16),
                                             ("synth_crank", 1, 64), ("synth_crank", 2, 400)])
def test_batched_chain_reproduces_reference_trace(name, seed, batch):
    r, s, eng = _engine(name)
    ref = _trace_moves(replay.golden_short_trace(name, seed))
    n = len(ref)
    chain = mcgen.NativeChain(eng, r, s, seed, record_moves=n)
    chain.run_batched(n, batch)
    assert [m[1] for m in ref] == chain.rec_mol.tolist()
    assert [m[3] for m in ref] == chain.rec_acc.tolist()
    for i, m in enumerate(ref):
        if m[2] >= VLE:
            assert chain.rec_dE[i] >= VLE, (i, m[2], chain.rec_dE[i])
        else:
            assert replay.rel(chain.rec_dE[i], m[2]) <= 1e-10, (i, m[2], chain.rec_dE[i])
    # the chain's final coordinates == the reference's accepted trial coordinates, bit for bit
    want = s.xyz.copy()
    for kind, mol, dE, acc, trial, _ in ref:
        if acc:
            want[s.mol_first[mol]:s.mol_first[mol + 1]] = trial
    assert np.array_equal(chain.positions(), want)
    assert np.array_equal(eng.positions(), want)
    # running totals after the last step
    tot = eng.totals()
    for k, v in zip(("pair", "ewald", "bond", "ext"), ref[-1][5]):
        assert replay.rel(tot[k], v) <= 1e-9, (k, tot[k], v)


@pytest.mark.parametrize("name,seed", [("confined_nvt", 1), ("synth_spring", 1), ("synth_crank", 1), ("synth_crank", 2)])
def test_device_built_trial_coordinates_are_bit_identical(name, seed):
    """One uploaded batch, read back step by step (pg_mc_trial_xyz) up to the first stop."""
    r, s, eng = _engine(name)
    ref = _trace_moves(replay.golden_short_trace(name, seed))
    g = mcgen.Generator.for_run(r, s.mol_first, seed)
    descs, rows, n_rows = [], [], 0
    for m in ref:
        kind, d, rv = g.next()
        assert kind == m[0] and d.mol == m[1]
        if kind == 2:     # pivot rows of this step start here (a crankshaft keeps its last-bead index in rv_offset)
            d.rv_offset = n_rows
        descs.append(d)
        rows.append(rv)
        n_rows += rv.shape[0]
        if m[2] >= VLE:
            break   # the stream behind this step is shifted; the device must stop here by itself
    extra = 5       # steps behind the stop: must be left untouched
    for _ in range(extra):
        kind, d, rv = g.next()
        if kind < 0:
            break
        if kind == 2:
            d.rv_offset = n_rows
        descs.append(d)
        rows.append(rv)
        n_rows += rv.shape[0]
    eng.mc_upload(descs, np.concatenate(rows) if n_rows else None)
    dE, acc, n_done, ms = eng.mc_run(0, len(descs))
    stops = [i for i, m in enumerate(ref) if m[2] >= VLE]
    want_done = (stops[0] + 1) if stops else len(ref)
    assert n_done == min(want_done, len(descs))
    for i in range(n_done):
        kind, mol, dE_ref, acc_ref, trial, _ = ref[i]
        got = eng.mc_trial_xyz(i, trial.shape[0])
        assert np.array_equal(got, trial), (i, kind, np.abs(got - trial).max())
        assert int(acc[i]) == acc_ref
        if dE_ref < VLE:
            assert replay.rel(dE[i], dE_ref) <= 1e-10


@pytest.mark.parametrize("name", ["bulk_nvt", "confined_nvt", "synth_spring", "synth_crank"])
def test_batched_path_equals_per_move_path(name):
    """Same seed through pg_delta_e/pg_commit (host-built trials) and through pg_mc_*: the same Markov chain."""
    n = 600
    r, s, eng_a = _engine(name)
    _, _, eng_b = _engine(name)
    a = mcgen.NativeChain(eng_a, r, s, 77, record_moves=n, record_trials=True)
    b = mcgen.NativeChain(eng_b, r, s, 77, record_moves=n)
    a.run_per_move(n)
    b.run_batched(n, 100)
    assert a.rec_mol.tolist() == b.rec_mol.tolist()
    assert a.rec_acc.tolist() == b.rec_acc.tolist()
    big = a.rec_dE >= VLE
    assert np.array_equal(big, b.rec_dE >= VLE)
    assert np.all(np.abs(a.rec_dE[~big] - b.rec_dE[~big]) <= 1e-10 * np.maximum(1.0, np.abs(a.rec_dE[~big])))
    assert np.array_equal(a.positions(), b.positions())
    assert np.array_equal(eng_a.positions(), eng_b.positions())
    ta, tb = eng_a.totals(), eng_b.totals()
    for k in ("pair", "ewald", "bond", "ext"):
        assert replay.rel(ta[k], tb[k]) <= 1e-10
    # the two paths can be interleaved on one engine: the per-move path continues the batched chain
    b.run_per_move(50)
    a.run_batched(50, 7)
    assert np.array_equal(a.positions(), b.positions())


def test_full_size_system_batched_equals_per_move():
    """BASELINE.json configs[4] (22 000 beads, 100-bead chains: pivots of 99 sequential steps on the device)."""
    from plum_b200 import synth
    import os
    r, s, types, params = synth.load(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    ids = types.ids(s.symbol)
    engs = []
    for _ in range(2):
        e = Engine(params, device=0, capacity_beads=s.n)
        e.upload(s.xyz, s.q, ids, s.mol_first)
        e.init_energy()
        engs.append(e)
    n = 300
    a = mcgen.NativeChain(engs[0], r, s, 5, record_moves=n, record_trials=True)
    b = mcgen.NativeChain(engs[1], r, s, 5, record_moves=n)
    a.run_per_move(n)
    b.run_batched(n, 128)
    assert a.rec_mol.tolist() == b.rec_mol.tolist()
    assert a.rec_acc.tolist() == b.rec_acc.tolist()
    assert np.all(np.abs(a.rec_dE - b.rec_dE) <= 1e-10 * np.maximum(1.0, np.abs(a.rec_dE)))
    assert np.array_equal(a.positions(), b.positions())
    assert 0 < int(a.rec_acc.sum()) < n
