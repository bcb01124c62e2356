"""GPU parity of the device-resident Markov chain (k_chain through pg_chain_*: random stream, proposals, energy change,
Metropolis test and commit in one kernel; cell-grid LJ, compact charged list, factorised reciprocal phases), C ABI:

  * against the REFERENCE: from PLUM_SEED's std::mt19937 state the chain reproduces the traces plum_ref wrote — the
    112-bead spring system (varied bond lengths) and the 1320-bead cut of the benchmark system S (several cells, partner
    chunks and k slices per CTA; 2000 steps x 2 seeds): molecule, move kind and accept bit of EVERY step identical,
    dE within 1e-10, the trial coordinates the device built bit-identical to the ones the reference wrote, final
    coordinates bit-identical, running totals within 1e-9 — for every cluster size;
  * against the per-move path (pg_delta_e / pg_commit with host-built trials) on the full 22 000-bead system;
  * the resident structures (cells, charged list) stay consistent with the coordinates (pg_chain_check), S(k) stays the
    structure factor of the coordinates, and the paths can be interleaved on one engine.
"""
import os

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


def _configure(eng, r, cluster, keep_trials=False, pivot_mode=0):
    bl, vary = mcgen.bond_settings(r)
    eng.chain_configure(r.phantom, r.move_size, r.move_prob, bl, vary_bond=vary, gc_freq=r.gc_freq if r.use_gc else 0,
                        cluster=cluster, keep_trials=keep_trials, pivot_mode=pivot_mode)


def _errors(got, ref):
    """(max plain relative error, max error relative to max(1, |ref|)) over the steps with a finite dE."""
    fin = ref < VLE
    d = np.abs(got[fin] - ref[fin])
    plain = float(np.max(d / np.maximum(np.abs(ref[fin]), 1e-300))) if fin.any() else 0.0
    scaled = float(np.max(d / np.maximum(np.abs(ref[fin]), 1.0))) if fin.any() else 0.0
    return plain, scaled


def _trace_moves(lines):
    out = []
    for i, ln in enumerate(lines):
        if ln.startswith("T "):
            t = ln.split()
            x = lines[i + 1].split()
            n = int(x[1])
            trial = np.array([[replay.hx(x[3 + 4 * k + a]) for a in range(3)] for k in range(n)])
            out.append((int(t[2]), int(t[3]), replay.hx(t[4]), int(t[5]), trial, [replay.hx(v) for v in t[6:10]]))
    return out


@pytest.mark.parametrize("seed,cluster,batch", [(1, 1, 400), (2, 1, 37), (1, 2, 400), (2, 4, 128), (1, 8, 400), (2, 16, 400)])
def test_chain_reproduces_reference_trace_spring(seed, cluster, batch):
    r, s, eng = _engine("synth_spring")
    ref = _trace_moves(replay.golden_short_trace("synth_spring", seed))
    n = len(ref)
    _configure(eng, r, cluster, keep_trials=True)
    eng.chain_seed(seed)
    want = s.xyz.copy()
    done = 0
    while done < n:
        b = min(batch, n - done)
        rec, stop, ms = eng.chain_run(b)
        assert len(rec) == b and stop == 0
        for i in range(b):
            kind, mol, dE_ref, acc_ref, trial, _ = ref[done + i]
            assert (int(rec["kind"][i]), int(rec["mol"][i]), int(rec["accept"][i])) == (kind, mol, acc_ref), (done + i, rec[i], ref[done + i][:4])
            got = eng.chain_trial_xyz(i, trial.shape[0])
            assert np.array_equal(got, trial), (done + i, kind, np.abs(got - trial).max())
            if dE_ref >= VLE:
                assert rec["dE"][i] >= VLE
            else:
                assert replay.rel(rec["dE"][i], dE_ref) <= 1e-10, (done + i, kind, rec["dE"][i], dE_ref)
            if acc_ref:
                want[s.mol_first[mol]:s.mol_first[mol + 1]] = trial
        done += b
    assert np.array_equal(eng.positions(), want)
    eng.chain_check()
    tot = eng.totals()
    for k, v in zip(("pair", "ewald", "bond", "ext"), ref[-1][5]):
        assert replay.rel(tot[k], v) <= 1e-9, (k, tot[k], v)
    # S(k) carried incrementally == S(k) of the final coordinates
    fresh = eng.recompute_totals()
    assert replay.rel(tot["ewald"], fresh["ewald"]) <= 1e-9
    eng.close()


@pytest.mark.parametrize("seed,cluster", [(1, 1), (2, 1), (1, 8), (2, 16)])
def test_chain_with_crankshaft_moves_reproduces_reference_trace(seed, cluster):
    """Molecule::Crankshaft (molecule.cc:239-265) inside the device-resident chain: the angle's sine and cosine come
    from the device's sincos (<= 2 ulp from glibc's, which the reference used), so a crankshaft's trial coordinates
    are compared at 1e-13 absolute instead of bit for bit, and once one has been accepted that molecule's later
    coordinates inherit the difference; molecule, move kind and accept bit of all 400 steps must still be the
    reference's, dE within 1e-10."""
    r, s, eng = _engine("synth_crank")
    ref = _trace_moves(replay.golden_short_trace("synth_crank", seed))
    n = len(ref)
    _configure(eng, r, cluster, keep_trials=True)
    eng.chain_seed(seed)
    want = s.xyz.copy()
    rec, stop, ms = eng.chain_run(n)
    assert len(rec) == n and stop == 0
    n_crank, worst = 0, 0.0
    touched = set()      # molecules whose coordinates have gone through an accepted crankshaft
    for i in range(n):
        kind, mol, dE_ref, acc_ref, trial, _ = ref[i]
        assert (int(rec["kind"][i]), int(rec["mol"][i]), int(rec["accept"][i])) == (kind, mol, acc_ref), (i, rec[i], ref[i][:4])
        got = eng.chain_trial_xyz(i, trial.shape[0])
        if kind == 3 or mol in touched:
            worst = max(worst, float(np.abs(got - trial).max()))
            assert np.abs(got - trial).max() <= 1e-11, (i, kind, np.abs(got - trial).max())
        else:
            assert np.array_equal(got, trial), (i, kind, np.abs(got - trial).max())
        n_crank += kind == 3
        if dE_ref >= VLE:
            assert rec["dE"][i] >= VLE
        else:
            assert replay.rel(rec["dE"][i], dE_ref) <= 1e-10, (i, kind, rec["dE"][i], dE_ref)
        if acc_ref:
            want[s.mol_first[mol]:s.mol_first[mol + 1]] = trial
            if kind == 3:
                touched.add(mol)
    assert n_crank >= 60
    assert np.abs(eng.positions() - want).max() <= 1e-11
    print(f"synth_crank seed {seed} cluster {cluster}: {n_crank} crankshaft steps, max |trial - ref| on touched molecules {worst:.2e}")
    eng.chain_check()
    tot = eng.totals()
    for k, v in zip(("pair", "ewald", "bond", "ext"), ref[-1][5]):
        assert replay.rel(tot[k], v) <= 1e-9, (k, tot[k], v)
    eng.close()


def _cut_fixture(seed):
    z = np.load(os.path.join(replay.GOLDEN, "long", f"synth_cut_seed{seed}.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture
def fleet_kernel(monkeypatch):
    """Single-CTA chains through the 448-thread, two-CTAs-per-SM build of k_chain (what fleets of more chains than SMs run)."""
    monkeypatch.setenv("PLUM_B200_CHAIN_THREADS", "448")


@pytest.mark.parametrize("seed,pivot_mode", [(1, 0), (2, 0), (2, 1), (1, 2), (2, 2)])
def test_fleet_build_of_the_chain_kernel_reproduces_reference_on_the_cut(fleet_kernel, seed, pivot_mode):
    test_chain_reproduces_reference_on_the_1320_bead_cut(seed, 1, pivot_mode)


@pytest.mark.parametrize("seed", [1, 2])
def test_fleet_build_of_the_chain_kernel_reproduces_spring_and_crankshaft_traces(fleet_kernel, seed):
    test_chain_reproduces_reference_trace_spring(seed, 1, 37 if seed == 2 else 400)
    test_chain_with_crankshaft_moves_reproduces_reference_trace(seed, 1)


@pytest.mark.parametrize("seed,cluster,pivot_mode", [(1, 1, 0), (2, 8, 0), (1, 4, 0), (2, 1, 1), (1, 8, 1),
                                                     (1, 1, 2), (2, 1, 2), (1, 16, 2), (2, 8, 2), (1, 2, 2)])
def test_chain_reproduces_reference_on_the_1320_bead_cut(seed, cluster, pivot_mode):
    """plum_ref itself, 2000 steps on 12 x 100-bead chains + 120 ions in S's box (same alpha, cutoffs and K = 3574).
    pivot_mode 0: the reference's operation order, trial coordinates bit-identical; pivot_mode 1: pivot arms as prefix
    sums, trial coordinates equal to 1e-11, everything else held to the same bars; pivot_mode 2: the energies from the
    prefix-sum arms while two warps build the exact ones — trial (= committed) coordinates bit-identical again."""
    r, s, eng = _engine("synth_cut")
    fx = _cut_fixture(seed)
    n = len(fx["kind"])
    _configure(eng, r, cluster, keep_trials=True, pivot_mode=pivot_mode)
    eng.chain_seed(seed)
    n_tr = len(fx["trial_off"]) - 1
    rec, stop, ms = eng.chain_run(n_tr)
    worst = 0.0
    for i in range(n_tr):
        trial = fx["trial_xyz"][fx["trial_off"][i]:fx["trial_off"][i + 1]]
        got = eng.chain_trial_xyz(i, trial.shape[0])
        if pivot_mode != 1:
            assert np.array_equal(got, trial), (i, int(fx["kind"][i]), np.abs(got - trial).max())
        else:
            worst = max(worst, float(np.abs(got - trial).max()))
    if pivot_mode == 1:
        print(f"pivot_mode 1: max |trial - reference trial| over {n_tr} steps = {worst:.3e}")
        assert worst <= 1e-11
    rec2, stop, ms = eng.chain_run(n - n_tr)
    rec = np.concatenate([rec, rec2])
    assert rec["kind"].tolist() == fx["kind"].tolist()
    assert rec["mol"].tolist() == fx["mol"].tolist()
    assert rec["accept"].tolist() == fx["accept"].tolist()
    big = fx["dE"] >= VLE
    assert np.array_equal(rec["dE"] >= VLE, big)
    plain, scaled = _errors(rec["dE"], fx["dE"])
    print(f"synth_cut seed {seed} cluster {cluster} pivot_mode {pivot_mode}: {n} steps, {int(big.sum())} overlap steps, accept ratio "
          f"{fx['accept'].mean():.3f}; max |dE - ref| / |ref| = {plain:.3e}, / max(1, |ref|) = {scaled:.3e}")
    assert scaled <= 1e-10
    eng.chain_check()
    tot = eng.totals()
    for k, v in zip(("pair", "ewald", "bond", "ext"), fx["tot"][-1]):
        assert replay.rel(tot[k], v) <= 1e-9, (k, tot[k], v)
    fresh = eng.recompute_totals()
    for k in ("pair", "ewald"):
        assert replay.rel(tot[k], fresh[k]) <= 1e-9, (k, tot[k], fresh[k])
    eng.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_per_move_kernel_reproduces_reference_on_the_1320_bead_cut(seed):
    """The same fixture through pg_delta_e / pg_commit (k_move<true>, host-built trials): the benchmarked per-move kernel
    pinned to the real reference at a multi-tile size."""
    r, s, eng = _engine("synth_cut")
    fx = _cut_fixture(seed)
    n = len(fx["kind"])
    a = mcgen.NativeChain(eng, r, s, seed, record_moves=n)
    a.run_per_move(n)
    assert a.rec_mol.tolist() == fx["mol"].tolist()
    assert a.rec_acc.tolist() == fx["accept"].tolist()
    big = fx["dE"] >= VLE
    assert np.array_equal(a.rec_dE >= VLE, big)
    plain, scaled = _errors(a.rec_dE, fx["dE"])
    print(f"synth_cut seed {seed} per-move: max |dE - ref| / |ref| = {plain:.3e}, / max(1, |ref|) = {scaled:.3e}")
    assert scaled <= 1e-10
    tot = eng.totals()
    for k, v in zip(("pair", "ewald", "bond", "ext"), fx["tot"][-1]):
        assert replay.rel(tot[k], v) <= 1e-9
    eng.close()


@pytest.mark.parametrize("name,cluster", [("synth_spring", 1), ("synth_cut", 2), ("synth_cut", 8)])
def test_chain_equals_per_move_path_and_interleaves(name, cluster):
    n = 500
    r, s, eng_a = _engine(name)
    _, _, eng_b = _engine(name)
    a = mcgen.NativeChain(eng_a, r, s, 77, record_moves=n)
    b = mcgen.NativeChain(eng_b, r, s, 77, record_moves=n)
    a.run_per_move(n)
    b.run_chain(n, batch=128, cluster=cluster)
    assert a.rec_mol.tolist() == b.rec_mol.tolist()
    assert a.rec_acc.tolist() == b.rec_acc.tolist()
    big = a.rec_dE >= VLE
    assert np.array_equal(big, b.rec_dE >= VLE)
    assert _errors(b.rec_dE, a.rec_dE)[1] <= 1e-10
    assert np.array_equal(a.positions(), b.positions())
    assert np.array_equal(eng_a.positions(), eng_b.positions())
    eng_b.chain_check()
    ta, tb = eng_a.totals(), eng_b.totals()
    for k in ("pair", "ewald", "bond", "ext"):
        assert replay.rel(ta[k], tb[k]) <= 1e-10
    # the per-move path continues the chain's state and the chain continues the per-move path's (structures rebuilt)
    b.run_per_move(60)
    a.run_chain(60, batch=7, cluster=cluster)
    assert a.rec_mol[:60].tolist() == b.rec_mol[:60].tolist()
    assert a.rec_acc[:60].tolist() == b.rec_acc[:60].tolist()
    assert np.array_equal(a.positions(), b.positions())
    eng_a.chain_check()
    eng_a.close(); eng_b.close()


@pytest.mark.parametrize("cluster", [1, 8])
def test_full_size_system_chain_equals_per_move(cluster):
    """BASELINE.json configs[4]: 22 000 beads, 100-bead chains, 4000 charged beads, 1787 half-space k vectors."""
    from plum_b200 import synth
    r, s, types, params = synth.load(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    ids = types.ids(s.symbol)
    engs = []
    for _ in range(2):
        e = Engine(params, device=0, capacity_beads=s.n)
        e.upload(s.xyz, s.q, ids, s.mol_first)
        e.init_energy()
        engs.append(e)
    n = 400
    a = mcgen.NativeChain(engs[0], r, s, 5, record_moves=n)
    b = mcgen.NativeChain(engs[1], r, s, 5, record_moves=n)
    a.run_per_move(n)
    b.run_chain(n, batch=200, cluster=cluster)
    assert a.rec_mol.tolist() == b.rec_mol.tolist()
    assert a.rec_acc.tolist() == b.rec_acc.tolist()
    assert np.array_equal(a.rec_dE >= VLE, b.rec_dE >= VLE)
    plain, scaled = _errors(b.rec_dE, a.rec_dE)
    print(f"S cluster {cluster}: chain vs per-move max |ddE| / |dE| = {plain:.3e}, / max(1, |dE|) = {scaled:.3e}; "
          f"accepted {int(a.rec_acc.sum())} of {n}")
    assert scaled <= 1e-10
    assert np.array_equal(a.positions(), b.positions())
    engs[1].chain_check()
    ta, tb = engs[0].totals(), engs[1].totals()
    for k in ("pair", "ewald"):
        assert replay.rel(ta[k], tb[k]) <= 1e-10
    fresh = engs[1].recompute_totals()
    for k in ("pair", "ewald"):
        assert replay.rel(tb[k], fresh[k]) <= 1e-9
    assert 0 < int(a.rec_acc.sum()) < n
    for e in engs:
        e.close()


def test_full_size_chain_moves_against_the_oracle():
    """The FAST path at N = 22 000 against the CPU oracle directly (not only against the per-move kernel): 400 steps of
    the device-resident chain with the trial coordinates kept; every ion move and the first 24 chain moves (>= 200
    moves in all) are re-evaluated by the oracle's pairwise restatement of ForceField::EnergyDifference on the
    configuration the chain had in front of that step.  Prints plain-relative and max(1, |dE|)-scaled errors."""
    from oracle.oracle_py import Oracle
    from plum_b200 import synth
    r, s, types, params = synth.load(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    ids = types.ids(s.symbol)
    eng = Engine(params, device=0, capacity_beads=s.n)
    eng.upload(s.xyz, s.q, ids, s.mol_first)
    eng.init_energy()
    _configure(eng, r, 8, keep_trials=True)
    eng.chain_seed(21)
    n = 400
    rec, stop, ms = eng.chain_run(n)
    assert len(rec) == n and stop == 0
    orc = Oracle(params)
    pos = s.xyz.copy()
    got, ref, n_chain_checked, kinds = [], [], 0, []
    for i in range(n):
        mol, kind, acc = int(rec["mol"][i]), int(rec["kind"][i]), int(rec["accept"][i])
        if mol < 0:
            continue
        f, l = int(s.mol_first[mol]), int(s.mol_first[mol + 1])
        trial = eng.chain_trial_xyz(i, l - f)
        check = (kind == 0) or n_chain_checked < 24
        if check:
            orc.upload(pos, s.q, ids, s.mol_first)
            mv = np.ones(l - f, dtype=np.uint8)
            do = orc.delta_e(mol, trial, mv)
            orc.commit(False)
            got.append(float(rec["dE"][i])); ref.append(do["dE"]); kinds.append(kind)
            n_chain_checked += kind != 0
        if acc:
            pos[f:l] = trial
    got, ref = np.array(got), np.array(ref)
    assert len(got) >= 200 and n_chain_checked == 24
    assert np.array_equal(got >= VLE, ref >= VLE)
    plain, scaled = _errors(got, ref)
    print(f"S chain vs oracle: {len(got)} moves ({n_chain_checked} chain moves) max |ddE| / |dE| = {plain:.3e}, / max(1, |dE|) = {scaled:.3e}")
    assert scaled <= 1e-10
    assert np.array_equal(eng.positions(), pos)
    eng.close()


def test_many_replicas_in_one_launch_walk_their_own_chains():
    """pg_chain_run_multi: 6 replicas with different seeds in one launch == each of them run alone."""
    r, s, types, params = replay.load_golden("synth_cut")
    ids = types.ids(s.symbol)
    seeds = [3, 4, 5, 6, 7, 8]
    n = 150

    def make():
        e = Engine(params, device=0, capacity_beads=s.n)
        e.upload(s.xyz, s.q, ids, s.mol_first)
        e.init_energy()
        return e
    alone = []
    for sd in seeds[:3]:
        e = make()
        c = mcgen.NativeChain(e, r, s, sd, record_moves=n)
        c.run_chain(n, batch=n, cluster=2)
        alone.append((c.rec_mol.copy(), c.rec_acc.copy(), c.rec_dE.copy(), c.positions()))
        e.close()
    engs = [make() for _ in seeds]
    chains = [mcgen.NativeChain(e, r, s, sd, record_moves=n) for e, sd in zip(engs, seeds)]
    mcgen.run_chain_multi(chains, n, batch=50, cluster=2)
    for i in range(3):
        assert chains[i].rec_mol.tolist() == alone[i][0].tolist()
        assert chains[i].rec_acc.tolist() == alone[i][1].tolist()
        assert np.array_equal(chains[i].rec_dE, alone[i][2])          # same cluster size: bit-identical sums
        assert np.array_equal(chains[i].positions(), alone[i][3])
    assert len({tuple(c.rec_mol.tolist()) for c in chains}) == len(seeds)
    for e in engs:
        e.chain_check()
        e.close()
