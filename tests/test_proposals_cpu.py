"""CPU pin of the host half of the batched Markov-chain path (plum_b200/host/mc_propose.h): walking a
std::mt19937 in the reference's draw order and building the trial coordinates with the arithmetic the device
kernel k_propose shares (plum_b200/csrc/pg_propose_math.h) must reproduce, BIT FOR BIT, the trial coordinates
the reference binary itself wrote into the committed traces (X lines of tests/golden/short/*.trace.gz, written
by oracle/_ref/plum_ref with PLUM_SEED / PLUM_TRACE_XYZ), the molecule and move kind of every step, and — via the
pre-drawn acceptance variate — every accept/reject decision."""
import math

import numpy as np
import pytest

import replay
from plum_b200 import mcgen

VLE = 1.0e8


def _walk(name, seed, lines, r, s, max_steps=None):
    g = mcgen.Generator.for_run(r, s.mol_first, seed)
    xyz = s.xyz.copy()
    step_done = 0
    n_moves = 0
    kinds = set()
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("G "):
            # the reference's CBMC consumes the stream from here on: the generator must have stopped right here
            step = int(ln.split()[1])
            while step_done < step - 1:
                kind, _, _ = g.next()
                assert kind == mcgen.KIND_NONE
                step_done += 1
            kind, _, _ = g.next()
            assert kind == mcgen.KIND_GC, (name, seed, step, kind)
            break
        if not ln.startswith("T "):
            i += 1
            continue
        t = ln.split()
        step, move_type, mol, dE, accept = int(t[1]), int(t[2]), int(t[3]), replay.hx(t[4]), int(t[5])
        x = lines[i + 1].split()
        assert x[0] == "X"
        n = int(x[1])
        moved = np.array([int(x[2 + 4 * k]) for k in range(n)])
        trial = np.array([[replay.hx(x[3 + 4 * k + a]) for a in range(3)] for k in range(n)])
        i += 2
        # steps that attempted nothing leave no T line
        while step_done < step - 1:
            kind, _, _ = g.next()
            assert kind == mcgen.KIND_NONE, (name, seed, step_done, kind)
            step_done += 1
        kind, d, rv = g.next()
        step_done += 1
        assert kind == move_type and d.mol == mol, (name, seed, step, kind, move_type, d.mol, mol)
        f, l = int(s.mol_first[mol]), int(s.mol_first[mol + 1])
        assert l - f == n
        mine = mcgen.apply(d, rv, xyz[f:l])
        assert np.array_equal(mine, trial), (name, seed, step, kind, np.abs(mine - trial).max())
        if kind == 3:     # crankshaft: only the beads strictly between the axis beads (molecule.cc:251-253)
            assert moved.tolist() == [int(d.i0 < k < d.rv_offset) for k in range(n)]
        else:
            assert moved.tolist() == ([1] + [0] * (n - 1) if kind == 0 else [1] * n)
        if dE >= VLE:
            assert accept == 0
            g.no_accept_draw()
        else:
            assert accept == int(d.u < math.exp(-r.beta * dE)), (name, seed, step, d.u, dE)
        if accept:
            xyz[f:l] = trial
        kinds.add(kind)
        n_moves += 1
        if max_steps and n_moves >= max_steps:
            break
    return n_moves, kinds


@pytest.mark.parametrize("name,seed", [("bulk_nvt", 1), ("bulk_nvt", 2), ("confined_nvt", 1), ("confined_nvt", 2),
                                       ("bulk_muvt", 1), ("confined_muvt", 2), ("synth_spring", 1), ("synth_spring", 2),
                                       ("synth_crank", 1), ("synth_crank", 2)])
def test_generator_reproduces_reference_trial_coordinates(name, seed):
    r, s, _, _ = replay.load_golden(name)
    lines = replay.golden_short_trace(name, seed)
    n_moves, kinds = _walk(name, seed, lines, r, s)
    if not r.use_gc:
        assert n_moves >= 200 and {0, 2} <= kinds, (n_moves, kinds)
    if name == "synth_crank":    # Molecule::Crankshaft (molecule.cc:239-265), Eigen's rotation order
        assert {0, 1, 2, 3, 4} <= kinds
    if name == "synth_spring":   # spring bonds: Pivot / RandomReptation draw a bond length per step
        assert r.use_bond and {0, 1, 2, 4} <= kinds
