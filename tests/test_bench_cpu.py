"""Host-side pieces of bench.py that run without a GPU: the reference arm's real-binary leg and the
algorithmic work formulas of SURVEY.md §8(d)."""
import os
import sys

import numpy as np
import pytest

import replay

sys.path.insert(0, replay.REPO)
import bench  # noqa: E402


@pytest.mark.skipif(not replay.have_plum_ref(), reason="oracle/_ref/plum_ref not built here")
def test_reference_arm_runs_the_real_binary_on_a_cut_of_S():
    out = bench.plum_ref_cut(n_chains=2, steps=12)
    assert out["kind"] == "reference" and out["cores"] == 1
    assert out["N"] == 2 * 100 + 2 * 10          # 2 chains x 100 beads + one counter-ion per charged monomer
    assert out["moves"] == out["ion_moves"] + out["chain_moves"] and 0 < out["moves"] <= 12
    assert out["measured_moves_per_s_at_this_N"] > 0
    # extrapolation is linear in N: the full system must come out slower than the cut
    assert 0 < out["extrapolated_moves_per_s_at_N22000"] < out["measured_moves_per_s_at_this_N"]


def test_algorithmic_work_formulas_match_survey_8d():
    """chain move of S = 2*2 194 950*36 + 2*10*3574*16 + 12*3574 flop and 40 N + 40 K + 48*100 bytes;
    ion move = 2*21 999*36 + 2*3574*16 + 12*3574 flop."""
    from plum_b200 import synth
    _, sysm, _, _ = synth.load(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    chain = next(m for m in range(sysm.n_mol) if sysm.mol_first[m + 1] - sysm.mol_first[m] == 100)
    ion = next(m for m in range(sysm.n_mol) if sysm.mol_first[m + 1] - sysm.mol_first[m] == 1)
    ev, fl, by = bench.alg_flops_per_move(sysm, np.array([chain, ion]))
    assert ev[0] == 100 * 21900 + 4950 and ev[1] == 21999
    assert fl[0] == 2 * 2194950 * 36 + 2 * 10 * 3574 * 16 + 12 * 3574
    assert fl[1] == 2 * 21999 * 36 + 2 * 1 * 3574 * 16 + 12 * 3574
    assert by[0] == 40 * 22000 + 40 * 3574 + 48 * 100 and by[1] == 40 * 22000 + 40 * 3574 + 48
