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
    out = bench.plum_ref_cut((2, 12, None))
    assert out["N"] == 2 * 100 + 2 * 10          # 2 chains x 100 beads + one counter-ion per charged monomer
    assert out["moves"] == out["ion_moves"] + out["chain_moves"] and 0 < out["moves"] <= 12
    assert out["moves_s"] > 0 and out["init_s"] > 0


@pytest.mark.skipif(not replay.have_plum_ref(), reason="oracle/_ref/plum_ref not built here")
def test_example_timing_of_the_reference_binary():
    """The examples block of the bench line: one reference example through the driver binary, difference of two runs."""
    out = bench.run_example(("confined_muvt", replay.PLUM_REF, 60, None, None))
    assert out["steps"] == 60 and out["steps_per_s"] > 0
    assert out["translational_steps"] + out["gc_steps"] <= 60


def test_algorithmic_work_formulas_match_survey_8d():
    """chain move of S = 2*2 194 950*36 + 2*10*3574*16 + 12*3574 flop; ion move = 2*21 999*36 + 2*3574*16 +
    12*3574 flop."""
    from plum_b200 import synth
    _, sysm, _, _ = synth.load(cache_dir=os.path.join(replay.REPO, "gpurun_out", "cache"))
    chain = next(m for m in range(sysm.n_mol) if sysm.mol_first[m + 1] - sysm.mol_first[m] == 100)
    ion = next(m for m in range(sysm.n_mol) if sysm.mol_first[m + 1] - sysm.mol_first[m] == 1)
    ev, fl, nq = bench.alg_flops_per_move(sysm.n, sysm.mol_first, sysm.q, np.array([chain, ion]))
    assert ev[0] == 100 * 21900 + 4950 and ev[1] == 21999
    assert fl[0] == 2 * 2194950 * 36 + 2 * 10 * 3574 * 16 + 12 * 3574
    assert fl[1] == 2 * 21999 * 36 + 2 * 1 * 3574 * 16 + 12 * 3574
    assert nq[0] == 10 and nq[1] == 1


def test_profile_summary_feeds_the_roofline():
    """bench.py takes the executed-work counters of k_chain from the tracked ncu summary, not from constants."""
    name, prof = bench.load_profile_summary()
    assert name and name.startswith("r") and prof
    for key in ("executed_fp64_flop_per_step", "warp_instructions_per_step", "dram_bytes_per_step", "steps_in_launch"):
        assert prof[key] > 0


def test_tracked_ncu_summary_carries_what_the_roofline_reads():
    """`roofline` of the bench line takes its executed-work counters from the newest profiles/r*_ncu_summary.json: the file
    must be there, carry the fleet capture (the launch shape the bench times) and every per-step figure bench.py reads,
    with plausible magnitudes — so that the fraction cannot silently lose its counters or go stale in shape."""
    name, prof = bench.load_profile_summary()
    assert name is not None and name.startswith("r") and prof["section"] == "k_chain_fleet"
    for key in ("executed_fp64_flop_per_step", "warp_instructions_per_step", "l2_bytes_per_step", "dram_bytes_per_step"):
        assert prof[key] > 0, key
    assert 1e5 < prof["executed_fp64_flop_per_step"] < 1e7 and 1e4 < prof["warp_instructions_per_step"] < 1e7
    assert prof["steps_in_launch"] >= 100 and "device_wide_pct" in prof
    # executed FP64 cannot be below what had to be computed by more than the share of launches that end in an overlap
    assert prof["executed_fp64_flop_per_step"] > 0.5e6
