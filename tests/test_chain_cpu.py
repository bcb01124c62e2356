"""CPU pins of the device-resident chain's host/device-shared pieces (plum_b200/csrc/pg_chain_gen.h):

  * its two-buffer mt19937 reproduces std::mt19937 draw for draw, also when looking ahead across a generation boundary;
  * its draw order of a translational step (Simulation::Run / TranslationalMove, src/simulation/simulation.cc:216-355; the
    generators of src/molecules/molecule.cc:103-312; randSphere src/utilities/misc.cc:95-109) equals, bit for bit,
    plum_b200/host/mc_propose.h — which tests/test_proposals_cpu.py pins to the trial coordinates plum_ref itself wrote —
    including grand-canonical stops, varied bond lengths and the acceptance draw that is skipped when dE >= 1e8;
  * numpy's legacy seeding equals std::mt19937(seed) (what Engine.chain_seed relies on).
"""
import os
import subprocess

import numpy as np
import pytest

import replay

REPO = replay.REPO


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("chain_gen") / "chain_gen_check"
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-ffp-contract=off", "-I", os.path.join(REPO, "include"),
                           "-I", os.path.join(REPO, "plum_b200", "host"), "-I", os.path.join(REPO, "plum_b200", "csrc"),
                           os.path.join(REPO, "tests", "native", "chain_gen_check.cc"), "-o", str(exe)])
    return str(exe)


@pytest.mark.parametrize("seed,steps,vary,gc", [(1, 20000, 0, 0), (2, 20000, 1, 0), (3, 20000, 0, 10), (4, 6000, 1, 7)])
def test_chain_generator_equals_host_generator(checker, seed, steps, vary, gc):
    out = subprocess.run([checker, str(seed), str(steps), str(vary), str(gc)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "\nok" in "\n" + out.stdout, out.stdout


def test_numpy_legacy_seeding_is_std_mt19937():
    from plum_b200 import mcgen
    for seed in (1, 5, 12345):
        key = np.random.RandomState(seed).get_state()
        assert key[2] == 624
        bg = np.random.MT19937()
        bg.state = {"bit_generator": "MT19937", "state": {"key": key[1], "pos": 624}}
        g = mcgen.Generator([1, 1], 0, 1.0, [1.0, 0, 0, 0, 0], 1.0, False, 0, seed)
        assert int(bg.random_raw()) == g.peek_raw()
