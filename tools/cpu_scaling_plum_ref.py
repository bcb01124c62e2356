#!/usr/bin/env python3
"""profiles/<tag>_cpu_scaling_plum_ref.jsonl: oracle/_ref/plum_ref ITSELF (one core each) on the four cuts of the benchmark system
S that BASELINE.md §3 names — 12 / 25 / 50 / 100 chains of 100 beads + counter-ions = N 1320 / 2750 / 5500 / 11000, same box,
alpha, K = 3574 and move mix — and the per-core fit linear in n_moved x N with its extrapolation to N = 22000.  The N = 11000 cut
needs ~29 GB of std::map nodes and minutes of energy initialisation, which is why bench.py's reference arm stops at 5500.
CPU only:  python tools/cpu_scaling_plum_ref.py profiles/r02_cpu_scaling_plum_ref.jsonl [--max-chains 100]"""
import argparse
import json
import multiprocessing as mp
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--max-chains", type=int, default=100)
    a = ap.parse_args()
    jobs = [(c, s, i) for i, (c, s) in enumerate([(12, 60), (25, 60), (50, 40), (100, 24)]) if c <= a.max_chains]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        cuts = pool.map(bench.plum_ref_cut, jobs)
    num = sum(c["moves_s"] * (c["ion_moves"] + 100.0 * c["chain_moves"]) * c["N"] for c in cuts)
    den = sum(((c["ion_moves"] + 100.0 * c["chain_moves"]) * c["N"]) ** 2 for c in cuts)
    a_fit = num / den
    t_move = a_fit * 22000 * (0.5 * 1 + 0.5 * 100)
    with open(a.out, "w") as f:
        for c in cuts:
            c["s_per_move"] = c["moves_s"] / max(c["moves"], 1)
            f.write(json.dumps(c) + "\n")
        f.write(json.dumps({"fit_s_per_moved_bead_partner": a_fit, "extrapolated_s_per_move_at_N22000": t_move,
                            "extrapolated_moves_per_s_per_core_at_N22000": 1.0 / t_move,
                            "note": "EXTRAPOLATED from the cuts above (plum_ref cannot hold N = 22000: 70-115 GB of map nodes); one core per "
                                    "run; move mix 0.5 ion / 0.5 chain; measured in the build container's CPU, not on the GPU box"}) + "\n")
    print(open(a.out).read())


if __name__ == "__main__":
    main()
