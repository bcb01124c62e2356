#!/usr/bin/env python3
"""Timing probe of the device-resident chain kernel (k_chain) on the benchmark system S (or --system synth_cut / synth_spring):
one chain at several cluster sizes, then many replicas in one launch.  Prints one JSON line per measurement."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="S")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--clusters", default="1,2,4,8,16")
    ap.add_argument("--replicas", default="")
    ap.add_argument("--multi-cluster", default="1")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--prof", action="store_true", help="per-phase clock sums of thread 0 (pgx_chain_prof)")
    ap.add_argument("--pivot-mode", type=int, default=0)
    ap.add_argument("--prof-rank", type=int, default=0)
    ap.add_argument("--skip", default="0", help="comma-separated phase-skip masks to time (1 R, 2 Q, 4 C, 8 intra, 16 commit, 32 pivot arms)")
    a = ap.parse_args()
    from plum_b200 import mcgen, synth
    from plum_b200.engine import Engine
    import replay
    if a.system == "S":
        r, s, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
    else:
        r, s, types, params = replay.load_golden(a.system)
    ids = types.ids(s.symbol)
    bl, vary = mcgen.bond_settings(r)

    def make(cluster):
        e = Engine(params, device=0, capacity_beads=s.n)
        e.upload(s.xyz, s.q, ids, s.mol_first)
        e.init_energy()
        e.chain_configure(r.phantom, r.move_size, r.move_prob, bl, vary_bond=vary, cluster=cluster, pivot_mode=a.pivot_mode)
        return e

    PHASES = ["header", "load+rows", "proposal", "lists+recip", "real", "cells", "intra", "walls", "reduce", "decide", "commit"]
    for g, skip in [(int(x), int(k)) for x in a.clusters.split(",") if x for k in a.skip.split(",")]:
        e = make(g)
        e.chain_seed(11)
        e.chain_run(200)                      # warm-up (builds the structures)
        if a.prof or skip:
            e.L.pgx_chain_prof.argtypes = [C.c_void_p, C.c_int, C.c_int]
            e.L.pgx_chain_prof_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
            assert e.L.pgx_chain_prof(e.h, 1 if a.prof else 0, skip | (min(a.prof_rank, g - 1) << 8)) == 0
        rec, stop, ms = e.chain_run(a.steps)
        if a.prof:
            buf = (C.c_uint64 * 80)()
            assert e.L.pgx_chain_prof_read(e.h, buf) == 0
            pr = np.array(buf[:], dtype=np.float64).reshape(5, 16)
            tot_cyc = pr[:, :11].sum()
            for k, name in enumerate(["bead", "com", "pivot", "crank", "rept"]):
                if pr[k, 15] > 0:
                    us = pr[k, :11] / tot_cyc * ms * 1e3 / pr[k, 15]
                    print(json.dumps({"prof_kind": name, "cluster": g, "skip": skip, "moves": int(pr[k, 15]),
                                      "us_per_move": float(us.sum()), "phases_us": {p_: round(float(u), 2) for p_, u in zip(PHASES, us)}}), flush=True)
        kinds = rec["kind"]
        out = {"system": a.system, "cluster": g, "skip": skip, "steps": int(len(rec)), "ms": ms, "us_per_step": 1e3 * ms / max(len(rec), 1),
               "moves_per_s": len(rec) / (ms * 1e-3), "accept": float(rec["accept"].mean()),
               "overlap_frac": float((rec["dE"] >= 1e8).mean()), "kind_frac": [float((kinds == k).mean()) for k in range(5)]}
        if a.check:
            e.chain_check()
            t, f = e.totals(), e.recompute_totals()
            out["drift_ewald"] = abs(t["ewald"] - f["ewald"]) / max(1.0, abs(f["ewald"]))
            out["drift_pair"] = abs(t["pair"] - f["pair"]) / max(1.0, abs(f["pair"]))
        print(json.dumps(out), flush=True)
        e.close()

    for g in [int(x) for x in a.multi_cluster.split(",") if x]:
        for R in [int(x) for x in a.replicas.split(",") if x]:
            engs = [make(g) for _ in range(R)]
            for i, e in enumerate(engs):
                e.chain_seed(100 + i)
            L = engs[0].L
            arr = (C.c_void_p * R)(*[e.h for e in engs])
            nd = (C.c_int * R)()
            ms = C.c_float(0.0)
            rc = L.pg_chain_run_multi(arr, R, 100, nd, C.byref(ms))
            assert rc == 0, L.pg_last_error(engs[0].h)
            t0 = time.perf_counter()
            rc = L.pg_chain_run_multi(arr, R, a.steps, nd, C.byref(ms))
            wall = time.perf_counter() - t0
            assert rc == 0, L.pg_last_error(engs[0].h)
            print(json.dumps({"system": a.system, "cluster": g, "replicas": R, "steps": a.steps, "ms": ms.value, "wall_ms": wall * 1e3,
                              "moves_per_s": R * a.steps / (ms.value * 1e-3)}), flush=True)
            for e in engs:
                e.close()


if __name__ == "__main__":
    main()
