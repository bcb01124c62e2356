#!/usr/bin/env python3
"""In-kernel %globaltimer breakdown of the generic k_move path on a reference example."""
import os, sys, time, ctypes as C
os.environ["PLUM_B200_TIMING"] = "1"
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import replay
from plum_b200.engine import Engine
from plum_b200._abi import PgDelta
for name in sys.argv[1:] or ("bulk_nvt", "confined_nvt"):
    r, s, types, params = replay.load_golden(name)
    eng = Engine(params, device=0, capacity_beads=s.n + 64)
    eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first); eng.init_energy()
    info = eng.ewald_info()
    print(name, "N", s.n, "nk_half", info.n_k_half, "real_cell", list(info.real_cell), "real_cutoff", info.real_cutoff, flush=True)
    eng.L.pgx_read_timing.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
    rng = np.random.default_rng(0)
    chains = [m for m in range(r.phantom, s.n_mol) if s.mol_first[m+1]-s.mol_first[m] > 1]
    ions = [m for m in range(r.phantom, s.n_mol) if s.mol_first[m+1]-s.mol_first[m] == 1]
    buf = np.zeros((8192, 8), dtype=np.uint64)
    d = PgDelta()
    for label, pool in (("ion", ions), ("chain", chains)):
        if not pool: continue
        for it in range(6):
            m = int(rng.choice(pool)); f, l = s.mol_first[m], s.mol_first[m+1]
            x = np.ascontiguousarray(s.xyz[f:l] + rng.normal(scale=0.3, size=(l-f, 3))); v = np.ones(l-f, dtype=np.uint8)
            t0 = time.perf_counter(); eng.delta_e_raw(m, x, v, d); t1 = time.perf_counter(); eng.L.pg_commit(eng.h, 0)
            n = eng.L.pgx_read_timing(eng.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 8192)
        t = buf[:n].astype(np.int64); b = t[:, 0].min()
        print(f"  {label}: host {1e6*(t1-t0):.1f} us, ctas {n}, last start {t[:,0].max()-b} ns, main done med/max {int(np.median(t[:,2]-b))}/{t[:,2].max()-b} ns, "
              f"atomic done max {t[:,4].max()-b} ns, end {t[:,7].max()-b} ns", flush=True)
        for c in list(range(min(n, 4))) + [n - 1]:
            print("    cta", c, [int(x_ - b) for x_ in t[c, :6]])
    eng.close()
