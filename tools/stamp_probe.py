#!/usr/bin/env python3
"""In-kernel %globaltimer breakdown of k_move (debug build flag PLUM_B200_TIMING=1)."""
import os, sys, ctypes as C
os.environ["PLUM_B200_TIMING"] = "1"
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from plum_b200 import synth
from plum_b200.engine import Engine
r, s, types, params = synth.load(cache_dir=os.path.join(REPO, "gpurun_out", "cache"))
eng = Engine(params, device=0, capacity_beads=s.n)
eng.upload(s.xyz, s.q, types.ids(s.symbol), s.mol_first); eng.init_energy()
eng.L.pgx_read_timing.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
rng = np.random.default_rng(0)
chains = [m for m in range(s.n_mol) if s.mol_first[m+1]-s.mol_first[m] > 1]
ions = [m for m in range(s.n_mol) if s.mol_first[m+1]-s.mol_first[m] == 1]
buf = np.zeros((8192, 8), dtype=np.uint64)
for name, pool in (("ion", ions), ("chain", chains)):
    agg = []
    for it in range(40):
        m = int(rng.choice(pool)); f, l = s.mol_first[m], s.mol_first[m+1]
        eng.delta_e(m, s.xyz[f:l] + rng.normal(scale=0.3, size=(l-f, 3)), np.ones(l-f, dtype=np.uint8)); eng.commit(False)
        n = eng.L.pgx_read_timing(eng.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 8192)
        t = buf[:n].astype(np.int64)
        t0 = t[:, 0].min()
        last = int(np.argmax(t[:, 7]))
        if it >= 10:
            agg.append([t[:, 0].max() - t0, (t[:, 1] - t[:, 0])[t[:, 1] > 0].mean() if (t[:,1]>0).any() else 0,
                        np.median(t[:, 2] - t[:, 0]), t[:, 2].max() - t0, t[:, 4].max() - t0,
                        t[last, 5] - t0, t[last, 6] - t0, t[last, 7] - t0, np.median(t[:,4]-t[:,3]), np.median(t[:,3]-t[:,2])])
    a = np.array(agg).mean(axis=0)
    print(f"{name}: ctas={n} last_cta_start={a[0]:.0f}ns stage={a[1]:.0f}ns median_main_done={a[2]:.0f}ns all_main_done={a[3]:.0f}ns "
          f"all_atomic_done={a[4]:.0f}ns final_start={a[5]:.0f}ns final_serial_start={a[6]:.0f}ns end={a[7]:.0f}ns | fence+atomic={a[8]:.0f}ns blocksum={a[9]:.0f}ns")

# detail of the last chain launch
t = buf[:n].astype(np.int64); t0 = t[:, 0].min()
d = t[:, 4] - t0
print("atomic_done percentiles (ns):", [int(np.percentile(d, q)) for q in (5, 25, 50, 75, 90, 95, 99, 100)])
order = np.argsort(-d)[:16]
print("slowest CTAs: index [start, stage_done, main_done(thread0), partial_written, atomic_done]")
for c in order:
    print(int(c), [int(x - t0) for x in t[c, :5]])
print("first 6 CTAs (k role):")
for c in range(6):
    print(int(c), [int(x - t0) for x in t[c, :5]])
print("CTAs 56..60 (intra role):")
for c in range(56, 61):
    print(int(c), [int(x - t0) for x in t[c, :5]])
# per-SM view
sm = t[:, 6]
role = np.where(np.arange(n) < 148, 0, 1)
print("per SM: n_ctas, n_helpers, last finish(ns) -- first 12 SMs and the 6 slowest")
rows = []
for sid in np.unique(sm):
    m_ = sm == sid
    rows.append((int(sid), int(m_.sum()), int((m_ & (role == 0)).sum()), int(d[m_].max()), int(d[m_].min())))
for r_ in rows[:12]: print(r_)
print("slowest SMs:", sorted(rows, key=lambda r_: -r_[3])[:6])
print("fastest SMs:", sorted(rows, key=lambda r_: r_[3])[:6])
hd = d[:148]; pdur = d[148:]
print("helpers finish: min/med/max", int(hd.min()), int(np.median(hd)), int(hd.max()))
print("pair finish: min/med/max", int(pdur.min()), int(np.median(pdur)), int(pdur.max()))

print("pair CTAs 148..170: [start, last_stage_done, main_done(thread0), partial_written, atomic_done] smid")
for c in range(148, 171):
    print(int(c), [int(x - t0) for x in t[c, :5]], int(t[c, 6]))
