import sys, ctypes as C, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import replay
from plum_b200.engine import Engine
r, s, types, params = replay.load_golden("synth_cut")
ids = types.ids(s.symbol)
engs=[]
for _ in range(2):
    e = Engine(params, device=0, capacity_beads=s.n); e.upload(s.xyz, s.q, ids, s.mol_first); engs.append(e)
t0 = engs[0].init_energy()
for rk, e in enumerate(engs): e.sk_attach_local(rk, engs)
def flags(e):
    b=(C.c_uint*33)(); e.L.pgx_sk_flags(e.h, b); return list(b[:3]), list(b[16:19]), b[32]
print("before", [flags(e) for e in engs])
for e in engs: e.recompute_sk_begin()
import time; time.sleep(0.5)
print("after begin", [flags(e) for e in engs])
for e in engs:
    try:
        print(e.recompute_sk_end())
    except Exception as ex:
        print("ERR", ex)
print("after end", [flags(e) for e in engs])
