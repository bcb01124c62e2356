#!/usr/bin/env python3
"""Wall time of Plum's driver on the four reference examples: B200 façade (bin/plum_gpu) vs the reference
binary on one host core (oracle/_ref/plum_ref)."""
import os, shutil, subprocess, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def run(binary, ex, steps):
    d = tempfile.mkdtemp()
    for f in ("run.in", "input_crd.dat", "input_top.dat"):
        shutil.copy(os.path.join(REPO, "tests/golden/examples", ex, f), d)
    txt = open(os.path.join(d, "run.in")).read().split("\n")
    txt = [f"s1_total_simulation_steps {steps}" if l.startswith("s1_total_simulation_steps") else l for l in txt]
    open(os.path.join(d, "run.in"), "w").write("\n".join(txt))
    t0 = time.perf_counter()
    subprocess.check_call([binary], stdin=open(os.path.join(d, "run.in")), stdout=open(os.path.join(d, "run.log"), "w"),
                          cwd=d, env=dict(os.environ, PLUM_SEED="1"))
    dt = time.perf_counter() - t0
    shutil.rmtree(d)
    return dt
gpu, ref = os.path.join(REPO, "bin/plum_gpu"), os.path.join(REPO, "oracle/_ref/plum_ref")
for ex in ("bulk_nvt", "confined_nvt", "bulk_muvt", "confined_muvt"):
    g0 = run(gpu, ex, 0); g = run(gpu, ex, 100000)
    r0 = run(ref, ex, 0); r = run(ref, ex, 3000)
    print(f"{ex}: plum_gpu init {g0:.2f}s, 1e5 steps {g - g0:.2f}s ({1e5/(g-g0):.0f} steps/s) | plum_ref init {r0:.2f}s, "
          f"3000 steps {r - r0:.2f}s ({3000/(r-r0):.0f} steps/s) | speed-up {(1e5/(g-g0))/(3000/(r-r0)):.0f}x", flush=True)
