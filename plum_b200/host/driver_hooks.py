"""Seed and trace hooks for Plum's MC driver (applied to a throw-away COPY of the
reference sources at build time, never to /root/reference and never committed).

The reference seeds its mt19937 from the wall clock (src/simulation/simulation.cc:134-136)
and has no per-move output, so neither "fixed RNG seed" nor "identical accept/reject
sequence" can be checked without two observation hooks:
  * PLUM_SEED=<n>      overrides the clock seed;
  * PLUM_TRACE=<file>  one line per MC step (format below); PLUM_TRACE_XYZ=1 adds the
                       inputs needed to replay the trace offline.
The same hooks go into both binaries: oracle/_ref/plum_ref (the reference's own
ForceField, built by oracle/build_ref.py) and bin/plum_gpu (this repo's façade, built by
plum_b200/host/build_host.py), so their traces are directly comparable.  No arithmetic
is changed.

Trace format (space separated, doubles as C99 %a):
  I <n_mol> <Epair> <Eewald> <Ebond> <Eext>                        after init
  T <step> <move_type> <mol_id> <dE> <accept> <Epair> <Eewald> <Ebond> <Eext>
  G <step> <I|D> <result> <weight> <n_mol> <Epair> <Eewald> <Ebond> <Eext>
     (result: insertion 0/1, deletion = deleted molecule index or -1;
      weight = Rosenbluth weight at the acceptance test, -1 if never reached)
  V <n_samples> <p_tensor[0..5]> <p_tensor_el[0..15]> <p_tensor_hs[0..15]>
     after every ForceField::CalcPressureVolScalingHSELSlit call (src/force_field/pressure.cc:187-387; the
     reference itself never prints these for systems without walls)
With PLUM_TRACE_XYZ=1 additionally:
  X <n> then n x (<moved> <x> <y> <z>)      trial coordinates of the molecule just tried
  A <n_beads> then n x (<mol> <symbol> <q> <x> <y> <z>)   beads appended by an accepted insertion
  B <current_len> <delete_id> <x1 y1 z1 q1> <x2 y2 z2 q2> <pair_e> <ewald_e> <result>
    <n_chain> then n_chain x (<x> <y> <z> <q>)   one ForceField::BeadsEnergy call
    (reference binary only: src/force_field/cbmc.cc:5-151)
"""
import os


def patch(path, anchor, replacement, count=1):
    with open(path) as f:
        s = f.read()
    if s.count(anchor) != count:
        raise SystemExit(f"driver_hooks: anchor not found exactly {count}x in {path}: {anchor[:60]!r}")
    with open(path, "w") as f:
        f.write(s.replace(anchor, replacement))


TRACE_DECL = r'''
#include <cstdio>
#include <cstdlib>
FILE* plum_trace_file() {
  static FILE* f = NULL; static bool init = false;
  if (!init) { init = true; const char* p = getenv("PLUM_TRACE"); if (p) f = fopen(p, "w"); }
  return f;
}
double plum_trace_weight = -1;
FILE* plum_trace_xyz_file() {
  static int on = -1;
  if (on < 0) { const char* p = getenv("PLUM_TRACE_XYZ"); on = (p && p[0] == '1') ? 1 : 0; }
  return on ? plum_trace_file() : NULL;
}
#define PLUM_TRACE_TOTALS(ff) \
  ((ff).UsePairPot() ? (ff).TotPairEnergy() : 0.0), ((ff).UseEwaldPot() ? (ff).TotEwaldEnergy() : 0.0), \
  ((ff).UseBondPot() ? (ff).TotBondEnergy() : 0.0), ((ff).UseExtPot() ? (ff).TotExtEnergy() : 0.0)
'''


def apply_sim_hooks(src):
    """Hooks in the driver (simulation.cc): seed, per-step trace."""
    sim = os.path.join(src, "simulation", "simulation.cc")
    # Declarations.
    patch(sim, "using namespace std; \n\nSimulation::Simulation() {",
          "using namespace std; \n" + TRACE_DECL + "\nSimulation::Simulation() {")
    # Hook 1: seed.
    patch(sim, "  rand_gen.seed(s);",
          "  if (getenv(\"PLUM_SEED\")) s = (unsigned)strtoul(getenv(\"PLUM_SEED\"), NULL, 10);\n"
          "  rand_gen.seed(s);\n"
          "  if (plum_trace_file()) { fprintf(plum_trace_file(), \"I %d %a %a %a %a\\n\", (int)mols.size(),"
          " PLUM_TRACE_TOTALS(force_field)); }")
    # Hook 2a: translational move trace, right after FinalizeEnergies.
    patch(sim, "      force_field.FinalizeEnergies(mols, accept, mol_id);\n",
          "      force_field.FinalizeEnergies(mols, accept, mol_id);\n"
          "      if (plum_trace_file()) { fprintf(plum_trace_file(), \"T %d %d %d %a %d %a %a %a %a\\n\","
          " step, move_type, mol_id, dE, (int)accept, PLUM_TRACE_TOTALS(force_field)); }\n"
          "      if (plum_trace_xyz_file()) { FILE* tf = plum_trace_xyz_file(); fprintf(tf, \"X %d\", mols[mol_id].Size());\n"
          "        for (int ti = 0; ti < mols[mol_id].Size(); ti++) fprintf(tf, \" %d %a %a %a\", (int)mols[mol_id].bds[ti].GetMoved(),\n"
          "          mols[mol_id].bds[ti].GetCrd(1,0), mols[mol_id].bds[ti].GetCrd(1,1), mols[mol_id].bds[ti].GetCrd(1,2));\n"
          "        fprintf(tf, \"\\n\"); }\n")
    # Hook 2b: GC trace.
    patch(sim, "  // Choose whether to attempt insertion or deletion.\n  bool insert = rand_gen() % 2;\n",
          "  // Choose whether to attempt insertion or deletion.\n  bool insert = rand_gen() % 2;\n"
          "  int plum_gc_result = -1; plum_trace_weight = -1;\n")
    patch(sim, "    bool accept = force_field.CBMCFChainInsertion(mols, rand_gen);\n",
          "    bool accept = force_field.CBMCFChainInsertion(mols, rand_gen);\n"
          "    plum_gc_result = (int)accept;\n")
    patch(sim, "      deleted_chain = force_field.CBMCFChainDeletion(mols, rand_gen);\n",
          "      deleted_chain = force_field.CBMCFChainDeletion(mols, rand_gen);\n"
          "      plum_gc_result = deleted_chain;\n")
    patch(sim, "  UpdateMolCounts();\n  force_field.UpdateMolCounts(mols);\n",
          "  UpdateMolCounts();\n  force_field.UpdateMolCounts(mols);\n"
          "  if (plum_trace_file()) { fprintf(plum_trace_file(), \"G %d %c %d %a %d %a %a %a %a\\n\", step,"
          " insert ? 'I' : 'D', plum_gc_result, plum_trace_weight, (int)mols.size(),"
          " PLUM_TRACE_TOTALS(force_field)); }\n")
    patch(sim, "      force_field.EnergyInitForAddedMolecule(mols);\n",
          "      if (plum_trace_xyz_file()) { FILE* tf = plum_trace_xyz_file(); int nb = 0;\n"
          "        for (int ti = n_mol; ti < (int)mols.size(); ti++) nb += mols[ti].Size();\n"
          "        fprintf(tf, \"A %d\", nb);\n"
          "        for (int ti = n_mol; ti < (int)mols.size(); ti++) for (int tj = 0; tj < mols[ti].Size(); tj++)\n"
          "          fprintf(tf, \" %d %s %a %a %a %a\", ti, mols[ti].bds[tj].Symbol().c_str(), mols[ti].bds[tj].Charge(),\n"
          "            mols[ti].bds[tj].GetCrd(0,0), mols[ti].bds[tj].GetCrd(0,1), mols[ti].bds[tj].GetCrd(0,2));\n"
          "        fprintf(tf, \"\\n\"); }\n"
          "      force_field.EnergyInitForAddedMolecule(mols);\n")


def apply_batch_hook(src):
    """bin/plum_gpu only: lets the façade run stretches of translational steps on the device
    (ForceField::TranslationalBatch).  This is the ONE change to the driver's control flow, and
    the patch a maintainer would carry: at the top of Simulation::Run's loop
    (src/simulation/simulation.cc:220) the façade is offered every step up to the next one at which Sample /
    Print* act; it returns how many it executed (0: a GC step or a crankshaft is next and the untouched code below
    runs it).  PLUM_B200_BATCH=0 disables it."""
    sim = os.path.join(src, "simulation", "simulation.cc")
    patch(sim, "  for (step = 1; step <= steps; step++) {\n    int rand_num = rand_gen();\n",
          "  for (step = 1; step <= steps; step++) {\n"
          "    if (force_field.BatchedMoves()) {\n"
          "      int plum_until = steps;\n"
          "      const int plum_freq[3] = {sample_freq, stat_out_freq, traj_out_freq};\n"
          "      for (int k = 0; k < 3; k++) if (plum_freq[k] > 0) {\n"
          "        long long nx = ((long long)(step + plum_freq[k] - 1) / plum_freq[k]) * plum_freq[k];\n"
          "        if (nx < plum_until) plum_until = (int)nx;\n"
          "      }\n"
          "      const int plum_n = force_field.TranslationalBatch(mols, rand_gen, plum_until - step + 1, step, move_size,\n"
          "                                                        move_prob, attempted, accepted);\n"
          "      if (plum_n > 0) {\n"
          "        step += plum_n - 1;\n"
          "        Sample(); PrintStat(); PrintTraj(); PrintLastCrd(); PrintLastTop(); PrintLastRhoZ();\n"
          "        continue;\n"
          "      }\n"
          "    }\n"
          "    int rand_num = rand_gen();\n")


def apply_pressure_hook(src):
    """Reference binary only: one V line per volume-perturbation pressure sample (pressure.cc:187-387)."""
    prs = os.path.join(src, "force_field", "pressure.cc")
    patch(prs, '#include "../utilities/constants.h"\n',
          '#include "../utilities/constants.h"\n#include <cstdio>\nFILE* plum_trace_file();\n')
    patch(prs, "  delete [] d_com_z;\n",
          "  if (plum_trace_file()) { FILE* tf = plum_trace_file(); fprintf(tf, \"V %d\", (int)vp_z);\n"
          "    for (int ti = 0; ti < 6; ti++) fprintf(tf, \" %a\", p_tensor[ti]);\n"
          "    for (int ti = 0; ti < 16; ti++) fprintf(tf, \" %a\", p_tensor_el[ti]);\n"
          "    for (int ti = 0; ti < 16; ti++) fprintf(tf, \" %a\", p_tensor_hs[ti]);\n"
          "    fprintf(tf, \"\\n\"); }\n"
          "  delete [] d_com_z;\n")


def apply_cbmc_hooks(src):
    """Hooks in the reference's own cbmc.cc (reference binary only)."""
    cbmc = os.path.join(src, "force_field", "cbmc.cc")
    # One line per BeadsEnergy call.
    patch(cbmc, "  energy = pair_e + ewald_e;\n  if (pair_e >= kVeryLargeEnergy) {\n    return kVeryLargeEnergy;\n  }\n  return energy;\n",
          "  energy = pair_e + ewald_e;\n"
          "  if (plum_trace_xyz_file()) { FILE* tf = plum_trace_xyz_file();\n"
          "    fprintf(tf, \"B %d %d %a %a %a %a %a %a %a %a %a %a %a\", current_len, delete_id,\n"
          "      bead1.GetCrd(1,0), bead1.GetCrd(1,1), bead1.GetCrd(1,2), bead1.Charge(),\n"
          "      bead2.GetCrd(1,0), bead2.GetCrd(1,1), bead2.GetCrd(1,2), bead2.Charge(), pair_e, ewald_e,\n"
          "      (pair_e >= kVeryLargeEnergy) ? kVeryLargeEnergy : energy);\n"
          "    int nch = current_len * ((gc_bead_charge != 0) ? 2 : 1); fprintf(tf, \" %d\", nch);\n"
          "    for (int ti = 0; ti < current_len; ti++) fprintf(tf, \" %a %a %a %a\", cbmc_chain[ti].GetCrd(1,0),\n"
          "      cbmc_chain[ti].GetCrd(1,1), cbmc_chain[ti].GetCrd(1,2), cbmc_chain[ti].Charge());\n"
          "    if (gc_bead_charge != 0) for (int ti = gc_chain_len; ti < gc_chain_len + current_len; ti++)\n"
          "      fprintf(tf, \" %a %a %a %a\", cbmc_chain[ti].GetCrd(1,0), cbmc_chain[ti].GetCrd(1,1),\n"
          "        cbmc_chain[ti].GetCrd(1,2), cbmc_chain[ti].Charge());\n"
          "    fprintf(tf, \"\\n\"); }\n"
          "  if (pair_e >= kVeryLargeEnergy) {\n    return kVeryLargeEnergy;\n  }\n  return energy;\n")
    # Rosenbluth weight at the two acceptance tests.
    patch(cbmc, '#include "../utilities/constants.h"\n',
          '#include "../utilities/constants.h"\n#include <cstdio>\nextern double plum_trace_weight;\n'
          'FILE* plum_trace_xyz_file();\n')
    patch(cbmc, "  if (rand_num < (exp(beta*chem_pot) * weight) * C) {",
          "  plum_trace_weight = weight;\n  if (rand_num < (exp(beta*chem_pot) * weight) * C) {")
    patch(cbmc, "  if (rand_num < C / (exp(beta*chem_pot) * weight)) {",
          "  plum_trace_weight = weight;\n  if (rand_num < C / (exp(beta*chem_pot) * weight)) {")


