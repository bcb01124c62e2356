// B200 façade behind Plum's ForceField interface — see force_field.h.
//
// What each member replaces in the reference (file:line relative to /root/reference):
//   Initialize            src/force_field/force_field.cc:46-363 (+ the potentials' ReadParameters)
//   EnergyDifference      src/force_field/force_field.cc:407-434   -> pg_delta_e
//   FinalizeEnergies      src/force_field/force_field.cc:436-451   -> pg_commit
//   CBMCFGenTrialBeads    src/force_field/cbmc.cc:161-212          -> pg_trial_energies (batched)
//   CBMCFChainInsertion   src/force_field/cbmc.cc:214-328
//   CBMCFChainDeletion    src/force_field/cbmc.cc:330-442          -> pg_delete_molecules
//   EnergyInitForAddedMolecule  src/force_field/force_field.cc:1088-1105 -> pg_insert_molecules
//   UpdateMolCounts       src/force_field/force_field.cc:1258-1277
// The random-number draws are made in exactly the reference's order: a trajectory is
// identical as long as every accept/reject and roulette decision agrees.
#include "force_field.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "../utilities/constants.h"
#include "../utilities/misc.h"
#include "plum_b200.h"
#include "mc_propose.h"
#include "mt_state.h"

#include <cstdio>
#include <cstdlib>
#include <ctime>
using namespace std;

// Trace hooks of the driver build (plum_b200/host/driver_hooks.py); absent otherwise.
FILE* plum_trace_file() __attribute__((weak));
FILE* plum_trace_xyz_file() __attribute__((weak));

// Set by the driver's trace hook build (plum_b200/host/driver_hooks.py); harmless otherwise.
double plum_trace_weight __attribute__((weak)) = -1;

namespace {
double Uniform(mt19937& g) { return (double)g() / g.max(); }

// PLUM_B200_PROFILE=1: wall time spent inside each ABI entry point, printed to stderr at exit.
enum { kTDelta, kTCommit, kTTrials, kTInsert, kTDelete, kTTotals, kTWallForce, kTVolScale, kTGcInsert, kTGcDelete, kTBatch, kTSites };
const char* const kSiteName[kTSites] = {"pg_delta_e", "pg_commit", "pg_trial_energies", "pg_insert_molecules",
                                        "pg_delete_molecules", "pg_get_totals", "pg_wall_force", "pg_vol_scaling_sample",
                                        "CBMCFChainInsertion", "CBMCFChainDeletion", "TranslationalBatch"};
bool prof_on = getenv("PLUM_B200_PROFILE") != NULL;
double prof_s[kTSites];
long prof_n[kTSites];
timespec prof_t_run;          // end of InitializeEnergy: everything behind it is the driver's run loop
bool prof_run_started = false;
struct SiteTimer {
  int site;
  timespec t0;
  explicit SiteTimer(int s) : site(s) { if (prof_on) clock_gettime(CLOCK_MONOTONIC, &t0); }
  ~SiteTimer() {
    if (!prof_on) return;
    timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    prof_s[site] += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    prof_n[site]++;
  }
};
}  // namespace

namespace {
struct McState {
  plum_mc::Proposer prop;
  plum_mc::Batch batch;
  plum_mc::BatchSizer sizer;
  int n_mol_configured = -1;   // -1: (re)build the molecule tables (set by UpdateMolCounts after every GC move)
  bool sizer_ready = false;
  int cooldown = 0;            // steps left to the driver's per-move path because batches kept stopping early
  double avg_len = 256.0;      // smoothed steps per batch (GC steps cut batches short, too)
  vector<double> dE, xyz;
  vector<uint8_t> acc;
  // device-resident chain (pg_chain_*): -1 not tried yet, 0 not offered for this system (the pg_mc_* path is used), 1 on
  int chain = -1;
  vector<pg_chain_step> steps;
  mt19937 rng_handed;          // the driver's generator as the device handed it back last time
  bool rng_resident = false;   // ... the device still holds that state: no upload unless the driver has drawn since
};

// std::mt19937 <-> 624 state words + position: the textual form operator<< / operator>> define.
void MtExport(const mt19937& g, uint32_t* state, int* pos) { plum_mt::export_state(g, state, pos); }
void MtImport(mt19937& g, const uint32_t* state, int pos) { plum_mt::import_state(g, state, pos); }
}  // namespace

ForceField::ForceField() : vp_z(0), engine(NULL), tot_valid(false), pending_mol(-1), mc_state(NULL) {
  for (int i = 0; i < 12; i++) p_tensor[i] = 0;
  for (int i = 0; i < 20; i++) p_tensor2[i] = p_tensor3[i] = p_tensor_hs[i] = p_tensor_el[i] = 0;
}

ForceField::~ForceField() {
  if (prof_on && prof_run_started) {
    timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double w = (t1.tv_sec - prof_t_run.tv_sec) + 1e-9 * (t1.tv_nsec - prof_t_run.tv_nsec);
    cerr << "plum_b200 profile: run_wall calls 1 total " << w << " s, " << 1e6 * w << " us/call" << endl;
  }
  if (prof_on)
    for (int i = 0; i < kTSites; i++)
      if (prof_n[i])
        cerr << "plum_b200 profile: " << kSiteName[i] << " calls " << prof_n[i] << " total " << prof_s[i] << " s, "
             << 1e6 * prof_s[i] / prof_n[i] << " us/call" << endl;
  if (engine) pg_destroy(engine);
}

void ForceField::Fail(const char* what, int rc) {
  cout << "  plum_b200: " << what << " failed (" << rc << "): " << pg_last_error(engine) << endl;
  cout << "  There is no CPU fallback. Exiting! Program complete." << endl;
  exit(1);
}

int ForceField::TypeId(const string& symbol) {
  map<string, int>::iterator it = type_of.find(symbol);
  if (it != type_of.end()) return it->second;
  if (engine) {
    cout << "  plum_b200: bead type \"" << symbol << "\" appeared after the engine was created."
         << " Exiting! Program complete." << endl;
    exit(1);
  }
  int id = (int)type_symbols.size();
  type_of[symbol] = id;
  type_symbols.push_back(symbol);
  return id;
}

// The s2/s3 part of run.in: same tokens, same order, flag names discarded.
void ForceField::ReadPotentialParameters(vector<Molecule>& mols) {
  string flag, symbol;
  cin >> flag >> vp_el_res;
  cin >> flag >> vp_bead_size;
  cin >> flag >> use_pair_pot;
  cin >> flag >> use_ewald_pot;
  cin >> flag >> use_bond_pot;
  cin >> flag >> use_bond_rigid;
  cin >> flag >> use_angle_pot;
  cin >> flag >> use_dihed_pot;
  cin >> flag >> use_ext_pot;
  cout << "  Force field parameters (B200 engine)." << endl;
  cout << setw(35) << "[P] g(r) bin resolution (ul): " << vp_el_res << endl;
  cout << setw(35) << "[P] Hard sphere size (ul)   : " << vp_bead_size << endl;
  cout << setw(35) << "Use pair potential?         : " << YesOrNo(use_pair_pot) << endl;
  cout << setw(35) << "Use Ewald sum potential?    : " << YesOrNo(use_ewald_pot) << endl;
  cout << setw(35) << "Use bond potential?         : " << YesOrNo(use_bond_pot) << endl;
  cout << setw(35) << "Use rigid bond?             : " << YesOrNo(use_bond_rigid) << endl;
  cout << setw(35) << "Use angle potential?        : " << YesOrNo(use_angle_pot) << endl;
  cout << setw(35) << "Use dihedral potential?     : " << YesOrNo(use_dihed_pot) << endl;
  cout << setw(35) << "Use external potential?     : " << YesOrNo(use_ext_pot) << endl;
  cin >> flag >> use_gc;
  cout << setw(35) << "Use grand canonical MC?     : " << YesOrNo(use_gc) << endl;
  if (use_gc) {
    cin >> flag >> chem_pot >> flag >> gc_deBroglie_prefactor >> flag >> gc_freq >> flag >> gc_chain_len >> flag >>
        gc_bead_charge >> flag >> gc_bead_symbol >> flag >> cbmc_no_of_trials;
    cout << setw(35) << "[GC] Chemical pot (kBT)     : " << chem_pot << endl;
    cout << setw(35) << "[GC] de Broglie wavelength^3: " << gc_deBroglie_prefactor << endl;
    cout << setw(35) << "[GC] GCMC move frequency    : " << gc_freq << endl;
    cout << setw(35) << "[GC] CBMC chain length      : " << gc_chain_len << endl;
    cout << setw(35) << "[GC] CBMC bead charge (e)   : " << gc_bead_charge << endl;
    cout << setw(35) << "[GC] CBMC bead type         : " << gc_bead_symbol << endl;
    cout << setw(35) << "[GC] CBMC trials            : " << cbmc_no_of_trials << endl;
  }

  if (use_pair_pot) {
    cin >> flag >> pair_name;
    if (pair_name == "TruncatedLJ") {
      cout << setw(35) << "[PP] Pair potential type    : " << "Truncated LJ" << endl;
      cin >> flag >> lj_cutoff;
      cout << setw(35) << "[PP] Pair potential cutoff  : " << lj_cutoff << endl;
      while (true) {
        cin >> flag >> symbol;
        if (symbol == "end") break;
        double sigma, epsilon;
        cin >> flag >> sigma >> flag >> epsilon;
        lj_sigmas[symbol] = sigma;
        lj_epsilons[symbol] = epsilon;
        TypeId(symbol);
        cout << setw(35) << "[PP] Sigma for bead type    : " << symbol << " - " << sigma << endl;
        cout << setw(35) << "[PP] Epsilon for bead type  : " << symbol << " - " << epsilon << endl;
      }
    } else if (pair_name == "HardSphere") {
      cout << setw(35) << "[PP] Pair potential type    : " << "Hard sphere" << endl;
      while (true) {
        cin >> flag >> symbol;
        if (symbol == "end") break;
        double radius;
        cin >> flag >> radius;
        hs_radii[symbol] = radius;
        TypeId(symbol);
        cout << setw(35) << "[PP] Bead type and r (ul)   : " << symbol << " - " << radius << endl;
      }
    } else {
      cout << "ForceField::ReadParameters: " << endl;
      cout << "  Undefined pair potential! Exiting." << endl;
      exit(1);
    }
  }
  if (use_ewald_pot) {
    cin >> flag >> ewald_name;
    if (ewald_name != "Coul") {
      cout << "ForceField::ReadParameters: " << endl;
      cout << "  Undefined Ewald potential! Exiting." << endl;
      exit(1);
    }
    cout << setw(35) << "[EP] Ewald potential type   : " << "Coul" << endl;
    cin >> flag >> lB >> flag >> dielectric >> flag >> alpha >> flag >> dipole_correction;
  }
  if (use_bond_pot) {
    cin >> flag >> bond_name;
    if (bond_name != "Spring") {
      cout << "ForceField::ReadParameters: " << endl;
      cout << "  Undefined bonded potential! Exiting." << endl;
      exit(1);
    }
    cout << setw(35) << "Bond potential type         : " << "Harmonic spring" << endl;
    cin >> flag >> bond_k >> flag >> bond_r0;
    cout << setw(37) << "Bond strength in eV/ul^2    = " << bond_k << endl;
    cout << setw(37) << "Resting bond length in ul   = " << bond_r0 << endl;
  }
  if (use_bond_rigid) {
    cin >> flag >> rigid_bond;
    cout << setw(35) << "[RB] Rigid bond length      : " << rigid_bond << "A" << endl;
  }
  if (use_ext_pot) {
    cin >> flag >> ext_name;
    if (ext_name == "TruncatedLJWall") {
      cout << setw(35) << "[eP] External potential type: " << "Truncated LJ wall" << endl;
      cin >> flag >> wall_cut >> flag >> wall_sig >> flag >> wall_eps;
      cout << setw(35) << "[eP] LJ cutoff              : " << wall_cut << endl;
      cout << setw(35) << "[eP] Sigma for wall         : " << wall_sig << endl;
      cout << setw(35) << "[eP] Epsilon for wall       : " << wall_eps << endl;
      while (true) {
        cin >> flag >> symbol;
        if (symbol == "end") break;
        double sigma, epsilon;
        cin >> flag >> sigma >> flag >> epsilon;
        wall_sigmas[symbol] = sigma;
        wall_epsilons[symbol] = epsilon;
        TypeId(symbol);
        cout << setw(35) << "[eP] Sigma for bead type    : " << symbol << " - " << sigma << endl;
        cout << setw(35) << "[eP] Epsilon for bead type  : " << symbol << " - " << epsilon << endl;
      }
    } else if (ext_name == "HardWall" || ext_name == "WellWall") {
      if (ext_name == "WellWall") {
        cout << "  Note: Currently, no correct pressure calculation" << endl;
        cout << "        routine exists for well wall potential." << endl;
        cin >> flag >> well_width;
        cin >> flag >> well_depth;
        cout << setw(31) << "Well witdh in unit length   = " << well_width << endl;
        cout << setw(31) << "Well depth in kT            = " << well_depth << endl;
      }
      while (true) {
        cin >> flag >> symbol;
        if (symbol == "end") break;
        double radius;
        cin >> flag >> radius;
        wall_sigmas[symbol] = radius;
        TypeId(symbol);
        cout << setw(35) << "[eP] Radius for bead type   : " << symbol << " - " << radius << endl;
      }
    } else {
      cout << "ForceField::ReadParameters: " << endl;
      cout << "  Undefined external potential! Exiting." << endl;
      exit(1);
    }
  }
  if (vp_bead_size < 0) {
    cout << "  Virial bead size has to be bigger than zero!" << "Exiting! Program complete." << endl;
    exit(1);
  }
  if (use_ext_pot) {
    cout << "  Note: [!!!] If a wall confining potential is used, the" << endl;
    cout << "        z-length of the simulation box needs to be 3 to" << endl;
    cout << "        5 times wider than the confinement for the" << endl;
    cout << "        electrostatics calculation to be correct!" << endl;
  }
  // Every symbol present in the coordinates gets a type id too (absent table entries read as 0,
  // like operator[] on the reference's std::map).
  for (int i = 0; i < (int)mols.size(); i++)
    for (int j = 0; j < mols[i].Size(); j++) TypeId(mols[i].bds[j].Symbol());
}

void ForceField::CreateEngine(vector<Molecule>& mols) {
  TypeId(gc_bead_symbol);
  const int nt = (int)type_symbols.size();
  vector<double> ljs(nt, 0.0), lje(nt, 0.0), hsr(nt, 0.0), ws(nt, 0.0), we(nt, 0.0);
  vector<int32_t> graft(nt, PG_GRAFT_NONE);
  for (int t = 0; t < nt; t++) {
    const string& s = type_symbols[t];
    if (lj_sigmas.count(s)) ljs[t] = lj_sigmas[s];
    if (lj_epsilons.count(s)) lje[t] = lj_epsilons[s];
    if (hs_radii.count(s)) hsr[t] = hs_radii[s];
    if (wall_sigmas.count(s)) ws[t] = wall_sigmas[s];
    if (wall_epsilons.count(s)) we[t] = wall_epsilons[s];
    if (s == "L") graft[t] = PG_GRAFT_LEFT;
    if (s == "R") graft[t] = PG_GRAFT_RIGHT;
  }
  pg_params p;
  for (int i = 0; i < 3; i++) p.box[i] = box_l[i];
  p.npbc = npbc;
  p.n_types = nt;
  p.beta = beta;
  p.pair_kind = !use_pair_pot ? PG_PAIR_NONE : (pair_name == "HardSphere" ? PG_PAIR_HARD_SPHERE : PG_PAIR_TRUNCATED_LJ);
  p._pad0 = 0;
  p.lj_cutoff = lj_cutoff;
  p.lj_sigma = ljs.data();
  p.lj_epsilon = lje.data();
  p.hs_radius = hsr.data();
  p.use_ewald = use_ewald_pot ? 1 : 0;
  p.dipole_correction = (use_ewald_pot && dipole_correction) ? 1 : 0;
  p.lB = lB;
  p.alpha = alpha;
  p.bond_kind = use_bond_pot ? PG_BOND_SPRING : PG_BOND_NONE;
  p._pad1 = 0;
  p.bond_k = bond_k;
  p.bond_r0 = bond_r0;
  p.ext_kind = !use_ext_pot ? PG_EXT_NONE
                            : (ext_name == "HardWall" ? PG_EXT_HARD_WALL
                                                      : (ext_name == "WellWall" ? PG_EXT_WELL_WALL : PG_EXT_TRUNCATED_LJ_WALL));
  p._pad2 = 0;
  p.wall_cut = wall_cut;
  p.wall_sigma = ws.data();
  p.wall_epsilon = we.data();
  p.graft_kind = graft.data();
  p.well_width = well_width;
  p.well_depth = well_depth;
  int device = 0;
  if (getenv("PLUM_B200_DEVICE")) device = atoi(getenv("PLUM_B200_DEVICE"));
  int n_beads = 0;
  for (int i = 0; i < (int)mols.size(); i++) n_beads += mols[i].Size();
  int rc = pg_create(&p, device, n_beads + 4096, &engine);
  if (rc) Fail("pg_create", rc);
  if (use_ewald_pot) {
    pg_ewald_info info;
    pg_get_ewald_info(engine, &info);
    cout << setw(35) << "[EP] Bjerrum length (ul)    : " << lB << endl;
    cout << setw(35) << "[EP] Dielectric constant    : " << dielectric << endl;
    cout << setw(35) << "[EP] Ewald alpha (ul^-2)    : " << alpha << endl;
    cout << setw(35) << "[EP] Real space cutoff (ul) : " << info.real_cutoff << endl;
    cout << setw(35) << "[EP] (cell sum number)      : "
         << ceil((info.real_cell[0] + info.real_cell[1] + info.real_cell[2]) / 3.0) << endl;
    cout << setw(35) << "[EP] k space cutoff (ul^-2) : " << info.repl_cutoff << endl;
    cout << setw(35) << "[EP] (k sum number)         : "
         << ceil((info.repl_cell[0] + info.repl_cell[1] + info.repl_cell[2]) / 3.0) << endl;
    cout << setw(35) << "[EP] k vectors on device    : " << info.n_k_half << " (half space of " << info.n_k << ")"
         << endl;
    cout << setw(35) << "[EP] Use dipole correction? : " << YesOrNo(dipole_correction) << endl;
  }
}

void ForceField::UploadSystem(vector<Molecule>& mols) {
  vector<double> xyz, q;
  vector<int32_t> type, first;
  first.push_back(0);
  for (int i = 0; i < (int)mols.size(); i++) {
    for (int j = 0; j < mols[i].Size(); j++) {
      Bead& b = mols[i].bds[j];
      xyz.push_back(b.GetCrd(0, 0));
      xyz.push_back(b.GetCrd(0, 1));
      xyz.push_back(b.GetCrd(0, 2));
      q.push_back(b.Charge());
      type.push_back(TypeId(b.Symbol()));
    }
    first.push_back((int32_t)q.size());
  }
  tot_valid = false;
  int rc = pg_upload_system(engine, (int)q.size(), xyz.data(), q.data(), type.data(), (int)mols.size(), first.data());
  if (rc) Fail("pg_upload_system", rc);
}

void ForceField::Initialize(double beta_in, int npbc_in, double box_l_in[3], vector<Molecule>& mols, int phantom_in,
                            int coion_in, int grafted_in, int grafted_counterion_in) {
  beta = beta_in;
  npbc = npbc_in;
  rigid_bond = -1;
  for (int i = 0; i < 3; i++) box_l[i] = box_l_in[i];
  vol = box_l[0] * box_l[1] * box_l[2];
  phantom = phantom_in;
  coion = coion_in;
  grafted = grafted_in;
  grafted_counterion = grafted_counterion_in;
  use_gc = false;
  chem_pot = 0; gc_freq = 1; gc_bead_charge = 0; gc_chain_len = 1;
  lj_cutoff = -1; lB = 0; dielectric = 0; alpha = 0; dipole_correction = false;
  bond_k = 0; bond_r0 = 0; wall_cut = -1; wall_sig = 0; wall_eps = 0; well_width = 0; well_depth = 0;

  ReadPotentialParameters(mols);

  // Chain bookkeeping, force_field.cc:236-284.
  mu_tot_ins = 50;
  chain_len = -1;
  if (use_gc) {
    chain_len = gc_chain_len;
  } else {
    for (int i = phantom; i < (int)mols.size(); i++) {
      chain_len = mols[i].Size();
      gc_bead_charge = mols[i].bds[0].Charge();
      if (chain_len > 1) break;
    }
    gc_deBroglie_prefactor = 1;
    gc_chain_len = chain_len;
    gc_bead_symbol = "P";
    cbmc_no_of_trials = 30;
  }
  UpdateMolCounts(mols);
  n_chain += grafted;   // the start-up count of force_field.cc:260-278 does not subtract grafted chains
  if (gc_chain_len < 0) gc_chain_len = 0;
  gc_chain_chg.assign(gc_chain_len, (double)gc_bead_charge);

  CreateEngine(mols);
  InitializeEnergy(mols);
}

void ForceField::InitializeEnergy(vector<Molecule>& mols) {
  UploadSystem(mols);
  pg_totals t;
  tot_valid = false;
  int rc = pg_init_energy(engine, &t);
  if (rc) Fail("pg_init_energy", rc);
  if (use_pair_pot) cout << "  Initialized pair potential." << endl;
  if (use_ewald_pot) cout << "  Initialized Ewald potential." << endl;
  if (use_bond_pot) cout << "  Initialized bond potential." << endl;
  if (use_ext_pot) cout << "  Initialized external potential." << endl;

  // CBMC scratch beads, force_field.cc:383-403: chain monomers then their counter-ions.
  cbmc_chain.clear();
  cbmc_trial_beads.clear();
  for (int i = 0; i < gc_chain_len; i++) cbmc_chain.push_back(Bead(gc_bead_symbol, -1, -1, gc_chain_chg[i], 0, 0, 0));
  for (int i = 0; i < cbmc_no_of_trials; i++)
    cbmc_trial_beads.push_back(Bead(gc_bead_symbol, -1, -1, gc_chain_len ? gc_chain_chg[0] : 0.0, 0, 0, 0));
  for (int i = 0; i < gc_chain_len; i++) cbmc_chain.push_back(Bead(gc_bead_symbol, -1, -1, -gc_chain_chg[i], 0, 0, 0));
  for (int i = 0; i < cbmc_no_of_trials; i++)
    cbmc_trial_beads.push_back(Bead(gc_bead_symbol, -1, -1, gc_chain_len ? -gc_chain_chg[0] : 0.0, 0, 0, 0));
  cbmc_trial_weights.assign(cbmc_no_of_trials, 0.0);
  if (prof_on && !prof_run_started) { clock_gettime(CLOCK_MONOTONIC, &prof_t_run); prof_run_started = true; }
}

// ------------------------------------------------------------------ per move
double ForceField::EnergyDifference(vector<Molecule>& mols, int moved_mol) {
  Molecule& m = mols[moved_mol];
  const int len = m.Size();
  vector<double> trial(3 * len);
  vector<uint8_t> moved(len);
  for (int i = 0; i < len; i++) {
    trial[3 * i] = m.bds[i].GetCrd(1, 0);
    trial[3 * i + 1] = m.bds[i].GetCrd(1, 1);
    trial[3 * i + 2] = m.bds[i].GetCrd(1, 2);
    moved[i] = m.bds[i].GetMoved() ? 1 : 0;
  }
  pg_delta d;
  int rc;
  { SiteTimer st_(kTDelta); rc = pg_delta_e(engine, moved_mol, trial.data(), moved.data(), &d); }
  if (rc) Fail("pg_delta_e", rc);
  pending_mol = moved_mol;
  return d.dE;
}

void ForceField::FinalizeEnergies(vector<Molecule>& mols, bool accept, int moved_mol) {
  (void)mols;
  if (pending_mol != moved_mol) return;   // EnergyDifference was skipped by the driver
  int rc;
  tot_valid = false;
  { SiteTimer st_(kTCommit); rc = pg_commit(engine, accept ? 1 : 0); }
  if (rc) Fail("pg_commit", rc);
  pending_mol = -1;
  SkDriftReset(1);
}


// ------------------------------------------------- batched translational steps
// PLUM_B200_BATCH: unset / 1 = batches where they pay (see BatchSizer), 0 = never, 2 = always (tests).
static int BatchMode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("PLUM_B200_BATCH"); mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }
  return mode;
}
bool ForceField::BatchedMoves() { return BatchMode() != 0; }

// S(k) is carried incrementally (S += dS on every accepted move), so rounding drift grows with the number of updates;
// every PLUM_B200_SK_RESET steps (default 2^20, 0 = never) it is recomputed from the resident coordinates
// (pg_recompute_sk: the totals the driver prints are accumulated like the reference's and are not touched).
void ForceField::SkDriftReset(int steps_done) {
  static const long long every = [] { const char* e = getenv("PLUM_B200_SK_RESET"); return e ? atoll(e) : (1LL << 20); }();
  if (every <= 0 || !use_ewald_pot) return;
  sk_steps += steps_done;
  if (sk_steps < every) return;
  sk_steps = 0;
  tot_valid = false;
  int rc = pg_recompute_sk(engine, NULL);
  if (rc) Fail("pg_recompute_sk", rc);
}

int ForceField::TranslationalBatch(vector<Molecule>& mols, mt19937& rand_gen, int max_steps, int first_step,
                                   double move_size, const double move_prob[5], int attempted[], int accepted[]) {
  if (max_steps <= 0) return 0;
  SiteTimer whole_(kTBatch);
  if (!mc_state) mc_state = new McState();
  McState& S = *static_cast<McState*>(mc_state);
  // Where most steps end in an overlap (dE >= 1e8: dense systems, long pivots) a batch stops after a handful of
  // steps and costs more than it saves: leave the next stretch to the driver's per-move path, then probe again.
  if (BatchMode() == 2) S.cooldown = 0;
  if (S.cooldown > 0) { S.cooldown--; return 0; }
  // the molecule list only changes through GC moves, which also change its length or leave it as it was
  if (S.n_mol_configured != (int)mols.size()) {
    plum_mc::Config cfg;
    cfg.phantom = phantom;
    for (int i = 0; i < (int)mols.size(); i++) cfg.mol_len.push_back(mols[i].Size());
    cfg.move_size = move_size;
    for (int i = 0; i < 5; i++) cfg.move_prob[i] = move_prob[i];
    if (use_bond_rigid) cfg.bond_len = rigid_bond;                       // simulation.cc:288-296
    if (use_bond_pot) { cfg.bond_len = bond_r0; cfg.vary_bond = true; }
    cfg.gc_freq = use_gc ? gc_freq : 0;
    S.prop.configure(cfg);
    if (!S.sizer_ready) { S.sizer.reset(4096); S.sizer_ready = true; }
    S.n_mol_configured = (int)mols.size();
  }
  FILE* tf = plum_trace_file ? plum_trace_file() : NULL;
  FILE* tx = plum_trace_xyz_file ? plum_trace_xyz_file() : NULL;
  int executed = 0;
  // ---- the device-resident chain: the generator itself goes to the device, nothing stops a stretch but a GC step.
  // Offered for single-image systems (pg_chain_configure says so otherwise); PLUM_B200_CHAIN=0
  // keeps the descriptor path below, PLUM_B200_CLUSTER sets the CTAs that share the chain (default 16),
  // PLUM_B200_PIVOT_MODE = 0 / 1 / 2 the way pivot arms are built (pg_chain_config.pivot_mode, default 2).
  if (S.chain != 0) {
    static const bool chain_off = [] { const char* e = getenv("PLUM_B200_CHAIN"); return e && e[0] == '0'; }();
    if (S.chain < 0 || S.n_mol_configured < 0) {
      pg_chain_config cc;
      memset(&cc, 0, sizeof(cc));
      cc.phantom = phantom;
      cc.gc_freq = use_gc ? gc_freq : 0;
      cc.vary_bond = use_bond_pot ? 1 : 0;
      const char* ec = getenv("PLUM_B200_CLUSTER");
      cc.cluster_ctas = ec ? atoi(ec) : 16;
      const char* ep = getenv("PLUM_B200_PIVOT_MODE");
      cc.pivot_mode = ep ? atoi(ep) : 2;   // default 2: exact arms for the state, prefix-sum arms for the energies
      cc.keep_trials = tx ? 1 : 0;
      cc.move_size = move_size;
      cc.bond_len = use_bond_pot ? bond_r0 : (use_bond_rigid ? rigid_bond : 0.0);
      for (int i = 0; i < 5; i++) cc.move_prob[i] = move_prob[i];
      S.chain = (!chain_off && pg_chain_configure(engine, &cc) == PG_OK) ? 1 : 0;
    }
  }
  if (S.chain == 1) {
    uint32_t state[624];
    int pos = 0;
    S.steps.resize(max_steps);
    int rc = 0, n_done = 0, stop = 0;
    { SiteTimer st_(kTDelta);
      if (!(S.rng_resident && rand_gen == S.rng_handed)) {
        MtExport(rand_gen, state, &pos);
        rc = pg_chain_set_rng(engine, state, pos);
        if (rc) Fail("pg_chain_set_rng", rc);
      }
      tot_valid = false;
      rc = pg_chain_run(engine, max_steps, &n_done, &stop, S.steps.data(), NULL);
      if (rc) Fail("pg_chain_run", rc);
      rc = pg_chain_get_rng(engine, state, &pos);
      if (rc) Fail("pg_chain_get_rng", rc); }
    MtImport(rand_gen, state, pos);
    S.rng_handed = rand_gen;
    S.rng_resident = true;
    for (int m = 0; m < n_done; m++) {
      const pg_chain_step& d = S.steps[m];
      if (d.kind < 0) continue;
      attempted[d.kind]++;
      if (d.accept) accepted[d.kind]++;
      if (tf) {
        fprintf(tf, "T %d %d %d %a %d", first_step + m, (int)d.kind, d.mol, d.dE, (int)d.accept);
        if (m == n_done - 1) {
          fprintf(tf, " %a %a %a %a\n", use_pair_pot ? TotPairEnergy() : 0.0, use_ewald_pot ? TotEwaldEnergy() : 0.0,
                  use_bond_pot ? TotBondEnergy() : 0.0, use_ext_pot ? TotExtEnergy() : 0.0);
        } else {
          fprintf(tf, " nan nan nan nan\n");
        }
        if (tx) {
          const int len = mols[d.mol].Size();
          vector<double> t(3 * len);
          rc = pg_chain_trial_xyz(engine, m, t.data(), len);
          if (rc) Fail("pg_chain_trial_xyz", rc);
          fprintf(tx, "X %d", len);
          for (int i = 0; i < len; i++)
            fprintf(tx, " %d %a %a %a", (d.kind == PG_MOVE_BEAD) ? (i == 0) : 1, t[3 * i], t[3 * i + 1], t[3 * i + 2]);
          fprintf(tx, "\n");
        }
      }
    }
    executed = n_done;
  }
  while (S.chain != 1 && executed < max_steps) {
    const int budget = min(S.sizer.next(), max_steps - executed);
    const int n_steps = S.prop.generate(rand_gen, budget, S.batch);
    if (n_steps == 0) break;   // a GC step is next
    plum_mc::Batch& b = S.batch;
    const int n_moves = (int)b.moves.size();
    int steps_done = n_steps;
    bool stopped = false;
    if (n_moves > 0) {
      S.dE.resize(n_moves);
      S.acc.resize(n_moves);
      int n_done = 0, rc;
      { SiteTimer st_(kTDelta);
        rc = pg_mc_upload(engine, n_moves, b.moves.data(), (int)(b.rvec.size() / 4), b.rvec.data());
        if (rc) Fail("pg_mc_upload", rc);
        tot_valid = false;
        rc = pg_mc_run(engine, 0, n_moves, S.dE.data(), S.acc.data(), &n_done, NULL); }
      if (rc) Fail("pg_mc_run", rc);
      stopped = n_done > 0 && S.dE[n_done - 1] >= kVeryLargeEnergy;
      if (!stopped && n_done != n_moves) Fail("pg_mc_run (batch ended without a stop)", PG_ERR_STATE);
      for (int m = 0; m < n_done; m++) {
        const pg_move_desc& d = b.moves[m];
        attempted[d.kind]++;
        if (S.acc[m]) accepted[d.kind]++;
        if (tf) {
          // running totals are not materialised per step inside a batch: "nan" except behind the last step
          fprintf(tf, "T %d %d %d %a %d", first_step + executed + b.step_of_move[m], d.kind, d.mol, S.dE[m], (int)S.acc[m]);
          if (m == n_done - 1) {
            fprintf(tf, " %a %a %a %a\n", use_pair_pot ? TotPairEnergy() : 0.0, use_ewald_pot ? TotEwaldEnergy() : 0.0,
                    use_bond_pot ? TotBondEnergy() : 0.0, use_ext_pot ? TotExtEnergy() : 0.0);
          } else {
            fprintf(tf, " nan nan nan nan\n");
          }
          if (tx) {
            const int len = mols[d.mol].Size();
            vector<double> t(3 * len);
            rc = pg_mc_trial_xyz(engine, m, t.data());
            if (rc) Fail("pg_mc_trial_xyz", rc);
            fprintf(tx, "X %d", len);
            for (int i = 0; i < len; i++)
              fprintf(tx, " %d %a %a %a", (d.kind == PG_MOVE_BEAD) ? (i == 0) : 1, t[3 * i], t[3 * i + 1], t[3 * i + 2]);
            fprintf(tx, "\n");
          }
        }
      }
      if (stopped) steps_done = b.rewind_after_overlap(rand_gen, n_done);
      S.sizer.update(n_done, stopped);
      if (!S.sizer.worthwhile()) { S.cooldown = 20000; S.sizer.run = 128.0; }   // probe again after the cool-down
    }
    if (budget >= 24) {   // (a short budget is the caller's choice — frequent sampling — not evidence)
      S.avg_len = 0.9 * S.avg_len + 0.1 * steps_done;
      if (S.avg_len < 24.0) { S.cooldown = 20000; S.avg_len = 64.0; }
    }
    executed += steps_done;
    if (!stopped && b.stop != plum_mc::STOP_FULL) break;   // the generator stands in front of a step the driver runs
    if (S.cooldown > 0) break;
  }
  SkDriftReset(executed);
  if (executed > 0) {
    // the device's coordinates are the accepted ones: bring the driver's beads (current and trial) up to date
    int n = 0;
    for (int i = 0; i < (int)mols.size(); i++) n += mols[i].Size();
    S.xyz.resize(3 * (size_t)n);
    int rc = pg_download_positions(engine, S.xyz.data());
    if (rc) Fail("pg_download_positions", rc);
    size_t k = 0;
    for (int i = 0; i < (int)mols.size(); i++)
      for (int j = 0; j < mols[i].Size(); j++, k += 3)
        for (int a = 0; a < 3; a++) {
          mols[i].bds[j].SetCrd(0, a, S.xyz[k + a]);
          mols[i].bds[j].SetCrd(1, a, S.xyz[k + a]);
        }
  }
  return executed;
}

// ---------------------------------------------------------------------- CBMC
void ForceField::TrialEnergies(int n_trials, const double* b1, const double* b2, int current_len, int delete_id,
                               double* energy_out) {
  const bool charged = (gc_bead_charge != 0);
  pg_trial_set set;
  set.n_trials = n_trials;
  set.use_bead2 = charged ? 1 : 0;
  set.type1 = TypeId(gc_bead_symbol);
  set.type2 = set.type1;
  set.q1 = gc_chain_len ? gc_chain_chg[0] : 0.0;
  set.q2 = -set.q1;
  set.current_len = current_len;
  set.skip_mol_first = -1;
  set.skip_mol_last = -1;
  set._pad = 0;
  if (delete_id >= 0) {
    set.skip_mol_first = delete_id;
    set.skip_mol_last = delete_id + (charged ? gc_chain_len : 0);
  }
  // Partial chain: current_len monomers followed by their current_len ions.
  const int n_chain_beads = current_len * (charged ? 2 : 1);
  vector<double> cx(3 * (n_chain_beads > 0 ? n_chain_beads : 1)), cq(n_chain_beads > 0 ? n_chain_beads : 1);
  vector<int32_t> ct(n_chain_beads > 0 ? n_chain_beads : 1, set.type1);
  for (int i = 0; i < n_chain_beads; i++) {
    Bead& b = (i < current_len) ? cbmc_chain[i] : cbmc_chain[gc_chain_len + (i - current_len)];
    cx[3 * i] = b.GetCrd(1, 0);
    cx[3 * i + 1] = b.GetCrd(1, 1);
    cx[3 * i + 2] = b.GetCrd(1, 2);
    cq[i] = b.Charge();
  }
  int rc;
  { SiteTimer st_(kTTrials); rc = pg_trial_energies(engine, &set, b1, b2, cx.data(), cq.data(), ct.data(), energy_out, NULL, NULL); }
  if (rc) Fail("pg_trial_energies", rc);
}

// potential_spring.cc:50-61 (host RNG logic only; consumes the same variates).
double ForceField::RandomBondLen(mt19937& rand_gen) {
  double len = 0;
  double sigma = sqrt(1 / (beta * bond_k));
  double a = (bond_r0 + 3 * sigma) * (bond_r0 + 3 * sigma);
  bool ready = false;
  while (!ready) {
    len = gasdev(bond_r0, sigma, rand_gen);
    ready = (Uniform(rand_gen) < (len * len / a));
  }
  return len;
}

// Generates the k candidates of one growth step (all random draws first, in the reference's
// order — BeadsEnergy itself draws nothing), then evaluates them in ONE batched launch.
double ForceField::CBMCFGenTrialBeads(Bead& end_bead, vector<Molecule>& mols, int current_len, mt19937& rand_gen,
                                      int delete_id) {
  (void)mols;
  const int k = cbmc_no_of_trials;
  const bool charged = (gc_bead_charge != 0);
  vector<double> b1(3 * k), b2(3 * k, 0.0), e(k);
  for (int i = 0; i < k; i++) {
    double bond_len = 0;
    if (use_bond_pot)
      bond_len = RandomBondLen(rand_gen);
    else if (use_bond_rigid)
      bond_len = rigid_bond;
    double c[3];
    randSphere(c, rand_gen);
    for (int j = 0; j < 3; j++) {
      c[j] *= bond_len;
      c[j] += end_bead.GetCrd(0, j);
    }
    cbmc_trial_beads[i].SetAllCrd(c);
    for (int j = 0; j < 3; j++) b1[3 * i + j] = c[j];
    if (charged) {
      c[0] = Uniform(rand_gen) * box_l[0];
      c[1] = Uniform(rand_gen) * box_l[1];
      c[2] = Uniform(rand_gen) * box_l[2];
      cbmc_trial_beads[i + k].SetAllCrd(c);
      for (int j = 0; j < 3; j++) b2[3 * i + j] = c[j];
    }
  }
  TrialEnergies(k, b1.data(), b2.data(), current_len, delete_id, e.data());
  double Wi = 0;
  for (int i = 0; i < k; i++) {
    cbmc_trial_weights[i] = exp(-beta * e[i]);
    Wi += cbmc_trial_weights[i];
  }
  return Wi;
}

bool ForceField::CBMCFChainInsertion(vector<Molecule>& mols, mt19937& rand_gen) {
  SiteTimer whole_(kTGcInsert);   // the whole grand-canonical attempt, host and device (PLUM_B200_PROFILE=1)
  bool accept = false;
  double weight = 1.0;
  const bool charged = (gc_bead_charge != 0);
  const int k = cbmc_no_of_trials;
  plum_trace_weight = -1;

  // First monomer (+ its ion) uniformly in the box.
  double xyz[3];
  const int initial_beads = charged ? 2 : 1;
  for (int i = 0; i < initial_beads; i++) {
    for (int j = 0; j < 3; j++) xyz[j] = Uniform(rand_gen) * box_l[j];
    cbmc_chain[i * gc_chain_len].SetAllCrd(xyz);
  }
  {
    double b1[3], b2[3] = {0, 0, 0}, e;
    for (int j = 0; j < 3; j++) b1[j] = cbmc_chain[0].GetCrd(1, j);
    if (charged)
      for (int j = 0; j < 3; j++) b2[j] = cbmc_chain[gc_chain_len].GetCrd(1, j);
    TrialEnergies(1, b1, b2, 0, -1, &e);
    weight *= exp(-beta * e);
  }
  if (weight <= 0) return false;

  for (int i = 1; i < gc_chain_len; i++) {
    double Wi = CBMCFGenTrialBeads(cbmc_chain[i - 1], mols, i, rand_gen, -1);
    weight *= Wi / k;
    if (weight <= 0) return accept;
    // Roulette selection on the cumulative weights.
    double rand_num = Uniform(rand_gen) * Wi;
    int current_bead = 0;
    double cumulate_weight = cbmc_trial_weights[0];
    while (cumulate_weight < rand_num && current_bead < k - 1) {
      current_bead++;
      cumulate_weight += cbmc_trial_weights[current_bead];
    }
    for (int j = 0; j < 3; j++) xyz[j] = cbmc_trial_beads[current_bead].GetCrd(0, j);
    cbmc_chain[i].SetAllCrd(xyz);
    if (charged) {
      for (int j = 0; j < 3; j++) xyz[j] = cbmc_trial_beads[current_bead + k].GetCrd(0, j);
      cbmc_chain[i + gc_chain_len].SetAllCrd(xyz);
    }
  }

  // Acceptance rule, cbmc.cc:273-300.
  UpdateMolCounts(mols);
  int spc1;
  if (gc_chain_len > 1) spc1 = n_chain;
  else if (gc_bead_charge >= 0) spc1 = n_cion - coion;
  else spc1 = n_aion - coion;
  int spc2;
  if (gc_bead_charge == 0) spc2 = 0;
  else if (gc_bead_charge > 0) spc2 = n_aion;
  else spc2 = n_cion;
  int spc1_add = 1;
  int spc2_add = charged ? gc_chain_len : 0;
  double factorial = 1;
  double m1 = gc_chain_len, m2 = 1;
  double vol_over_lam = 1;
  for (int i = spc1 + 1; i <= spc1 + spc1_add; i++) factorial *= i;
  for (int i = spc2 + 1; i <= spc2 + spc2_add; i++) factorial *= i;
  vol_over_lam *= pow(vol / (gc_deBroglie_prefactor / pow(m1, 1.5)), spc1_add);
  vol_over_lam *= pow(vol / (gc_deBroglie_prefactor / pow(m2, 1.5)), spc2_add);
  double C = vol_over_lam * (1.0 / (double)factorial);
  double rand_num = Uniform(rand_gen);
  plum_trace_weight = weight;
  if (rand_num < (exp(beta * chem_pot) * weight) * C) {
    accept = true;
    mols.push_back(Molecule());
    for (int i = 0; i < gc_chain_len; i++) mols[(int)mols.size() - 1].AddBead(cbmc_chain[i]);
    for (int i = 0; i < mols[(int)mols.size() - 1].Size() - 1; i++) mols[(int)mols.size() - 1].AddBond(i, i + 1);
    if (charged) {
      for (int i = gc_chain_len; i < gc_chain_len * 2; i++) {
        mols.push_back(Molecule());
        mols[(int)mols.size() - 1].AddBead(cbmc_chain[i]);
      }
    }
  }
  return accept;
}

void ForceField::EnergyInitForAddedMolecule(vector<Molecule>& mols) {
  int added = 1;
  if (gc_bead_charge != 0) added += gc_chain_len;
  vector<int32_t> lens, type;
  vector<double> xyz, q;
  for (int i = (int)mols.size() - added; i < (int)mols.size(); i++) {
    lens.push_back(mols[i].Size());
    for (int j = 0; j < mols[i].Size(); j++) {
      Bead& b = mols[i].bds[j];
      for (int a = 0; a < 3; a++) xyz.push_back(b.GetCrd(0, a));
      q.push_back(b.Charge());
      type.push_back(TypeId(b.Symbol()));
    }
  }
  int rc;
  tot_valid = false;
  { SiteTimer st_(kTInsert); rc = pg_insert_molecules(engine, added, lens.data(), xyz.data(), q.data(), type.data(), NULL); }
  if (rc) Fail("pg_insert_molecules", rc);
}

int ForceField::CBMCFChainDeletion(vector<Molecule>& mols, mt19937& rand_gen) {
  SiteTimer whole_(kTGcDelete);
  UpdateMolCounts(mols);
  const bool charged = (gc_bead_charge != 0);
  const int k = cbmc_no_of_trials;
  plum_trace_weight = -1;
  int spc1;
  if (gc_chain_len > 1) spc1 = n_chain;
  else if (gc_bead_charge >= 0) spc1 = n_cion - coion;
  else spc1 = n_aion - coion;

  int delete_id = floor(Uniform(rand_gen) * spc1);
  if (delete_id >= spc1) delete_id = spc1 - 1;
  if (charged)
    delete_id = phantom + coion + grafted + grafted_counterion + delete_id * (1 + gc_chain_len);
  else
    delete_id = phantom + coion + grafted + grafted_counterion + delete_id;
  if (delete_id < 0 || delete_id >= (int)mols.size()) return -1;   // nothing of the GC species to delete

  // Retrace the chain: first monomer (+ion) against everything but the chain itself.
  double weight = 1.0;
  const int second_mol = charged ? delete_id + 1 : delete_id;
  double b1[3], b2[3] = {0, 0, 0}, e;
  for (int j = 0; j < 3; j++) b1[j] = mols[delete_id].bds[0].GetCrd(0, j);
  if (charged)
    for (int j = 0; j < 3; j++) b2[j] = mols[second_mol].bds[0].GetCrd(0, j);
  TrialEnergies(1, b1, b2, 0, delete_id, &e);
  weight *= exp(-beta * e);
  cbmc_chain[0].SetAllCrd(b1);
  if (charged) cbmc_chain[gc_chain_len].SetAllCrd(b2);

  for (int i = 1; i < gc_chain_len; i++) {
    double xyz[3];
    for (int j = 0; j < 3; j++) xyz[j] = mols[delete_id].bds[i].GetCrd(0, j);
    cbmc_chain[i].SetAllCrd(xyz);
    if (charged) {
      for (int j = 0; j < 3; j++) xyz[j] = mols[delete_id + i + 1].bds[0].GetCrd(0, j);
      cbmc_chain[i + gc_chain_len].SetAllCrd(xyz);
    }
    // k fresh candidates (their draws are consumed), slot 0 replaced by the real bead.
    CBMCFGenTrialBeads(cbmc_chain[i - 1], mols, i, rand_gen, delete_id);
    for (int j = 0; j < 3; j++) b1[j] = cbmc_chain[i].GetCrd(1, j);
    if (charged)
      for (int j = 0; j < 3; j++) b2[j] = cbmc_chain[i + gc_chain_len].GetCrd(1, j);
    TrialEnergies(1, b1, b2, i, delete_id, &e);
    cbmc_trial_weights[0] = exp(-beta * e);
    double Wi = 0;
    for (int j = 0; j < k; j++) Wi += cbmc_trial_weights[j];
    weight *= Wi / k;
  }

  int spc2;
  if (gc_bead_charge == 0) spc2 = 0;
  else if (gc_bead_charge > 0) spc2 = n_aion;
  else spc2 = n_cion;
  int spc1_del = 1;
  int spc2_del = charged ? gc_chain_len : 0;
  double factorial = 1;
  double m1 = gc_chain_len, m2 = 1;
  double lam_over_vol = 1;
  for (int i = spc1; i > spc1 - spc1_del; i--) factorial *= i;
  for (int i = spc2; i > spc2 - spc2_del; i--) factorial *= i;
  lam_over_vol *= pow((gc_deBroglie_prefactor / pow(m1, 1.5)) / vol, spc1_del);
  lam_over_vol *= pow((gc_deBroglie_prefactor / pow(m2, 1.5)) / vol, spc2_del);
  double C = lam_over_vol * (double)factorial;
  double rand_num = Uniform(rand_gen);
  plum_trace_weight = weight;
  if (rand_num < C / (exp(beta * chem_pot) * weight)) {
    // Monovalent counter-ions, as the reference assumes (potential_pair.cc:251-255).
    int counterion = 0;
    for (int i = 0; i < mols[delete_id].Size(); i++) counterion += (int)abs(round(mols[delete_id].bds[i].Charge()));
    int rc;
    tot_valid = false;
    { SiteTimer st_(kTDelete); rc = pg_delete_molecules(engine, delete_id, delete_id + counterion, NULL); }
    if (rc) Fail("pg_delete_molecules", rc);
    return delete_id;
  }
  return -1;
}

// Widom-style chemical potential by ghost CBMC insertions (cbmc.cc:444-531): the same growth as
// CBMCFChainInsertion without the acceptance step, mu_tot_ins times.  It draws random numbers, so it
// is on the trajectory whenever s1_calc_chem_pot is 1.
double ForceField::CalcChemicalPotentialF(vector<Molecule>& mols, mt19937& rand_gen) {
  double total = 0;
  const bool charged = (gc_bead_charge != 0);
  const int k = cbmc_no_of_trials;

  int spc1;
  if (gc_chain_len > 1) spc1 = n_chain;
  else if (gc_bead_charge >= 0) spc1 = n_cion - coion;
  else spc1 = n_aion - coion;
  int spc2;
  if (gc_bead_charge == 0) spc2 = 0;
  else if (gc_bead_charge > 0) spc2 = n_aion;
  else spc2 = n_cion;
  int spc1_add = 1;
  int spc2_add = charged ? gc_chain_len : 0;
  double factorial = 1;
  double m1 = gc_chain_len, m2 = 1;
  double vol_over_lam = 1;
  for (int i = spc1 + 1; i <= spc1 + spc1_add; i++) factorial *= i;
  for (int i = spc2 + 1; i <= spc2 + spc2_add; i++) factorial *= i;
  vol_over_lam *= pow(vol / 15.625 / (gc_deBroglie_prefactor / pow(m1, 1.5)), spc1_add);
  vol_over_lam *= pow(vol / 15.625 / (gc_deBroglie_prefactor / pow(m2, 1.5)), spc2_add);
  double C = vol_over_lam * (1.0 / (double)factorial);

  for (int c = 0; c < mu_tot_ins; c++) {
    double weight = 1.0;
    double xyz[3];
    const int initial_beads = charged ? 2 : 1;
    for (int i = 0; i < initial_beads; i++) {
      for (int j = 0; j < 3; j++) xyz[j] = Uniform(rand_gen) * box_l[j];
      cbmc_chain[i * gc_chain_len].SetAllCrd(xyz);
    }
    double b1[3], b2[3] = {0, 0, 0}, dE;
    for (int j = 0; j < 3; j++) b1[j] = cbmc_chain[0].GetCrd(1, j);
    if (charged)
      for (int j = 0; j < 3; j++) b2[j] = cbmc_chain[gc_chain_len].GetCrd(1, j);
    TrialEnergies(1, b1, b2, 0, -1, &dE);
    if (dE == kVeryLargeEnergy) weight *= 0;
    else weight *= exp(-beta * dE);

    for (int i = 1; i < gc_chain_len; i++) {
      double Wi = CBMCFGenTrialBeads(cbmc_chain[i - 1], mols, i, rand_gen, -1);
      weight *= Wi / k;
      if (weight <= 0) break;
      double rand_num = Uniform(rand_gen) * Wi;
      int current_bead = 0;
      double cumulate_weight = cbmc_trial_weights[0];
      while (cumulate_weight < rand_num && current_bead < k - 1) {
        current_bead++;
        cumulate_weight += cbmc_trial_weights[current_bead];
      }
      for (int j = 0; j < 3; j++) xyz[j] = cbmc_trial_beads[current_bead].GetCrd(0, j);
      cbmc_chain[i].SetAllCrd(xyz);
      if (charged) {
        for (int j = 0; j < 3; j++) xyz[j] = cbmc_trial_beads[current_bead + k].GetCrd(0, j);
        cbmc_chain[i + gc_chain_len].SetAllCrd(xyz);
      }
    }
    total += weight * C;
  }
  return total / (double)mu_tot_ins;
}

// ------------------------------------------------------------------ samplers
// Bulk volume-perturbation sampler, src/force_field/pressure.cc:187-387.  The energy change of the virtual
// stretch (all pair loops, wall, bond and dipole terms, pressure.cc:208-338) is one pg_vol_scaling_sample call on
// the resident configuration; the accumulation and the Yethiraj / de Miguel averages below are the reference's
// (pressure.cc:340-384).  The reference never prints them for systems without walls (GetPressure returns "nan",
// pressure.cc:490-500); with PLUM_TRACE set both binaries write them as a "V" line (driver_hooks.py).
void ForceField::CalcPressureVolScalingHSELSlit(vector<Molecule>& mols) {
  (void)mols;
  vp_z++;
  const double dz = kDz;
  pg_vol_sample smp;
  int rc;
  { SiteTimer st_(kTVolScale); rc = pg_vol_scaling_sample(engine, phantom, dz, &smp); }
  if (rc) Fail("pg_vol_scaling_sample", rc);
  for (int i = 0; i < 16; i++) { p_tensor_el[i] += smp.el[i]; p_tensor_hs[i] += smp.hs[i]; }
  p_tensor[5] += smp.bond;
  p_tensor[4] += smp.dipole;
  const double dU = smp.dU;
  const double vol = box_l[0] * box_l[1] * box_l[2];
  p_tensor[2] += (n_mol - phantom) / vol;
  p_tensor[3] += pow(1.0 + dz / box_l[2], n_mol - phantom) * exp(-beta * dU);
  p_tensor[0] = 0;
  for (int i = 0; i < 4; i++)
    for (int j = i; j < 4; j++) p_tensor[0] += p_tensor_el[i * 4 + j] + p_tensor_hs[i * 4 + j];
  p_tensor[0] += p_tensor[4] + p_tensor[5];
  p_tensor[0] /= (box_l[0] * box_l[1] * dz);
  p_tensor[0] = (beta * p_tensor[2] - p_tensor[0]) / vp_z;
  for (int i = 0; i < 4; i++)
    for (int j = i; j < 4; j++) {
      const int index = i * 4 + j;
      p_tensor2[index] = -p_tensor_hs[index] / (box_l[0] * box_l[1] * dz * vp_z);
      p_tensor3[index] = -p_tensor_el[index] / (box_l[0] * box_l[1] * dz * vp_z);
    }
  p_tensor2[17] = -p_tensor[5] / (box_l[0] * box_l[1] * dz * vp_z);
  p_tensor2[18] = -p_tensor[4] / (box_l[0] * box_l[1] * dz * vp_z);
  p_tensor2[19] = beta * p_tensor[2] / vp_z;
  p_tensor[1] = beta / (box_l[0] * box_l[1] * dz) * log(p_tensor[3] / vp_z);
  if (plum_trace_file && plum_trace_file()) {
    FILE* tf = plum_trace_file();
    fprintf(tf, "V %d", vp_z);
    for (int i = 0; i < 6; i++) fprintf(tf, " %a", p_tensor[i]);
    for (int i = 0; i < 16; i++) fprintf(tf, " %a", p_tensor_el[i]);
    for (int i = 0; i < 16; i++) fprintf(tf, " %a", p_tensor_hs[i]);
    fprintf(tf, "\n");
  }
}
// Slab wall-force pressure, src/force_field/pressure.cc:404-484.  The six force sums of the current
// configuration come from one pg_wall_force launch (site-site LJ + Ewald real/reciprocal z-forces between
// the wall sites and everything else, plus the plate LJ force); the running averages below are the
// reference's.  The reference never initialises vp_z on this branch (force_field.cc:293-322 sets it only
// for bulk systems); it starts at zero here.
void ForceField::CalcPressureForceLJELSlit(vector<Molecule>& mols) {
  (void)mols;
  vp_z++;
  double f[6];
  int rc;
  { SiteTimer st_(kTWallForce); rc = pg_wall_force(engine, phantom, f); }
  if (rc) Fail("pg_wall_force", rc);
  p_tensor[6] += f[0];
  p_tensor[7] += f[1];
  p_tensor[9] += f[3];
  p_tensor[10] += f[4];
  if (vp_z == 1) {   // the walls are fixed: wall-wall terms once
    p_tensor[8] += f[2];
    p_tensor[11] += f[5];
    p_tensor[2] = p_tensor[8] / (box_l[0] * box_l[1]);
    p_tensor[5] = p_tensor[11] / (box_l[0] * box_l[1]);
  }
  p_tensor[0] = p_tensor[6] / (vp_z * box_l[0] * box_l[1]);
  p_tensor[1] = p_tensor[7] / (vp_z * box_l[0] * box_l[1]);
  p_tensor[3] = p_tensor[9] / (vp_z * box_l[0] * box_l[1]);
  p_tensor[4] = p_tensor[10] / (vp_z * box_l[0] * box_l[1]);
}

// pressure.cc:490-500
string ForceField::GetPressure() {
  std::ostringstream foo;
  if (use_ext_pot)
    foo << p_tensor[0] << " " << p_tensor[1] << " " << p_tensor[2] << " " << p_tensor[3] << " " << p_tensor[4] << " "
        << p_tensor[5];
  else
    foo << "nan";
  return foo.str();
}

// ----------------------------------------------------------------- utilities
int ForceField::GCFrequency() { return gc_freq; }
bool ForceField::UseGC() { return use_gc; }
bool ForceField::UsePairPot() { return use_pair_pot; }
bool ForceField::UseEwaldPot() { return use_ewald_pot; }
bool ForceField::UseBondPot() { return use_bond_pot; }
bool ForceField::UseBondRigid() { return use_bond_rigid; }
bool ForceField::UseAnglePot() { return use_angle_pot; }
bool ForceField::UseDihedPot() { return use_dihed_pot; }
bool ForceField::UseExtPot() { return use_ext_pot; }

// The four running totals come from one pg_get_totals per configuration: the driver asks for them one by one
// (PrintStat, the trace hooks), the cache is dropped by every call that changes the resident state.
const double* ForceField::Totals() {
  if (!tot_valid) {
    pg_totals t;
    int rc;
    { SiteTimer st_(kTTotals); rc = pg_get_totals(engine, &t); }
    if (rc) Fail("pg_get_totals", rc);
    tot_cache[0] = t.pair; tot_cache[1] = t.ewald; tot_cache[2] = t.bond; tot_cache[3] = t.ext;
    tot_valid = true;
  }
  return tot_cache;
}
double ForceField::TotPairEnergy() { return Totals()[0]; }
double ForceField::TotEwaldEnergy() { return Totals()[1]; }
double ForceField::TotBondEnergy() { return Totals()[2]; }
double ForceField::TotExtEnergy() { return Totals()[3]; }

double ForceField::RigidBondLen() { return rigid_bond; }

void ForceField::SetBoxLen(double box_l_in[3]) {
  box_l[0] = box_l_in[0];
  box_l[1] = box_l_in[1];
  box_l[2] = box_l_in[2];
}

void ForceField::UpdateMolCounts(vector<Molecule>& mols) {
  if (mc_state) static_cast<McState*>(mc_state)->n_mol_configured = -1;   // the molecule list may have changed
  n_mol = (int)mols.size();
  n_chain = 0;
  n_cion = 0;
  n_aion = 0;
  for (int i = phantom; i < n_mol; i++) {
    if (mols[i].Size() > 1) n_chain++;
    else if (mols[i].bds[0].Charge() >= 0) n_cion++;
    else if (mols[i].bds[0].Charge() < 0) n_aion++;
  }
  n_chain -= grafted;
}

double ForceField::EqBondLen() {
  if (use_bond_pot) return bond_r0;
  cout << "  Illegal request of bond length! Exiting. Program complete." << endl;
  exit(1);
  return 0;
}
