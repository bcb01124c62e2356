/** B200 façade with the public interface of Plum's `ForceField`.
 *
 *  Drop-in replacement for src/force_field/force_field.h:22-294 of the reference:
 *  the MC driver (src/simulation/simulation.cc) is compiled UNCHANGED against this
 *  header and finds the same class name and the same public member functions with
 *  the same argument meaning.  Every energy is evaluated by the sm_100a kernels
 *  behind the C ABI in include/plum_b200.h; this class only
 *    - consumes the same run.in tokens in the same order (order is the schema),
 *    - keeps the host-side control flow that owns the random stream (CBMC growth,
 *      roulette selection, acceptance tests), and
 *    - forwards coordinates / decisions across the ABI.
 *  There is no CPU fallback: a failing ABI call prints the error and exit(1)s,
 *  like every error path of the reference does.
 *
 *  Build: plum_b200/host/build_host.py places this file at
 *  <tmp>/src/force_field/force_field.h next to the untouched driver sources.
 */
#ifndef SRC_FORCE_FIELD_FORCE_FIELD_H_
#define SRC_FORCE_FIELD_FORCE_FIELD_H_

// The driver relies on what the reference's force_field.h pulls in transitively
// (potential_pair.h: <iostream>, <iomanip>, <numeric>, <map>, ../utilities/misc.h).
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "../molecules/bead.h"
#include "../molecules/molecule.h"
#include "../utilities/misc.h"

struct pg_engine;

using namespace std;

class ForceField {
 private:
  // Same basic parameters as the reference (force_field.h:24-66).
  double beta;
  int npbc;
  double box_l[3];
  double p_tensor[12];   // wall-force pressure: [0..5] running averages, [6..11] cumulative sums (pressure.cc:389-403);
                         // volume-perturbation pressure (no walls): the six entries of pressure.cc:176-186
  double p_tensor2[20], p_tensor3[20], p_tensor_hs[20], p_tensor_el[20];   // per-species tables of pressure.cc:184-185, :366-377
  int vp_z;              // pressure samples taken
  int n_mol, phantom, coion, grafted, grafted_counterion;
  int chain_len, n_chain, n_cion, n_aion;
  bool use_gc, use_pair_pot, use_ewald_pot, use_bond_pot, use_bond_rigid;
  bool use_angle_pot, use_dihed_pot, use_ext_pot;
  string gc_bead_symbol;
  int gc_chain_len;
  int gc_bead_charge;
  vector<double> gc_chain_chg;
  int gc_freq;
  double chem_pot;
  double gc_deBroglie_prefactor;
  int cbmc_no_of_trials;
  vector<double> cbmc_trial_weights;
  vector<Bead> cbmc_trial_beads;
  vector<Bead> cbmc_chain;
  int mu_tot_ins;
  double rigid_bond;
  double vol;
  double vp_el_res, vp_bead_size;
  // Potential parameters read from stdin (what the reference's potential objects hold).
  string pair_name, ewald_name, bond_name, ext_name;
  double lj_cutoff;
  map<string, double> lj_sigmas, lj_epsilons, hs_radii;
  double lB, dielectric, alpha;
  bool dipole_correction;
  double bond_k, bond_r0;
  double wall_cut, wall_sig, wall_eps, well_width, well_depth;
  map<string, double> wall_sigmas, wall_epsilons;
  // Device side.
  pg_engine* engine;
  bool tot_valid;        // tot_cache holds the engine's running totals (one pg_get_totals per step at most)
  double tot_cache[4];   // pair, ewald, bond, ext
  const double* Totals();
  map<string, int> type_of;        // bead symbol -> dense type id
  vector<string> type_symbols;
  int pending_mol;                  // molecule of the trial awaiting FinalizeEnergies (-1: none)
  long long sk_steps = 0;           // steps since the last full S(k) recompute (SkDriftReset)
  void* mc_state;                   // batched translational steps: generator, batch buffers (force_field.cc)

  int TypeId(const string& symbol);
  void Fail(const char* what, int rc);
  void ReadPotentialParameters(vector<Molecule>& mols);
  void CreateEngine(vector<Molecule>& mols);
  void UploadSystem(vector<Molecule>& mols);
  /** ForceField::BeadsEnergy (cbmc.cc:5-151) for a batch of (monomer, ion) candidates. */
  void TrialEnergies(int n_trials, const double* b1, const double* b2, int current_len, int delete_id,
                     double* energy_out);
  double RandomBondLen(mt19937& rand_gen);

 public:
  ForceField();
  ~ForceField();
  void Initialize(double, int, double[3], vector<Molecule>&, int, int, int, int);
  void InitializeEnergy(vector<Molecule>&);

  double EnergyDifference(vector<Molecule>&, int);
  void FinalizeEnergies(vector<Molecule>&, bool, int);

  /** NEW (no reference counterpart): runs up to `max_steps` consecutive steps of
   *  Simulation::Run (simulation.cc:216-237) whose move is translational — proposal, energy change,
   *  Metropolis test and commit on the device (pg_mc_*), the random stream drawn ahead on the host in the
   *  reference's order — and returns how many steps were executed.  0 means the next step is a GC step or
   *  uses a move the device does not offer (crankshaft): the driver's own code runs that one.  On return
   *  `mols` holds the accepted coordinates, `rand_gen` stands where the reference's generator would, and
   *  attempted[] / accepted[] (simulation.h:110-111) are updated.  `first_step` only labels trace lines. */
  int TranslationalBatch(vector<Molecule>& mols, mt19937& rand_gen, int max_steps, int first_step, double move_size,
                         const double move_prob[5], int attempted[], int accepted[]);
  /** PLUM_B200_BATCH=0 turns the batched steps off (the driver then runs every step itself). */
  bool BatchedMoves();
  /** Periodic full recompute of the structure factor (drift reset); called with the steps just executed. */
  void SkDriftReset(int steps_done);

  // Pressure samplers: SURVEY.md §8(f) "next" #1 — not on the per-move path.  They
  // consume no random numbers, so leaving them out does not change the trajectory.
  void CalcPressureVolScalingHSELSlit(vector<Molecule>&);
  void CalcPressureForceLJELSlit(vector<Molecule>&);
  string GetPressure();

  int GCFrequency();
  double CBMCFGenTrialBeads(Bead&, vector<Molecule>&, int, mt19937&, int);
  bool CBMCFChainInsertion(vector<Molecule>&, mt19937&);
  int CBMCFChainDeletion(vector<Molecule>&, mt19937&);
  double CalcChemicalPotentialF(vector<Molecule>&, mt19937&);
  void EnergyInitForAddedMolecule(vector<Molecule>&);

  bool UseGC();
  bool UsePairPot();
  bool UseEwaldPot();
  bool UseBondPot();
  bool UseBondRigid();
  bool UseAnglePot();
  bool UseDihedPot();
  bool UseExtPot();
  double TotPairEnergy();
  double TotEwaldEnergy();
  double TotBondEnergy();
  double TotExtEnergy();
  double RigidBondLen();
  void SetBoxLen(double[]);
  void UpdateMolCounts(vector<Molecule>&);
  double EqBondLen();
};

#endif
