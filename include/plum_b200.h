/* plum_b200 — C ABI of the B200-native energy engine for Plum's per-trial-move
 * energy path.
 *
 * This header is the drop-in boundary.  Everything above it (the MC driver,
 * run.in parsing, move generators, Metropolis test) is Plum's own C++ and stays
 * as it is; everything below it is hand-written sm_100a CUDA.  The host façade
 * `plum_b200/host/force_field.{h,cc}` has the public signature of the
 * reference's `class ForceField` (src/force_field/force_field.h:207-294) and
 * forwards to these entry points; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain C types only: pointers and sizes, no C++/torch types;
 *  - every entry point returns 0 on success, a negative pg_status on failure;
 *    `pg_last_error(h)` gives a message (CUDA error strings included);
 *  - all pointers are HOST pointers owned by the caller unless the name ends
 *    in `_dev`;
 *  - beads are indexed densely in molecule order (the iteration order of every
 *    reference loop, e.g. src/force_field/potential_pair.cc:181-197); a
 *    molecule is a contiguous bead range;
 *  - all arithmetic is FP64.  Energies are in kBT, lengths in "ul".
 *  - one handle = one system replica on one GPU; a handle is not thread safe.
 *  - there is no CPU fallback: if no CUDA device is usable pg_create fails.
 */
#ifndef PLUM_B200_H_
#define PLUM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 1

/* The reference's sentinel for "infinite" energy (src/utilities/constants.h:16). */
#define PG_VERY_LARGE_ENERGY 1.0e8

typedef enum pg_status {
  PG_OK = 0,
  PG_ERR_INVALID = -1,   /* bad argument */
  PG_ERR_CUDA = -2,      /* CUDA runtime error, see pg_last_error */
  PG_ERR_NO_DEVICE = -3, /* no usable sm_100 device: there is no CPU fallback */
  PG_ERR_STATE = -4,     /* call out of order (e.g. commit without a trial) */
  PG_ERR_CAPACITY = -5,  /* exceeds a compiled-in or allocated capacity */
  PG_ERR_TIMEOUT = -6    /* device did not answer within the watchdog time */
} pg_status;

/* pair_kind: which PotentialPair subclass (src/force_field/force_field.cc:123-139) */
enum { PG_PAIR_NONE = 0, PG_PAIR_TRUNCATED_LJ = 1, PG_PAIR_HARD_SPHERE = 2 };
/* bond_kind (src/force_field/force_field.cc:154-165) */
enum { PG_BOND_NONE = 0, PG_BOND_SPRING = 1 };
/* ext_kind (src/force_field/force_field.cc:174-195) */
enum { PG_EXT_NONE = 0, PG_EXT_TRUNCATED_LJ_WALL = 1, PG_EXT_HARD_WALL = 2, PG_EXT_WELL_WALL = 3 };
/* graft_kind per bead type: symbol "L" / "R" select the tethered branch of the
 * LJ wall (src/force_field/potential_truncated_lj_wall.cc:82-115). */
enum { PG_GRAFT_NONE = 0, PG_GRAFT_LEFT = 1, PG_GRAFT_RIGHT = 2 };

/* Force-field parameters.  Bead "symbols" (strings in the reference) are mapped
 * by the caller to dense int type ids 0..n_types-1; a symbol missing from a
 * reference std::map reads as 0 there (operator[]), so absent entries are 0 here. */
typedef struct pg_params {
  double box[3];            /* ForceField::box_l, the unpadded box             */
  int32_t npbc;             /* 2 or 3 (src/simulation/simulation.cc:152)       */
  int32_t n_types;
  double beta;

  int32_t pair_kind;        /* PG_PAIR_*                                       */
  int32_t _pad0;
  double lj_cutoff;         /* <0: WCA (potential_truncated_lj.cc:69-76)       */
  const double* lj_sigma;   /* [n_types]                                       */
  const double* lj_epsilon; /* [n_types]                                       */
  const double* hs_radius;  /* [n_types] (potential_hard_sphere.cc:38-49)      */

  int32_t use_ewald;        /* PotentialEwaldCoul                              */
  int32_t dipole_correction;
  double lB;                /* Bjerrum length                                  */
  double alpha;             /* Ewald alpha, ul^-2 (it is kappa^2)              */

  int32_t bond_kind;        /* PG_BOND_*                                       */
  int32_t _pad1;
  double bond_k;            /* potential_spring.cc:15-20                       */
  double bond_r0;

  int32_t ext_kind;         /* PG_EXT_*                                        */
  int32_t _pad2;
  double wall_cut;          /* m_cut (potential_truncated_lj_wall.cc:23)       */
  const double* wall_sigma;   /* [n_types] LJ wall sigma, or radius for hard/well walls */
  const double* wall_epsilon; /* [n_types]                                     */
  const int32_t* graft_kind;  /* [n_types] PG_GRAFT_*                          */
  double well_width;        /* potential_well_wall.cc:41-55                    */
  double well_depth;
} pg_params;

/* Derived Ewald set-up (src/force_field/potential_ewald_coul.cc:29-132). */
typedef struct pg_ewald_info {
  double ewald_box[3];   /* padded in z when dipole correction is on           */
  double box_vol;
  double real_cutoff;
  int32_t real_cell[3];
  int32_t n_k;           /* k vectors with 0 < k2 <= repl_cutoff in the cube   */
  double repl_cutoff;
  int32_t repl_cell[3];
  int32_t n_k_half;      /* the half-space actually stored on the device       */
} pg_ewald_info;

/* Components of one trial move's energy change, and the orchestrated total the
 * driver compares with kVeryLargeEnergy (src/force_field/force_field.cc:407-434). */
typedef struct pg_delta {
  double dE;          /* what ForceField::EnergyDifference returns               */
  double pair;        /* PotentialPair::EnergyDifference                         */
  double ext;         /* PotentialExternal::EnergyDifference (exactly 1e8 if out)*/
  double ewald;       /* real + recip (+ dipole term supplied by the caller)     */
  double bond;        /* PotentialBond::EnergyDifference                         */
  double real;        /* Ewald real-space part                                   */
  double recip;       /* Ewald reciprocal part via dS(k)                         */
  double mz_current;  /* sum q*z over all beads, current coordinates             */
  int32_t stage;      /* 0 full; 1 returned after pair; 2 returned after ext     */
  int32_t n_overlap;  /* pairs that hit the r<=0 / hard-core sentinel            */
} pg_delta;

/* Running totals as the reference accumulates them (E_tot of each potential). */
typedef struct pg_totals {
  double pair, ewald, bond, ext;
  double real, recip, self, dipole;   /* Ewald split; dipole = current_dipl_E   */
} pg_totals;

typedef struct pg_engine pg_engine;

/* ---- life cycle --------------------------------------------------------- */
/* Replaces ForceField::Initialize's potential construction
 * (src/force_field/force_field.cc:118-195).  `device` is a CUDA ordinal.
 * `capacity_beads` sizes the resident arrays (>= n ever present; GC grows). */
int pg_create(const pg_params* params, int device, int capacity_beads, pg_engine** out);
int pg_destroy(pg_engine* h);
const char* pg_last_error(const pg_engine* h);
int pg_abi_version(void);
int pg_get_ewald_info(const pg_engine* h, pg_ewald_info* out);

/* ---- state --------------------------------------------------------------- */
/* Uploads the whole system (replaces the map fills of
 * PotentialPair::EnergyInitialization potential_pair.cc:55-100 and
 * PotentialEwald::EnergyInitialization potential_ewald.cc:176-230):
 * xyz is [n][3]; mol_first has n_mol+1 entries (bead range of each molecule). */
int pg_upload_system(pg_engine* h, int n_beads, const double* xyz, const double* q,
                     const int32_t* type, int n_mol, const int32_t* mol_first);
/* Full recompute of every total from resident coordinates, including the full
 * S(k) (ForceField::InitializeEnergy force_field.cc:365-381).  Resets the
 * running totals to the recomputed values. */
int pg_init_energy(pg_engine* h, pg_totals* out);
/* Recompute totals WITHOUT touching the running totals or S(k) (drift check). */
int pg_recompute_totals(pg_engine* h, pg_totals* out);
int pg_get_totals(const pg_engine* h, pg_totals* out);
int pg_download_positions(pg_engine* h, double* xyz /* [n][3] */);
int pg_num_beads(const pg_engine* h);

/* ---- the per-move hot path ------------------------------------------------ */
/* ForceField::EnergyDifference (force_field.cc:407-434).  `mol` is the moved
 * molecule; trial_xyz is [len][3] for ALL its beads (trial == current for the
 * unmoved ones); moved is [len] 0/1 (Bead::GetMoved).  Synchronous: returns when
 * the result is on the host.  `dipole_lag` is added to the Ewald term: the
 * reference's trial_dipl_E - current_dipl_E quirk is a host-side state machine
 * (potential_ewald.cc:527-531) fed from out->mz_current. */
int pg_delta_e(pg_engine* h, int mol, const double* trial_xyz, const uint8_t* moved,
               pg_delta* out);
/* The same call in two halves, for a host thread that drives several replicas
 * (one handle each) and overlaps their round trips: _begin stages the trial and
 * launches, _poll returns 1 while the device has not answered, 0 once `out` is
 * filled (the trial then awaits pg_commit), < 0 on error.  The trial_xyz / moved
 * buffers may be reused as soon as _begin returns. */
int pg_delta_e_begin(pg_engine* h, int mol, const double* trial_xyz, const uint8_t* moved);
int pg_delta_e_poll(pg_engine* h, pg_delta* out);
/* ForceField::FinalizeEnergies (force_field.cc:436-451): accept applies the
 * position deltas and S(k) += dS(k) on the device and adds the components to
 * the running totals; reject drops the trial.  Asynchronous on the device. */
int pg_commit(pg_engine* h, int accept);

/* ---- device-resident replay (bench `value` leg; inputs already in HBM) ---- */
/* A packed list of proposals is uploaded once; pg_replay runs them back to
 * back with the Metropolis decision taken on the device from the recorded
 * uniform variate (u < exp(-beta dE), skipped when dE >= 1e8 exactly like
 * src/simulation/simulation.cc:327-332).  No host round trip per move. */
typedef struct pg_proposal {
  int32_t mol;        /* moved molecule                                        */
  int32_t xyz_offset; /* offset (in beads) of its trial coordinates in xyz     */
  double u;           /* uniform variate for the acceptance test               */
} pg_proposal;
int pg_replay_upload(pg_engine* h, int n_moves, const pg_proposal* moves,
                     int n_xyz_beads, const double* trial_xyz, const uint8_t* moved);
/* Runs moves [first, first+count); dE_out/accept_out (host, may be NULL) are
 * filled after the batch.  elapsed_ms is CUDA-event time on the engine stream. */
int pg_replay_run(pg_engine* h, int first, int count, double* dE_out, uint8_t* accept_out,
                  float* elapsed_ms);
/* Same proposals, energy kernel only (no commit: the state does not advance): the
 * CUDA-event time of `count` back-to-back launches of the dominant kernel, for the
 * roofline figure. */
int pg_replay_time_delta(pg_engine* h, int first, int count, float* elapsed_ms);
/* Builds (captures + instantiates) the CUDA graph of a replay batch ahead of time so that the
 * first pg_replay_run / pg_replay_time_delta (with_commit 1 / 0) of that batch does not pay for it. */
int pg_replay_prepare(pg_engine* h, int first, int count, int with_commit);

/* ---- device-side proposals + Metropolis test (batched Markov-chain steps) -- */
/* Replaces the per-move round trip of Simulation::TranslationalMove
 * (src/simulation/simulation.cc:241-355) for a run of consecutive translational
 * steps.  The caller walks its own std::mt19937 ahead (the draw ORDER of
 * simulation.cc:247-331, molecule.cc:103-312 and misc.cc:95-109 is the caller's
 * business: plum_b200/host/mc_propose.h does it) and hands over, per step, only
 * what does not depend on the coordinates: the chosen molecule, the move kind and
 * the random vectors.  The device builds the trial coordinates from the RESIDENT
 * coordinates with the reference's exact operation order (round-to-nearest mul /
 * add / div / sqrt, never fused: Molecule::BeadTranslate molecule.cc:103-134,
 * COMTranslate :136-153, Pivot :155-237, RandomReptation :268-312), evaluates dE
 * (k_move), takes the Metropolis decision u < exp(-beta dE) and commits — all
 * without the host.  The reference does NOT draw the acceptance variate when
 * dE >= 1e8 (simulation.cc:327-332), which shifts every later draw: a batch
 * therefore stops after such a step (`n_done`), and the caller rewinds its
 * generator to that point and continues with a new batch.
 * Crankshaft (molecule.cc:239-265): the descriptor carries the axis beads, the
 * angle and its sine / cosine (the caller's libm, i.e. the reference's); the
 * device normalises the axis and builds the rotation in the operation order of
 * Eigen's AngleAxisd::toRotationMatrix.                         */
enum { PG_MOVE_BEAD = 0, PG_MOVE_COM = 1, PG_MOVE_PIVOT = 2, PG_MOVE_CRANKSHAFT = 3, PG_MOVE_REPTATION = 4 };
typedef struct pg_move_desc {
  int32_t mol;        /* moved molecule                                                        */
  int32_t kind;       /* PG_MOVE_*                                                             */
  int32_t i0;         /* PIVOT: the pivot bead; REPTATION: direction (+1 forward, -1 backward);
                         CRANKSHAFT: first bead of the axis                                    */
  int32_t rv_offset;  /* PIVOT: first of its len-1 rows in `rvec`, in draw order;
                         CRANKSHAFT: last bead of the axis                                     */
  double s;           /* BEAD: 3*move_size/|v| (0-length v: |v|); PIVOT: move_size_rand; REPTATION: bond_len;
                         CRANKSHAFT: the angle                                                 */
  double v[3];        /* BEAD: randSphere vector; COM: displacement; REPTATION: randSphere vector;
                         CRANKSHAFT: v[0] = sin(angle), v[1] = cos(angle)                      */
  double vlen;        /* REPTATION: |v|                                                        */
  double u;           /* uniform variate of the acceptance test                                */
} pg_move_desc;
/* rvec: [n_rvec][4] = (randSphere x, y, z, bond_len of that step), PIVOT only. */
int pg_mc_upload(pg_engine* h, int n_moves, const pg_move_desc* moves, int n_rvec, const double* rvec);
/* Runs steps [first, first+count) of the uploaded batch.  *n_done = steps whose outcome is final
 * (== count unless a step hit dE >= 1e8; that step IS done — it is a rejection — later ones are
 * not).  The batch has stopped iff dE_out[n_done-1] >= 1e8 — also when that was its last step:
 * the caller's generator must then be put back in front of that step's acceptance draw.  dE_out / accept_out (may be NULL) are filled for the done steps.  elapsed_ms (may be
 * NULL) is CUDA-event time on the engine stream.  pg_mc_run = pg_mc_begin + pg_mc_end; the halves
 * let one host thread keep several replicas busy. */
int pg_mc_run(pg_engine* h, int first, int count, double* dE_out, uint8_t* accept_out, int* n_done,
              float* elapsed_ms);
int pg_mc_begin(pg_engine* h, int first, int count);
int pg_mc_end(pg_engine* h, double* dE_out, uint8_t* accept_out, int* n_done, float* elapsed_ms);
/* Trial coordinates the device built for step m of the batch ([len][3]; for tests). */
int pg_mc_trial_xyz(pg_engine* h, int m, double* xyz);

/* ---- device-resident Markov chain (whole translational steps, no host in the loop) ---- */
/* Replaces the loop body of Simulation::Run for a stretch of translational steps
 * (src/simulation/simulation.cc:216-355: the first draw of the step, TranslationalMove's pre-selection of a
 * chain and an ion, the move type, the generator's own draws — Molecule::BeadTranslate / COMTranslate /
 * Pivot / RandomReptation, src/molecules/molecule.cc:103-312, randSphere src/utilities/misc.cc:95-109 — the
 * energy change, the Metropolis test :327-332 and FinalizeEnergies) by ONE kernel that keeps the driver's
 * std::mt19937 stream on the device: the caller hands over the generator's 624 state words + position
 * (what operator<< of std::mt19937 prints), the device consumes it draw for draw like the reference —
 * including NOT drawing the acceptance variate when dE >= 1e8 — and hands the state back.  Nothing stops a
 * chain except max_steps or a grand-canonical step (gc_freq > 0 and first draw % gc_freq == 0: the chain
 * stops IN FRONT of it, the generator untouched, and the caller runs that step itself).
 * Offered for single-image systems (3 periodic axes, one box for LJ and Ewald, real-space cutoff < L/2: what
 * pg_create selects k_move<true> for); pg_chain_configure says so otherwise and the caller stays on
 * pg_mc_* / pg_delta_e.  A crankshaft's sine and cosine come from the device's sincos (<= 2 ulp from the host
 * libm's): its trial coordinates agree with the reference's to ~1e-15 instead of bit for bit.  Positions, S(k) and the running totals are the engine's own: the
 * other entry points see the chain's result (pg_download_positions, pg_get_totals, pg_delta_e ...).     */
typedef struct pg_chain_config {
  int32_t phantom;        /* molecules [0, phantom) never move (simulation.cc:255)                       */
  int32_t gc_freq;        /* > 0: ForceField::GCFrequency() of a grand-canonical run, else 0             */
  int32_t vary_bond;      /* UseBondPot(): the generators vary the bond length (simulation.cc:293-296)   */
  int32_t cluster_ctas;   /* CTAs (SMs) that share one chain: 1..16, 0 = 1                               */
  int32_t keep_trials;    /* tests: keep every step's trial coordinates (pg_chain_trial_xyz)             */
  int32_t pivot_mode;     /* 0: Molecule::Pivot in the reference's operation order (trial coordinates bit-identical
                             to the reference's; one dependent step per bead of an arm); 1: the same arm as a prefix
                             sum of independent bond vectors (coordinates equal to a few ulp, not bit for bit);
                             2: both — the energy change is evaluated from the prefix-sum coordinates while two warps
                             build the exact arm, and the exact coordinates are what an accepted move commits (and
                             what pg_chain_trial_xyz returns): bit-identical trajectories at the speed of mode 1, dE
                             equal to mode 0's to ~1e-13 relative                                            */
  double move_size;       /* s1_MC_move_size                                                             */
  double bond_len;        /* RigidBondLen() or EqBondLen() (simulation.cc:288-296)                       */
  double move_prob[5];    /* bead, COM, pivot, crankshaft, reptation (simulation.cc:46-50)               */
} pg_chain_config;
typedef struct pg_chain_step {
  double dE;              /* what ForceField::EnergyDifference returned                                  */
  int32_t mol;            /* moved molecule, -1: the step attempted nothing                              */
  int8_t kind;            /* PG_MOVE_*, -1: the step attempted nothing                                   */
  uint8_t accept;
  uint8_t stage;          /* as pg_delta.stage                                                           */
  uint8_t _pad;
} pg_chain_step;
int pg_chain_configure(pg_engine* h, const pg_chain_config* cfg);
/* The generator: 624 state words and the position (0..624), libstdc++'s std::mt19937 layout.  get returns
 * an EQUIVALENT state (same future outputs; a position of 624 comes back as 0 of the next generation). */
int pg_chain_set_rng(pg_engine* h, const uint32_t* state624, int position);
int pg_chain_get_rng(pg_engine* h, uint32_t* state624, int* position);
/* Runs up to max_steps steps.  *n_done steps are over; *stop_kind 0: max_steps reached, 1: the next step is
 * a grand-canonical step.  steps (may be NULL) receives n_done records; elapsed_ms is CUDA-event time.   */
int pg_chain_run(pg_engine* h, int max_steps, int* n_done, int* stop_kind, pg_chain_step* steps, float* elapsed_ms);
int pg_chain_begin(pg_engine* h, int max_steps);
int pg_chain_end(pg_engine* h, int* n_done, int* stop_kind, pg_chain_step* steps, float* elapsed_ms);
/* n independent replicas (one engine each, same device, same cluster size) in ONE launch on hs[0]'s stream;
 * every replica continues from its own resident generator state.  n_done[n] may be NULL.                */
int pg_chain_run_multi(pg_engine** hs, int n, int max_steps, int* n_done, float* elapsed_ms);
/* The same with the whole boundary crossing in one call and one synchronisation: rng_io [n][625] (624 state words +
 * position per replica, in and out); rng_up [n] (may be NULL = all) 0: replica i continues from the generator state
 * resident on the device (the caller has not drawn since the last call handed it back); steps [n][max_steps] (may be
 * NULL); xyz [n] pointers to [n_beads][3] accepted coordinates (may be NULL; all engines then hold the same number
 * of beads).                                                                                             */
int pg_chain_run_multi_io(pg_engine** hs, int n, int max_steps, uint32_t* rng_io, const uint8_t* rng_up,
                          pg_chain_step* steps, double* const* xyz, int* n_done, float* elapsed_ms);
/* Records [first, first+count) of the engine's last chain (also after pg_chain_run_multi).              */
int pg_chain_steps(pg_engine* h, int first, int count, pg_chain_step* steps);
/* tests: trial coordinates of step `step` of the last chain ([n_beads][3]; keep_trials), and a consistency
 * check of the resident spatial structures against the resident coordinates (*n_bad inconsistencies).   */
int pg_chain_trial_xyz(pg_engine* h, int step, double* xyz, int n_beads);
int pg_chain_check(pg_engine* h, int* n_bad);
/* Work counters of this engine's chains since the last reset: out3 = {pair configurations evaluated inside a cutoff
 * (LJ, real space, intra-molecular), steps that evaluated an energy change, accepted steps}.              */
int pg_chain_counters(pg_engine* h, uint64_t* out3, int reset);

/* ---- configurational-bias trial energies (ForceField::BeadsEnergy) ------- */
/* One launch evaluates n_trials candidate (monomer, counter-ion) pairs against
 * every resident bead except molecules [skip_mol_first, skip_mol_last] (the
 * chain being retraced, cbmc.cc:22-23; pass -1,-1 for insertion) and against
 * the partial chain grown so far (cbmc.cc:54-90).
 *   bead1_xyz [n_trials][3], bead2_xyz [n_trials][3] (ignored if !use_bead2)
 *   chain_xyz/chain_q: current_len monomers followed by current_len ions
 *   out_energy[n_trials]: what BeadsEnergy returns (1e8 sentinel included)
 *   out_pair / out_ewald (may be NULL): its pair_e and ewald_e partial sums.
 * dipole: the caller passes mz_base = sum q z over the partners the reference
 * includes (cbmc DipoleEDiff, potential_ewald_coul.cc:473-530) or NaN to have
 * the device compute it. */
typedef struct pg_trial_set {
  int32_t n_trials;
  int32_t use_bead2;
  int32_t type1, type2;
  double q1, q2;
  int32_t current_len;     /* beads already grown                              */
  int32_t skip_mol_first;  /* -1: nothing skipped                              */
  int32_t skip_mol_last;
  int32_t _pad;
} pg_trial_set;
int pg_trial_energies(pg_engine* h, const pg_trial_set* set, const double* bead1_xyz,
                      const double* bead2_xyz, const double* chain_xyz, const double* chain_q,
                      const int32_t* chain_type, double* out_energy, double* out_pair,
                      double* out_ewald);

/* ---- grand-canonical bookkeeping ------------------------------------------ */
/* Append n_new_mol molecules at the end (accepted insertion: cbmc.cc:306-320 +
 * ForceField::EnergyInitForAddedMolecule force_field.cc:1088-1105).  Adds their
 * energy with everything (and among themselves) to the running totals and
 * updates S(k).  `added` returns the energy components that were added. */
int pg_insert_molecules(pg_engine* h, int n_new_mol, const int32_t* mol_len, const double* xyz,
                        const double* q, const int32_t* type, pg_totals* added);
/* Remove molecules [mol_first, mol_last] (accepted deletion: the four
 * AdjustEnergyUponMolDeletion, cbmc.cc:422-436), compacting the arrays. */
int pg_delete_molecules(pg_engine* h, int mol_first, int mol_last, pg_totals* removed);

/* ---- reciprocal space, shardable ------------------------------------------ */
/* One sample of the slab wall-force pressure, the force sums of
 * ForceField::CalcPressureForceLJELSlit (src/force_field/pressure.cc:404-469) on the
 * resident configuration: out6 = {LJ ion, LJ polymer, LJ wall-wall, EL ion, EL polymer,
 * EL wall-wall}.  The first n_phantom molecules are the wall sites (one bead each, the
 * first half on the z = 0 plate).  The caller keeps the running averages
 * (p_tensor[6..11], wall-wall terms from the first sample only).  The LJ site-site
 * term follows the evident intent of PotentialTruncatedLJ::PairForce
 * (potential_truncated_lj.cc:87-122), whose test of an uninitialised r6 is undefined
 * behaviour in the reference. */
int pg_wall_force(pg_engine* h, int n_phantom, double* out6);

/* One sample of the volume-perturbation pressure estimator: the energy change of the resident
 * configuration under a virtual stretch of the box along z by dz, split by species pair —
 * the loops of ForceField::CalcPressureVolScalingHSELSlit (src/force_field/pressure.cc:187-338;
 * the reference fixes dz = kDz = 1e-5, src/utilities/constants.h:55).  Species: 0 cation,
 * 1 anion, 2 polymer bead, 3 surface site (the first n_phantom molecules); tables are indexed
 * i*4+j with i <= j like p_tensor_el / p_tensor_hs.  The caller keeps the running sums and the
 * Yethiraj / de Miguel averages (pressure.cc:346-384).  The pairwise reciprocal-space sum of
 * PairEnergyReplForP (potential_ewald_coul.cc:228-250) is evaluated through per-species
 * structure factors of both geometries; see plum_b200/csrc/pg_volscale.cu. */
typedef struct pg_vol_sample {
  double el[16];   /* electrostatic dU per species pair (p_tensor_el increments)        */
  double hs[16];   /* LJ / hard-sphere and wall dU per species pair (p_tensor_hs)      */
  double bond;     /* bond-potential dU (p_tensor[5] increment)                        */
  double dipole;   /* dipole-correction dU (p_tensor[4] increment)                     */
  double dU;       /* everything above                                                 */
  int32_t n_free;  /* n_mol - n_phantom                                                */
  int32_t _pad;
} pg_vol_sample;
int pg_vol_scaling_sample(pg_engine* h, int n_phantom, double dz, pg_vol_sample* out);

/* Full S(k) over all charged beads for the k slice [k_first, k_first+k_count)
 * of the half-space list; writes interleaved (re,im) to sk_dev (device pointer,
 * may alias a torch tensor) or, if NULL, into the engine's own S(k).  Used by
 * the k-sharded recompute: each rank fills its slice, NCCL all-gathers. */
int pg_sk_compute_slice(pg_engine* h, int k_first, int k_count, double* sk_dev);
/* Adopt a full S(k) (n_k_half complex numbers, device pointer) as the engine's. */
int pg_sk_set(pg_engine* h, const double* sk_dev);
/* Reciprocal energy of the k slice from the engine's S(k). */
int pg_sk_energy(pg_engine* h, int k_first, int k_count, double* e_out);
int pg_sk_download(pg_engine* h, double* sk_host /* [n_k_half][2] */);
/* Full recompute of S(k) from the resident coordinates into the engine's own S(k) — the drift reset of the incrementally
 * updated structure factor (the reference's counterpart of a full recompute is PotentialEwald::EnergyInitialization,
 * src/force_field/potential_ewald.cc:176-230).  The running energy totals are NOT touched (the reference accumulates
 * them, src/force_field/force_field.cc:436-451); *recip_energy (may be NULL) receives the reciprocal energy of the fresh
 * S(k).  When peers are attached the recompute is k-sharded: this rank computes a contiguous slice of the k list and
 * the reduction epilogue stores it into every peer's S(k) through peer-mapped memory (NVLink) and handshakes with flags
 * — a collective: every attached rank must call it.  _begin / _end let one thread drive several ranks.             */
int pg_recompute_sk(pg_engine* h, double* recip_energy);
int pg_recompute_sk_begin(pg_engine* h);
int pg_recompute_sk_end(pg_engine* h, double* recip_energy, float* elapsed_ms);
/* Peers of the k-sharded recompute.  Every rank exports a 64-byte handle (cudaIpcMemHandle_t) of its exchange block
 * [S(k) | flags]; the caller exchanges the handles (any transport) and attaches all of them: handles is [world][64],
 * the entry of `rank` itself is ignored.  _attach_local is the same for engines of ONE process (other GPUs with peer
 * access, or the same GPU).  world <= 16; all ranks hold the same k list (same force field and box).               */
int pg_sk_export(pg_engine* h, void* handle64);
int pg_sk_attach(pg_engine* h, int rank, int world, const void* handles);
int pg_sk_attach_local(pg_engine* h, int rank, int world, pg_engine* const* peers);
int pg_sk_detach(pg_engine* h);

/* ---- instrumentation ------------------------------------------------------ */
/* Number of kernels this engine has launched since creation. */
uint64_t pg_launch_count(const pg_engine* h);
/* The CUDA stream the engine launches on (cudaStream_t as void*). */
void* pg_stream(const pg_engine* h);
/* FP64 FMA throughput microbenchmark (GFLOP/s, FMA = 2 flop) used as the
 * roofline denominator the driver's MEASURED_PEAKS.json does not carry. */
int pg_measure_fp64_peak(pg_engine* h, double* gflops);
/* L2 streaming-read bandwidth microbenchmark (GB/s, 32 MiB L2-resident buffer). */
int pg_measure_l2_peak(pg_engine* h, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* PLUM_B200_H_ */
