// plum_b200 — sm_100a kernels of the per-trial-move energy path (device side).
// Host side / C ABI: pg_engine.cu.  Arithmetic core: pg_math.cuh.
//
// Kernel 1  k_delta      fused "bead group vs everything" energy change:
//                        LJ + Ewald real-space over partner tiles, dS(k) and the
//                        reciprocal-space change over k tiles, and (last CTA to
//                        finish) wall + bond + self + dipole terms, the reference's
//                        early-return orchestration and, in replay mode, the
//                        Metropolis decision.
// Kernel 2  k_commit     apply an accepted trial: positions, S(k) += dS(k), totals.
// Kernel 3  k_tot_pairs  all-pairs LJ + real-space totals (energy initialisation).
// Kernel 4  k_sk_slice   full S(k) for a slice of the k list (shardable).
// Kernel 5  k_tot_final  O(N) / O(K) totals: walls, bonds, self, dipole, |S|^2.
// Kernel 6  k_trials     batched Rosenbluth trial energies (CBMC).
#ifndef PLUM_B200_PG_KERNELS_CUH_
#define PLUM_B200_PG_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "pg_math.cuh"

#define PG_TILE 128       // threads per CTA == partners per tile
#define PG_GCHUNK 32      // max group beads per CTA chunk
#define PG_KTILE 128      // k vectors per CTA in k_delta
#define PG_MAX_PARTIALS 8192

enum { PG_MODE_MOVE = 0, PG_MODE_INSERT = 1, PG_MODE_DELETE = 2 };

// Device-resident bookkeeping: running totals (the reference's E_tot members),
// the dipole lag state (potential_ewald.cc:224-228,527-531,615-622) and the
// pending trial's components.
struct PgState {
  double E_pair, E_ewald, E_bond, E_ext;
  double E_real, E_recip, E_self;     // informational split of E_ewald
  double cur_dipl, trial_dipl;
  // pending trial
  double dE, d_pair, d_ext, d_ewald, d_bond, d_real, d_recip, d_self, d_dipole;
  double mz_current, mz_group_old, mz_group_new;
  int stage, n_overlap, accept, mode;
  int n_beads;
  unsigned int done_counter;
  unsigned int seq;
  int pad;
};

// What the host reads back after a trial (lives in mapped pinned host memory).
struct PgResult {
  double dE, pair, ext, ewald, bond, real, recip, mz_current, self_e, dipole;
  int stage, n_overlap, accept, pad;
  volatile unsigned int seq;   // written last, after __threadfence_system()
  unsigned int pad2;
};

struct PgDeltaArgs {
  // resident system
  const double2* xy;
  const double2* zq;
  const int* type;
  int n_partners;        // resident beads that act as partners
  // group
  int g0, glen;          // bead range of the group in the resident arrays
  int mode;              // PG_MODE_*
  int has_old, has_new;
  const double* trial;   // [glen][3] trial / new coordinates
  const double* gq;      // [glen] charges (INSERT: not yet resident)
  const int* gtype;      // [glen]
  const uint8_t* moved;  // [glen]
  int chain_len_first;   // beads of the first molecule of the group (HS bonded exclusion, bonds)
  int bond_first, bond_len;  // group-relative bead range whose spring energy changes
  // reciprocal space
  const int* kl;         // [nk][4] integer triplets (lx,ly,lz,pad), half space
  const double* ek2;     // [nk]
  const double2* S;      // [nk] current structure factor
  double2* dS;           // [nk] out: change of the structure factor
  int nk;
  // tiling
  int n_tiles, n_chunks, chunk_size, n_pair_ctas, n_k_ctas;
  // scratch + outputs
  double* partial;       // [n_ctas][4]
  int* partial_i;        // [n_ctas]
  PgState* state;
  PgResult* result;      // mapped host memory (may be NULL in replay)
  // replay
  int decide_on_device;
  double u;              // uniform variate for the device-side Metropolis test
  double* replay_dE;     // optional device log
  uint8_t* replay_acc;
  int replay_index;
  unsigned int seq;
};

#endif
