// plum_b200 — host side of the engine and the C ABI (include/plum_b200.h).
//
// One pg_engine = one system replica resident on one GPU.  The host keeps only
// small mirrors (molecule table, per-bead charge/type) so it can describe a bead
// group to the kernels without reading the device back; coordinates, S(k) and
// the running energy totals are device-authoritative.
//
// Reference behaviour replaced (file:line relative to /root/reference):
//   ForceField::Initialize / InitializeEnergy   src/force_field/force_field.cc:46-405
//   ForceField::EnergyDifference                src/force_field/force_field.cc:407-434
//   ForceField::FinalizeEnergies                src/force_field/force_field.cc:436-451
//   ForceField::BeadsEnergy                     src/force_field/cbmc.cc:5-151
//   ForceField::EnergyInitForAddedMolecule      src/force_field/force_field.cc:1088-1105
//   *::AdjustEnergyUponMolDeletion              src/force_field/cbmc.cc:422-436
//   PotentialEwaldCoul::ReadParameters          src/force_field/potential_ewald_coul.cc:29-132
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/plum_b200.h"
#include "pg_kernels.cu"
#include "pg_move.cu"
#include "pg_trials.cu"
#include "pg_pressure.cu"
#include "pg_volscale.cu"
#include "pg_propose.cu"
#include "pg_chain.cu"
#include "pg_sk.cu"

namespace {

// src/utilities/constants.h:5-62 (truncated on purpose: parity with the reference)
const double kPi = 3.14159265359;
const double kEwaldCutoff = 1E-7;
const int kDiCorrection = 3;
const double k213 = 1.25992104989;
const double k216 = 1.12246204831;

struct ReplayMove {
  int mol, g0, glen, off, nq;  // off: bead offset into the packed proposal buffers
  double u;
};

}  // namespace

struct pg_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  PgDev P;
  pg_ewald_info info;
  double alpha = 0.0;           // Ewald alpha as given (tables of the stretched box, pg_vol_scaling_sample)
  std::string err;
  uint64_t launches = 0;

  // resident system
  int cap = 0, n = 0, n_mol = 0;
  double2 *xy = nullptr, *zq = nullptr;
  int *type = nullptr, *mol = nullptr;
  double2 *t_xy = nullptr, *t_zq = nullptr;   // compaction scratch
  int *t_type = nullptr, *t_mol = nullptr;
  std::vector<int> mol_first;   // host mirror [n_mol+1]
  std::vector<double> h_q;      // host mirror of charges
  std::vector<int> h_type;      // host mirror of types

  // reciprocal space
  int nk = 0;
  std::vector<int> h_kl;        // [nk][4]
  std::vector<double> h_ek2;
  std::vector<double> h_kvec;   // [nk][4] (kx, ky, kz, ek2)
  double4* d_kvec = nullptr;
  int* d_kl = nullptr;
  double* d_ek2 = nullptr;
  double2 *d_S = nullptr, *d_dS = nullptr, *d_Stmp = nullptr;

  // group staging: two pinned + two device blocks (ping-pong: the previous trial's block must
  // stay intact until its deferred commit has been applied by the next k_move launch)
  int gcap = 0;
  char* h_stage2[2] = {nullptr, nullptr};
  char* d_stage2[2] = {nullptr, nullptr};
  char* h_stage = nullptr;      // current slot (aliases h_stage2[slot])
  char* d_stage = nullptr;
  int slot = 0;
  size_t stage_bytes = 0;
  bool inflight = false;        // pg_delta_e_begin issued, pg_delta_e_poll has not returned 0 yet
  bool fast = false;            // k_move<true> hot loop is valid for this force field

  // deferred commit of the last trial (applied by the next k_move or by flush_commit)
  struct { bool valid; int accept; int g0, glen; const char* d_group; int cap; const char* h_group; bool on_device; }
      pc = {false, 0, 0, 0, nullptr, 0, nullptr, true};

  // scratch
  int partial_cap = 0;
  double* d_partial = nullptr;
  int* d_partial_i = nullptr;
  double* d_out8 = nullptr;     // 8 doubles
  double* h_out8 = nullptr;     // pinned
  PgState* d_state = nullptr;
  unsigned long long* d_timing = nullptr;   // debug stamps (PLUM_B200_TIMING=1)
  int timing_ctas = 0;
  PgDev* d_P = nullptr;          // device copy of P for non-inlined device functions
  PgResult* h_result = nullptr; // mapped pinned
  PgResult* d_result = nullptr; // device alias of h_result
  PgMailRec* h_mail = nullptr;  // mapped pinned: k_move's self-validating result records
  PgMailRec* d_mail = nullptr;
  int move_slots = 0;           // resident k_move CTAs on the device (one wave)
  int move_per_sm = 0;          // ... per SM
  int move_per_sm_fixed = 0;    // PLUM_B200_CTAS_PER_SM override (0: adaptive)
  int n_sm = 0;
  unsigned int seq = 0;
  bool use_mailbox = true;
  bool use_inline = true;       // small groups travel in the kernel parameters (PLUM_B200_NO_INLINE=1 disables)

  // pending trial
  bool pending = false;
  int pend_mode = 0, pend_g0 = 0, pend_glen = 0, pend_cap = 0;
  const char* pend_group = nullptr;
  const char* pend_hgroup = nullptr;
  bool pend_on_device = true;

  // CBMC staging
  // k_trials: staging for growth steps too large for the kernel parameters, partial sums, mailbox
  int ccap = 0;       // chain capacity of the staging block
  double *h_trial_in = nullptr, *d_trial_in = nullptr;    // b1[3 TR_MAXT] | b2[3 TR_MAXT] | chain_xyz[3 ccap] | chain_q[ccap]
  int *h_trial_ct = nullptr, *d_trial_ct = nullptr;
  double2* d_sk_partial = nullptr; size_t sk_partial_cap = 0;   // k_sk_slice [chunk][k]
  double2* d_tr_partial = nullptr; size_t tr_partial_cap = 0;
  double* d_tr_mz = nullptr; size_t tr_mz_cap = 0;
  unsigned int* d_tr_counter = nullptr;
  PgMailRec *h_tmail = nullptr, *d_tmail = nullptr;       // mapped, [3 TR_MAXT]
  int trial_slots = 0;                                    // resident k_trials CTAs

  // replay
  std::vector<ReplayMove> rp_moves;
  char* d_rp = nullptr;
  size_t rp_beads = 0;
  double* d_rp_dE = nullptr;
  uint8_t* d_rp_acc = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // CUDA graphs of replay batches, keyed by (first, count, commit): the per-move launches of a batch are
  // captured once so that the replay is not bound by the host's launch rate
  struct RpGraph { int first, count, commit; cudaGraphExec_t exec; };
  std::vector<RpGraph> rp_graphs;
  bool use_graph = true;        // PLUM_B200_NO_GRAPH=1 disables

  // batched Markov-chain steps with device-side proposals (pg_mc_*): same packed layout as the replay,
  // but the trial coordinates are written by k_propose and the buffers only ever grow
  bool rp_mc = false;           // rp_moves / d_rp currently hold an MC batch
  size_t rp_bytes_cap = 0, rp_log_cap = 0;
  PgPropDev* d_mc_moves = nullptr; size_t mc_moves_cap = 0;
  double4* d_mc_rv = nullptr; size_t mc_rv_cap = 0;
  char* h_mc_pin = nullptr; size_t mc_pin_cap = 0;   // pinned upload staging
  int* d_stop = nullptr;         // device memory: every CTA of every k_move / k_propose of a batch reads it first
  int* h_stop = nullptr;         // pinned copy, filled by a 4-byte async copy behind the batch
  double *h_mc_dE = nullptr, *d_mc_dE = nullptr;     // per-step logs in mapped pinned memory (device alias d_*):
  uint8_t *h_mc_acc = nullptr, *d_mc_acc = nullptr;  // the finalising CTA of k_move posts them straight to the host
  size_t mc_log_cap = 0;
  std::vector<int> mc_mol;      // molecule of every step of the batch (wave construction)
  int mc_first = 0, mc_count = 0;
  bool mc_inflight = false;

  // device-resident Markov chain (k_chain, pg_chain_*): resident spatial structures + io buffers
  PgChainHost ch;

  // full S(k) recompute (pg_sk.cu): compact list of the charged beads (rebuilt from the host mirror of the charges when
  // the bead set changes), the exchange block [S(k) | flags] other ranks map, the peers of a k-sharded recompute
  int* d_qidx = nullptr; size_t qidx_cap = 0; int nq_tot = 0; bool qidx_valid = false;
  char* d_xchg = nullptr;              // one allocation: nk_alloc double2 (d_S points here) + 2 * SK_MAX_PEERS flags
  unsigned int* d_flags = nullptr;
  unsigned int* d_sk_counter = nullptr;
  int sk_kmax[3] = {0, 0, 0};
  double sk_kunit[3] = {0, 0, 0};
  std::vector<PgSkItem> h_sk_items;     // k columns in segments of <= SK_SEG consecutive lz (pg_sk.cu)
  std::vector<int> h_sk_item_n;
  PgSkItem* d_sk_items = nullptr;
  int* d_sk_item_n = nullptr;
  PgSkPeers peers;                      // world <= 1: not sharded
  void* ipc_opened[SK_MAX_PEERS] = {nullptr};
  bool sk_inflight = false;
};

#define PG_CUDA(h, call)                                                                     \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
      return PG_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

namespace {

static std::string g_create_err;
// engines alive in this process, per CUDA device: how many replicas share a GPU decides how wide one
// k_move grid should be (move_slots_now)
static std::atomic<int> g_live_engines[64];

// Stage layout for a group of `cap` beads.
struct StageView {
  double* trial;   // [cap][3]
  double* gq;      // [cap]
  int* gtype;      // [cap]
  int* qidx;       // [cap] group-relative indices of the moved AND charged beads (first nq entries)
  uint8_t* moved;  // [cap]
};
size_t stage_size(int cap) { return sizeof(double) * 4 * (size_t)cap + sizeof(int) * 2 * (size_t)cap + (size_t)cap + 64; }
StageView stage_view(char* base, int cap) {
  StageView v;
  v.trial = reinterpret_cast<double*>(base);
  v.gq = v.trial + 3 * (size_t)cap;
  v.gtype = reinterpret_cast<int*>(v.gq + cap);
  v.qidx = v.gtype + cap;
  v.moved = reinterpret_cast<uint8_t*>(v.qidx + cap);
  return v;
}
// Fills qidx from gq / moved; returns nq.
int stage_fill_qidx(const StageView& v, int off, int glen) {
  int nq = 0;
  for (int i = 0; i < glen; i++)
    if (v.moved[off + i] && v.gq[off + i] != 0.0) v.qidx[off + nq++] = i;
  return nq;
}
int lanes_per_k(int nq) {
  int l = 2;
  while (l < 2 * nq && l < 32) l <<= 1;
  return l;
}

// PotentialEwaldCoul::ReadParameters, potential_ewald_coul.cc:29-132.
void ewald_setup(pg_engine* h, const pg_params* p) {
  PgDev& P = h->P;
  pg_ewald_info& I = h->info;
  memset(&I, 0, sizeof(I));
  double box_l[3] = {p->box[0], p->box[1], p->box[2]};
  if (p->dipole_correction) {
    double min_padding = 150;
    if (box_l[2] * kDiCorrection > min_padding)
      box_l[2] = box_l[2] * kDiCorrection;
    else
      box_l[2] += min_padding;
  }
  double box_vol = box_l[0] * box_l[1] * box_l[2];
  double lB = p->lB, alpha = p->alpha;
  h->alpha = alpha;
  double real_cutoff = 1;
  while (0.5 * lB * 1 * 1 * erfc(sqrt(alpha) * real_cutoff) / real_cutoff > kEwaldCutoff) real_cutoff += 1;
  int real_cell[3], repl_cell[3], ceto[3];
  for (int i = 0; i < 3; i++) real_cell[i] = (int)ceil(real_cutoff / box_l[i]);
  double repl_cutoff = alpha;
  double min_box_vol = pow(box_l[0], 3);
  while (lB * 1 * 1 / (2 * kPi * min_box_vol) * (4 * kPi * kPi) * exp(-repl_cutoff / (4 * alpha)) / repl_cutoff >
         kEwaldCutoff)
    repl_cutoff += alpha;
  for (int i = 0; i < 3; i++) repl_cell[i] = (int)ceil(sqrt(repl_cutoff) * box_l[i] / (2 * kPi));
  for (int i = 0; i < 3; i++) ceto[i] = 2 * repl_cell[i] + 1;
  std::vector<double> kx(ceto[0]), ky(ceto[1]), kz(ceto[2]);
  for (int l = -repl_cell[0]; l <= repl_cell[0]; l++) kx[l + repl_cell[0]] = l * 2 * kPi / box_l[0];
  for (int l = -repl_cell[1]; l <= repl_cell[1]; l++) ky[l + repl_cell[1]] = l * 2 * kPi / box_l[1];
  for (int l = -repl_cell[2]; l <= repl_cell[2]; l++) kz[l + repl_cell[2]] = l * 2 * kPi / box_l[2];
  h->h_kl.clear();
  h->h_ek2.clear();
  h->h_kvec.clear();
  int n_k = 0;
  for (int ix = 0; ix < ceto[0]; ix++)
    for (int iy = 0; iy < ceto[1]; iy++)
      for (int iz = 0; iz < ceto[2]; iz++) {
        double k2 = kx[ix] * kx[ix] + ky[iy] * ky[iy] + kz[iz] * kz[iz];
        double ek2 = exp(-k2 / (4 * alpha)) / k2;
        if (!(k2 > 0 && k2 <= repl_cutoff)) continue;
        n_k++;
        int ax = ix - repl_cell[0], ay = iy - repl_cell[1], az = iz - repl_cell[2];
        // S(-k) = conj(S(k)) and ek2(-k) == ek2(k) bit for bit: keep the half space, weight 2
        bool pos = (ax > 0) || (ax == 0 && ay > 0) || (ax == 0 && ay == 0 && az > 0);
        if (!pos) continue;
        h->h_kl.push_back(ax); h->h_kl.push_back(ay); h->h_kl.push_back(az); h->h_kl.push_back(0);
        h->h_ek2.push_back(ek2);
        h->h_kvec.push_back(kx[ix]); h->h_kvec.push_back(ky[iy]); h->h_kvec.push_back(kz[iz]); h->h_kvec.push_back(ek2);
      }
  h->nk = (int)h->h_ek2.size();
  for (int i = 0; i < 3; i++) {
    I.ewald_box[i] = box_l[i];
    I.real_cell[i] = real_cell[i];
    I.repl_cell[i] = repl_cell[i];
    P.ebox[i] = box_l[i];
    P.inv_ebox[i] = 1.0 / box_l[i];
    P.real_cell[i] = real_cell[i];
  }
  I.box_vol = box_vol;
  I.real_cutoff = real_cutoff;
  I.repl_cutoff = repl_cutoff;
  I.n_k = n_k;
  I.n_k_half = h->nk;
  P.lB = lB;
  P.sqrt_alpha = sqrt(alpha);
  P.real_cutoff = real_cutoff;
  P.rc_relaxed = real_cutoff * (1.0 + 1e-9);
  P.rc2_relaxed = P.rc_relaxed * P.rc_relaxed;
  P.recip_pref = 2 * kPi * lB / box_vol;
  P.self_pref = -lB * sqrt(alpha / kPi);
  P.dipole_pref = lB * 2 * kPi / (box_vol / 1.0);
  // only the central image can pass r <= real_cutoff when every axis is periodic and
  // the relaxed cutoff stays below half the shortest cell edge
  bool single = (p->npbc == 3);
  for (int i = 0; i < 3; i++)
    if (!(P.rc_relaxed < 0.5 * box_l[i] * (1.0 - 1e-9))) single = false;
  P.single_image = single ? 1 : 0;
  // a bead's own periodic images: position independent, so summed once here (potential_ewald.cc self-image term)
  P.real_self_unit = 0.5 * pg_pair_real_d(P, 0.0, 0.0, 0.0, 1.0);
}

int build_params(pg_engine* h, const pg_params* p) {
  PgDev& P = h->P;
  memset(&P, 0, sizeof(P));
  if (p->n_types < 1 || p->n_types > PG_MAX_TYPES) {
    g_create_err = "n_types out of range (1.." + std::to_string(PG_MAX_TYPES) + ")";
    return PG_ERR_CAPACITY;
  }
  if (p->npbc < 2 || p->npbc > 3) {
    g_create_err = "npbc must be 2 or 3";
    return PG_ERR_INVALID;
  }
  for (int i = 0; i < 3; i++) {
    if (!(p->box[i] > 0)) { g_create_err = "box lengths must be positive"; return PG_ERR_INVALID; }
    P.box[i] = p->box[i];
    P.inv_box[i] = 1.0 / p->box[i];
    P.half_box[i] = 0.5 * p->box[i];
    P.pbc[i] = (i < p->npbc) ? 1 : 0;
    P.ebox[i] = p->box[i];
    P.inv_ebox[i] = 1.0 / p->box[i];
  }
  if (p->pair_kind == PG_PAIR_TRUNCATED_LJ && (!p->lj_sigma || !p->lj_epsilon)) { g_create_err = "TruncatedLJ needs lj_sigma and lj_epsilon"; return PG_ERR_INVALID; }
  if (p->pair_kind == PG_PAIR_HARD_SPHERE && !p->hs_radius) { g_create_err = "HardSphere needs hs_radius"; return PG_ERR_INVALID; }
  if (p->pair_kind < PG_PAIR_NONE || p->pair_kind > PG_PAIR_HARD_SPHERE) { g_create_err = "unknown pair_kind"; return PG_ERR_INVALID; }
  P.n_types = p->n_types;
  P.beta = p->beta;
  P.pair_kind = p->pair_kind;
  P.lj_cutoff = p->lj_cutoff;
  for (int a = 0; a < p->n_types; a++)
    for (int b = 0; b < p->n_types; b++) {
      int tp = a * PG_MAX_TYPES + b;
      if (p->pair_kind == PG_PAIR_TRUNCATED_LJ) {
        // potential_truncated_lj.cc:55-82
        double sigma = (p->lj_sigma[a] + p->lj_sigma[b]) / 2;
        double epsilon = sqrt((p->lj_epsilon[a] * p->lj_epsilon[b]));
        double r6_ref, rcut;
        if (p->lj_cutoff < 0) {
          r6_ref = pow((1.0 / k216), 6);
          rcut = k216 * sigma;
        } else {
          r6_ref = pow((sigma / p->lj_cutoff), 6);
          rcut = p->lj_cutoff;
        }
        P.lj_sigma[tp] = sigma;
        P.lj_eps4[tp] = 4 * epsilon;
        P.lj_eref[tp] = 4 * epsilon * (r6_ref * r6_ref - r6_ref);
        P.lj_rcut[tp] = rcut;
        double rr = rcut * (1.0 + 1e-9);
        P.lj_rcut2_relaxed[tp] = rr * rr;
        P.lj_rcut2_relaxed_max = std::max(P.lj_rcut2_relaxed_max, rr * rr);
      } else if (p->pair_kind == PG_PAIR_HARD_SPHERE) {
        P.hs_allowed[tp] = p->hs_radius[a] + p->hs_radius[b];
      }
    }
  P.use_ewald = p->use_ewald ? 1 : 0;
  P.dipole = (p->use_ewald && p->dipole_correction) ? 1 : 0;
  h->nk = 0;
  memset(&h->info, 0, sizeof(h->info));
  if (p->use_ewald) {
    if (!(p->alpha > 0) || !(p->lB >= 0)) { g_create_err = "Ewald alpha must be > 0"; return PG_ERR_INVALID; }
    ewald_setup(h, p);
  }
  for (int i = 0; i < 3; i++) P.same_box[i] = (P.ebox[i] == P.box[i]) ? 1 : 0;
  P.bond_kind = p->bond_kind;
  P.bond_k = p->bond_k;
  P.bond_r0 = p->bond_r0;
  P.ext_kind = p->ext_kind;
  P.wall_cut = p->wall_cut;
  P.well_width = p->well_width;
  P.well_depth = p->well_depth;
  for (int t = 0; t < p->n_types; t++) {
    P.wall_sigma[t] = p->wall_sigma ? p->wall_sigma[t] : 0.0;
    P.wall_eps[t] = p->wall_epsilon ? p->wall_epsilon[t] : 0.0;
    P.graft[t] = p->graft_kind ? p->graft_kind[t] : 0;
    if (p->ext_kind == PG_EXT_TRUNCATED_LJ_WALL) {
      // potential_truncated_lj_wall.cc:69-70,117-118: reference energy of the branch in use
      double r3_ref = (p->wall_cut < 0 || P.graft[t] != 0) ? pow((1.0 / k213), 3) : pow((1.0 / p->wall_cut), 3);
      P.wall_r3ref_e[t] = 2.59807621135 * P.wall_eps[t] * (r3_ref * r3_ref - r3_ref);
    }
  }
  return PG_OK;
}

int flush_commit(pg_engine* h);

// Cached replay / MC-batch state bakes device pointers (xy, zq, d_partial), bead ranges and grid shapes into CUDA graphs
// and ReplayMove records.  Anything that reallocates those arrays drops the graphs; anything that changes the molecule
// table (upload, insertion, deletion) also drops the recorded moves, so that a later pg_replay_run / pg_mc_begin reports
// PG_ERR_INVALID instead of launching with stale ranges.
void replay_invalidate(pg_engine* h, bool drop_moves) {
  for (auto& g : h->rp_graphs) cudaGraphExecDestroy(g.exec);
  h->rp_graphs.clear();
  if (drop_moves) {
    h->rp_moves.clear();
    h->mc_mol.clear();
    h->rp_mc = false;
  }
}

int ensure_group(pg_engine* h, int glen) {
  if (glen <= h->gcap) return PG_OK;
  int rc = flush_commit(h);   // the pending trial's block is about to be freed
  if (rc) return rc;
  int cap = std::max(256, glen * 2);
  h->stage_bytes = stage_size(cap);
  for (int s = 0; s < 2; s++) {
    if (h->h_stage2[s]) cudaFreeHost(h->h_stage2[s]);
    if (h->d_stage2[s]) cudaFree(h->d_stage2[s]);
    h->h_stage2[s] = nullptr; h->d_stage2[s] = nullptr;
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_stage2[s], h->stage_bytes, cudaHostAllocDefault));
    PG_CUDA(h, cudaMalloc((void**)&h->d_stage2[s], h->stage_bytes));
  }
  h->slot = 0;
  h->h_stage = h->h_stage2[0];
  h->d_stage = h->d_stage2[0];
  h->gcap = cap;
  return PG_OK;
}

// Next staging slot (the other one keeps the previous trial until its commit is applied).
void next_slot(pg_engine* h) {
  h->slot ^= 1;
  h->h_stage = h->h_stage2[h->slot];
  h->d_stage = h->d_stage2[h->slot];
}

int ensure_partials(pg_engine* h, int n_ctas) {
  if (n_ctas <= h->partial_cap) return PG_OK;
  int cap = std::max(4096, n_ctas * 2);
  if (h->d_partial) { PG_CUDA(h, cudaStreamSynchronize(h->stream)); replay_invalidate(h, false); }
  if (h->d_partial) cudaFree(h->d_partial);
  if (h->d_partial_i) cudaFree(h->d_partial_i);
  h->d_partial = nullptr; h->d_partial_i = nullptr;
  PG_CUDA(h, cudaMalloc((void**)&h->d_partial, sizeof(double) * 8 * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&h->d_partial_i, sizeof(int) * (size_t)cap));
  h->partial_cap = cap;
  return PG_OK;
}

int ensure_capacity(pg_engine* h, int n_need) {
  if (n_need <= h->cap) return PG_OK;
  int cap = std::max(n_need * 2, 1024);
  double2 *xy, *zq, *txy, *tzq;
  int *type, *mol, *ttype, *tmol;
  PG_CUDA(h, cudaMalloc((void**)&xy, sizeof(double2) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&zq, sizeof(double2) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&type, sizeof(int) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&mol, sizeof(int) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&txy, sizeof(double2) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&tzq, sizeof(double2) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&ttype, sizeof(int) * (size_t)cap));
  PG_CUDA(h, cudaMalloc((void**)&tmol, sizeof(int) * (size_t)cap));
  if (h->n > 0) {
    PG_CUDA(h, cudaMemcpyAsync(xy, h->xy, sizeof(double2) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(zq, h->zq, sizeof(double2) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(type, h->type, sizeof(int) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(mol, h->mol, sizeof(int) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
    PG_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  replay_invalidate(h, false);
  cudaFree(h->xy); cudaFree(h->zq); cudaFree(h->type); cudaFree(h->mol);
  cudaFree(h->t_xy); cudaFree(h->t_zq); cudaFree(h->t_type); cudaFree(h->t_mol);
  h->xy = xy; h->zq = zq; h->type = type; h->mol = mol;
  h->t_xy = txy; h->t_zq = tzq; h->t_type = ttype; h->t_mol = tmol;
  h->cap = cap;
  return PG_OK;
}

// One step of a host spin on mapped memory: yields now and then (long wait or oversubscribed host: let a
// sibling replica's thread run), and every ~1M spins checks the stream for a failed kernel and the 30 s
// deadline.  Returns PG_OK to keep spinning.
struct SpinWait {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  int step(pg_engine* h) {
    if ((++spins & 0x1fff) == 0) sched_yield();
    if ((spins & 0xfffff) != 0) return PG_OK;
    cudaError_t q = cudaStreamQuery(h->stream);
    if (q != cudaSuccess && q != cudaErrorNotReady) {
      h->err = std::string("kernel failed: ") + cudaGetErrorString(q);
      return PG_ERR_CUDA;
    }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) {
      h->err = "device did not answer within 30 s";
      return PG_ERR_TIMEOUT;
    }
    return PG_OK;
  }
};

// Wait for the result mailbox (or the stream) — the driver's call is synchronous.
int wait_result(pg_engine* h, unsigned int seq) {
  if (h->use_mailbox) {
    SpinWait w;
    while (h->h_result->seq != seq) {
      int rc = w.step(h);
      if (rc) return rc;
    }
    return PG_OK;
  }
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->h_result->seq != seq) { h->err = "result sequence mismatch"; return PG_ERR_STATE; }
  return PG_OK;
}

// Launch k_delta for the group currently described in h_stage (already filled).
int launch_delta(pg_engine* h, int mode, int g0, int glen, int chain_len_first, int bond_first, int bond_len,
                 const char* d_group, int group_cap, bool decide_on_device, double u, int replay_index,
                 bool want_result) {
  PgDeltaArgs A;
  memset(&A, 0, sizeof(A));
  A.xy = h->xy; A.zq = h->zq; A.type = h->type;
  A.n_partners = h->n;
  A.g0 = g0; A.glen = glen; A.mode = mode;
  A.has_old = (mode != PG_MODE_INSERT);
  A.has_new = (mode != PG_MODE_DELETE);
  StageView dv = stage_view(const_cast<char*>(d_group), group_cap);
  A.trial = dv.trial; A.gq = dv.gq; A.gtype = dv.gtype; A.moved = dv.moved;
  A.chain_len_first = chain_len_first;
  A.bond_first = bond_first; A.bond_len = bond_len;
  A.kl = h->d_kl; A.ek2 = h->d_ek2; A.S = h->d_S; A.dS = h->d_dS; A.nk = h->nk;
  A.n_tiles = std::max(1, (h->n + PG_TILE - 1) / PG_TILE);
  A.n_chunks = std::max(1, (glen + PG_GCHUNK - 1) / PG_GCHUNK);
  if (h->P.use_ewald && !h->P.single_image) {
    // multi-image real space (100+ images per pair): a thread should carry few group beads — split the
    // group until the grid covers the device (small systems have one or two partner tiles)
    const int want = (2 * std::max(1, h->n_sm) + A.n_tiles - 1) / A.n_tiles;
    A.n_chunks = std::min(std::max(glen, 1), std::max(A.n_chunks, want));
  }
  A.chunk_size = std::max(1, (glen + A.n_chunks - 1) / A.n_chunks);
  A.n_chunks = std::max(1, (glen + A.chunk_size - 1) / A.chunk_size);
  A.n_pair_ctas = A.n_tiles * A.n_chunks;
  A.n_k_ctas = h->P.use_ewald ? (h->nk + PG_KTILE - 1) / PG_KTILE : 0;
  int n_ctas = A.n_pair_ctas + A.n_k_ctas;
  int rc = ensure_partials(h, n_ctas);
  if (rc) return rc;
  A.partial = h->d_partial; A.partial_i = h->d_partial_i;
  A.state = h->d_state;
  A.result = want_result ? h->d_result : nullptr;
  A.decide_on_device = decide_on_device ? 1 : 0;
  A.u = u;
  A.replay_dE = decide_on_device ? h->d_rp_dE : nullptr;
  A.replay_acc = decide_on_device ? h->d_rp_acc : nullptr;
  A.replay_index = replay_index;
  A.seq = ++h->seq;
  k_delta<<<n_ctas, PG_TILE, 0, h->stream>>>(h->P, A);
  h->launches++;
  PG_CUDA(h, cudaGetLastError());
  return PG_OK;
}

int launch_commit(pg_engine* h, int accept_flag, int mode, int g0, int glen, const char* d_group, int group_cap) {
  StageView dv = stage_view(const_cast<char*>(d_group), group_cap);
  int nthreads = std::max(std::max(glen, h->nk), 1);
  int nb = (nthreads + 255) / 256;
  k_commit<<<nb, 256, 0, h->stream>>>(h->P, accept_flag, mode, g0, glen, dv.trial, dv.gq, dv.gtype, h->xy, h->zq,
                                      h->type, h->d_S, h->d_dS, h->P.use_ewald ? h->nk : 0, h->d_state);
  h->launches++;
  PG_CUDA(h, cudaGetLastError());
  return PG_OK;
}

// Grid of a k_move launch, at most the CTAs resident at once (one wave on the 148 SMs).  Helper CTAs
// come first in block-index order — for a big move exactly one per SM — and share the reciprocal part
// (lanes-per-k chosen so a helper's k slice fits its threads in one pass whenever possible) and the
// intra-molecular pairs; the remaining CTAs split the (partner tile x moved bead) units evenly.
// CTA slots one k_move grid may take.  A single Markov chain on the GPU is latency-bound and wants the
// widest grid that still fits one wave (all but one slot per SM).  When several replicas share the GPU
// (engines of this process on the same device; separate processes set PLUM_B200_CTAS_PER_SM) throughput
// is what matters and narrow grids win: kernels of different replicas overlap, tails and imbalance of one
// hide behind the main phase of another.  Measured on S with 16 replicas and 128-thread CTAs
// (tools/sweep_replicas.py): 7 slots/SM 283 k moves/s, 5: 341 k, 3: 379 k, 2: 376 k; one replica: 7 slots
// 73 k, 4 slots 66 k.  The generic path runs small systems: always every slot.
int move_slots_now(const pg_engine* h) {
  const int n_sm = std::max(1, h->n_sm);
  if (h->move_per_sm_fixed) return h->move_per_sm_fixed * n_sm;
  if (!h->fast) return h->move_per_sm * n_sm;
  const int hi = std::max(1, h->move_per_sm - 1), lo = std::min(3, hi);
  const int live = (h->device >= 0 && h->device < 64) ? g_live_engines[h->device].load(std::memory_order_relaxed) : 1;
  if (live <= 1) return hi * n_sm;
  if (live >= 8) return lo * n_sm;
  return std::max(lo, hi - ((live - 1) * (hi - lo) + 3) / 7) * n_sm;
}

// The by-value scalar block of k_move / k_trials.
void fill_move_dev(const pg_engine* h, PgMoveDev& D) {
  const PgDev& P = h->P;
  for (int i = 0; i < 3; i++) {
    D.box[i] = P.box[i]; D.inv_box[i] = P.inv_box[i]; D.ebox[i] = P.ebox[i]; D.inv_ebox[i] = P.inv_ebox[i];
    D.pbc[i] = P.pbc[i];
    D.fbox[i] = (float)P.box[i];
    D.half_box[i] = P.half_box[i]; D.same_box[i] = P.same_box[i];
  }
  D.rc2_relaxed = P.rc2_relaxed; D.ljc2max = P.lj_rcut2_relaxed_max; D.recip_pref = P.recip_pref;
  D.dipole_pref = P.dipole_pref; D.beta = P.beta;
  D.real_cutoff = P.real_cutoff; D.rc_relaxed = P.rc_relaxed; D.lB = P.lB; D.sqrt_alpha = P.sqrt_alpha;
  const bool multi = P.use_ewald && !P.single_image;
  D.img_cx = multi ? P.real_cell[0] : 0; D.img_cy = multi ? P.real_cell[1] : 0; D.img_cz = multi ? P.real_cell[2] : 0;
  D.img_ny = 2 * D.img_cy + 1;
  D.img_split = (2 * D.img_cx + 1) * D.img_ny;
  D.pair_kind = P.pair_kind; D.use_ewald = P.use_ewald; D.dipole = P.dipole; D.bond_kind = P.bond_kind;
  D.ext_kind = P.ext_kind;
}

int move_grid(pg_engine* h, int glen, int nq, PgMoveArgs& A) {
  fill_move_dev(h, A.D);
  A.n_tiles = (int)std::max<long long>(1, ((long long)h->n * A.D.img_split + MV_THREADS - 1) / MV_THREADS);
  const int slots = move_slots_now(h), n_sm = std::max(1, h->n_sm);
  const long long units = (long long)A.n_tiles * std::max(glen, 1);
  const long long npairs = (long long)glen * (glen - 1) / 2;
  if (units >= (1LL << 31) || npairs * A.D.img_split >= (1LL << 31)) { h->err = "move too large for 32-bit work indexing"; return -1; }
  const int nk = (h->P.use_ewald ? h->nk : 0);
  // generic path: 2 CTAs/SM of multi-image work per unit, so helpers never get an SM each
  const bool big = h->fast && units >= (long long)(slots - n_sm);
  int lpk = lanes_per_k(nq);
  int helpers;
  if (big) {
    helpers = n_sm;
  } else {
    // small move: as few helpers as hold the k slice (at >= 2 lanes per k) and the intra pairs
    helpers = (int)std::max<long long>(((long long)nk * lpk + MV_THREADS - 1) / MV_THREADS,
                                       (npairs * A.D.img_split + MV_THREADS - 1) / MV_THREADS);
    helpers = std::min(std::max(helpers, (nk > 0 || npairs > 0) ? 1 : 0), n_sm);
  }
  if (helpers > 0) {
    // shrink lanes-per-k until a helper's k slice fits one pass of its threads
    const int k_per_helper = (nk + helpers - 1) / helpers;
    while (lpk > 2 && k_per_helper * lpk > MV_THREADS) lpk >>= 1;
  }
  A.lpk = lpk;
  A.n_helpers = helpers;
  const long long room = std::max(1, slots - helpers);
  A.n_pair_ctas = (int)std::max(1LL, std::min(units, room));
  // Throughput mode (several replicas share the device): a pair CTA takes whole partner tiles and walks all the
  // moved beads, so the partner-side work (loads, box fractions, the warp's bounding extent) is done once per
  // tile instead of once per (tile, bead chunk) — fewer instructions per move at the price of a longer single
  // launch, which the other replicas' kernels hide.  A lone chain keeps the finest one-wave split.
  {
    static const int mode = [] { const char* e = getenv("PLUM_B200_TILE_CTAS"); return e ? atoi(e) : -1; }();
    const int live = (h->device >= 0 && h->device < 64) ? g_live_engines[h->device].load(std::memory_order_relaxed) : 1;
    const bool whole = (mode < 0) ? (live >= 4) : (mode != 0);
    if (whole && big && glen > MV_GCHUNK && A.n_tiles > 0) {
      static const int tiles_per_cta = [] { const char* e = getenv("PLUM_B200_TILES_PER_CTA"); return e ? std::max(1, atoi(e)) : 1; }();
      // (measured, tools/sweep_tile_ctas.sh, 24 replicas: helpers n_sm/1 -> 412 k, /4 -> 460 k, /8 -> 461 k, /16 -> 464 k, /32 -> 446 k moves/s)
      static const int helpers_div = [] { const char* e = getenv("PLUM_B200_HELPERS_DIV"); return e ? std::max(1, atoi(e)) : 8; }();
      if (helpers_div > 1) {
        helpers = std::max(1, n_sm / helpers_div);
        int l = lanes_per_k(nq);
        const int k_per_helper = (nk + helpers - 1) / helpers;
        while (l > 2 && k_per_helper * l > MV_THREADS) l >>= 1;
        A.lpk = l;
        A.n_helpers = helpers;
      }
      const long long room2 = std::max(1, slots - helpers);
      const long long per = std::max<long long>((A.n_tiles + room2 - 1) / room2, tiles_per_cta);
      A.n_pair_ctas = (int)std::max(1LL, (A.n_tiles + per - 1) / per);
    } else if (whole && h->fast && glen <= MV_GCHUNK) {
      // small moves (every single-bead move): what limits the throughput of many co-running replicas is CTA-slot
      // time, and a CTA of a small move is nearly all fixed latency — so fewer CTAs, each walking several tiles
      // (measured, 24 replicas: 1 tile per CTA 465 k, 2 -> 491 k, 4 -> 500 k moves/s device-resident; the host-driven
      // rate is flat to slightly lower beyond 2 because a single launch gets longer)
      static const int small_tiles = [] { const char* e = getenv("PLUM_B200_SMALL_TILES"); return e ? std::max(1, atoi(e)) : 2; }();
      static const int small_hdiv = [] { const char* e = getenv("PLUM_B200_SMALL_HELPERS_DIV"); return e ? std::max(1, atoi(e)) : 4; }();
      if (small_hdiv > 1 && helpers > 1) {
        helpers = std::max(1, (helpers + small_hdiv - 1) / small_hdiv);
        A.n_helpers = helpers;
      }
      const long long u_per = (long long)small_tiles * std::max(glen, 1);
      A.n_pair_ctas = (int)std::max(1LL, std::min<long long>((units + u_per - 1) / u_per, std::max(1, slots - helpers)));
    }
  }
  return A.n_helpers + A.n_pair_ctas;
}

// One launch per move: energy change of the trial in `d_group` + deferred commit of h->pc.
int launch_move(pg_engine* h, int g0, int glen, int nq, const char* d_group, int group_cap, bool decide_on_device,
                double u, int replay_index, bool want_result, bool log_replay, const char* h_inline = nullptr) {
  PgMoveArgs A;
  memset(&A, 0, sizeof(A));
  A.xy = h->xy; A.zq = h->zq; A.type = h->type; A.n = h->n;
  A.g0 = g0; A.glen = glen;
  StageView dv = stage_view(const_cast<char*>(d_group), group_cap);
  A.trial = dv.trial; A.gq = dv.gq; A.gtype = dv.gtype; A.moved = dv.moved;
  A.qidx = dv.qidx; A.nq = h->P.use_ewald ? nq : 0;
  if (h_inline) {
    // small group: the trial data ride in the kernel parameters (no H2D copy before the launch)
    StageView hv = stage_view(const_cast<char*>(h_inline), group_cap);
    A.inl_n = glen;
    for (int i = 0; i < glen; i++) {
      A.inl_trial[3 * i] = hv.trial[3 * i]; A.inl_trial[3 * i + 1] = hv.trial[3 * i + 1]; A.inl_trial[3 * i + 2] = hv.trial[3 * i + 2];
      A.inl_gq[i] = hv.gq[i]; A.inl_gtype[i] = hv.gtype[i]; A.inl_qidx[i] = hv.qidx[i]; A.inl_moved[i] = hv.moved[i];
    }
  }
  if (h->pc.valid) {
    StageView pv = stage_view(const_cast<char*>(h->pc.d_group), h->pc.cap);
    A.prev_valid = 1; A.prev_accept = h->pc.accept; A.pg0 = h->pc.g0; A.pglen = h->pc.glen; A.ptrial = pv.trial;
    if (!h->pc.on_device) {
      StageView hv = stage_view(const_cast<char*>(h->pc.h_group), h->pc.cap);
      A.pinl_n = h->pc.glen;
      for (int i = 0; i < 3 * h->pc.glen; i++) A.pinl_trial[i] = hv.trial[i];
    }
  }
  A.kvec = h->d_kvec; A.S = h->d_S; A.dS = h->d_dS; A.nk = h->P.use_ewald ? h->nk : 0;
  const int n_ctas = move_grid(h, glen, A.nq, A);
  if (n_ctas < 0) return PG_ERR_CAPACITY;
  int rc = ensure_partials(h, n_ctas);
  if (rc) return rc;
  A.partial = h->d_partial;
  A.state = h->d_state;
  A.mail = want_result ? h->d_mail : nullptr;
  A.decide_on_device = decide_on_device ? 1 : 0;
  A.u = u;
  A.replay_dE = log_replay ? h->d_rp_dE : nullptr;
  A.replay_acc = log_replay ? h->d_rp_acc : nullptr;
  A.replay_index = replay_index;
  A.seq = ++h->seq;
  A.Pg = h->d_P;
  A.timing = (h->d_timing && n_ctas <= 8192) ? h->d_timing : nullptr;
  h->timing_ctas = n_ctas;
  if (h->fast)
    k_move<true><<<n_ctas, MV_THREADS, 0, h->stream>>>(A);
  else
    k_move<false><<<n_ctas, MV_THREADS, 0, h->stream>>>(A);
  h->launches++;
  h->pc.valid = false;   // the launch applies it
  PG_CUDA(h, cudaGetLastError());
  return PG_OK;
}

// Apply a commit that no k_move launch has picked up yet (needed before anything else reads
// positions, S(k) or the totals).
int flush_commit(pg_engine* h) {
  if (!h->pc.valid) return PG_OK;
  h->pc.valid = false;
  if (!h->pc.on_device)
    PG_CUDA(h, cudaMemcpyAsync(const_cast<char*>(h->pc.d_group), h->pc.h_group, stage_size(h->pc.cap), cudaMemcpyHostToDevice,
                               h->stream));
  return launch_commit(h, h->pc.accept, PG_MODE_MOVE, h->pc.g0, h->pc.glen, h->pc.d_group, h->pc.cap);
}

// Spin on k_move's mailbox: every 16-byte record carries the launch's sequence number, so a
// record whose seq matches is complete (no device-side system fence on the critical path).
int wait_mail(pg_engine* h, unsigned int seq, pg_delta* o) {
  volatile PgMailRec* m = h->h_mail;
  SpinWait w;
  double val[MV_NSLOT];
  int aux[MV_NSLOT];
  for (int s = 0; s < MV_NSLOT; s++) {
    for (;;) {
      if (__atomic_load_n(&h->h_mail[s].seq, __ATOMIC_ACQUIRE) == seq) {
        // (acquire: on a weakly ordered host the payload loads must not be satisfied before the seq load)
        val[s] = m[s].value;
        aux[s] = m[s].aux;
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        if (m[s].seq == seq) break;
      }
      int rc = w.step(h);
      if (rc) return rc;
    }
  }
  o->dE = val[0]; o->pair = val[1]; o->ext = val[2]; o->ewald = val[3]; o->bond = val[4];
  o->real = val[5]; o->recip = val[6]; o->mz_current = val[7];
  o->stage = aux[0] & 0xff; o->n_overlap = aux[1];
  return PG_OK;
}


// Compact list of the charged beads, from the host mirror of the charges.
int ensure_qidx(pg_engine* h) {
  if (h->qidx_valid) return PG_OK;
  std::vector<int> idx;
  for (int i = 0; i < h->n; i++)
    if (h->h_q[i] != 0.0) idx.push_back(i);
  if (idx.size() > h->qidx_cap || !h->d_qidx) {
    cudaFree(h->d_qidx); h->d_qidx = nullptr; h->qidx_cap = 0;
    const size_t cap = std::max<size_t>(idx.size() * 2, 1024);
    PG_CUDA(h, cudaMalloc((void**)&h->d_qidx, sizeof(int) * cap));
    h->qidx_cap = cap;
  }
  if (!idx.empty()) {
    PG_CUDA(h, cudaMemcpyAsync(h->d_qidx, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->nq_tot = (int)idx.size();
  h->qidx_valid = true;
  return PG_OK;
}

// S(k) of the k slice [k_first, k_first + k_count) into `out` (device): k tiles x chunks of the charged list, then the
// ordered reduction over the chunks (pg_sk.cu).  The chunking is a function of the number of charged beads only, so a
// slice computed alone is bit-identical to the same rows of the full result.  With `exchange` the reduction also stores
// the slice into the attached peers' S(k) and handshakes with them.
int launch_sk_slice(pg_engine* h, int k_first, int k_count, double2* out, bool exchange = false) {
  if (k_count <= 0 && !exchange) return PG_OK;
  int rc = ensure_qidx(h);
  if (rc) return rc;
  const int nq = h->nq_tot;
  const int chunk = SK_BEADS * std::max(1, (nq + SK_BEADS * 256 - 1) / (SK_BEADS * 256));
  const int n_chunks = std::max(1, (nq + chunk - 1) / chunk);
  const size_t need = (size_t)n_chunks * (size_t)std::max(k_count, 1);
  if (need > h->sk_partial_cap) {
    cudaFree(h->d_sk_partial); h->d_sk_partial = nullptr; h->sk_partial_cap = 0;
    PG_CUDA(h, cudaMalloc((void**)&h->d_sk_partial, sizeof(double2) * need));
    h->sk_partial_cap = need;
  }
  PgSkArgs A;
  memset(&A, 0, sizeof(A));
  A.xy = h->xy; A.zq = h->zq; A.qidx = h->d_qidx; A.nq = nq;
  A.k_first = k_first; A.k_count = k_count;
  // the items that overlap the slice (items are sorted by their first k)
  int i0 = 0, i1 = (int)h->h_sk_items.size();
  while (i0 < i1 && h->h_sk_items[i0].k0 + h->h_sk_item_n[i0] <= k_first) i0++;
  while (i1 > i0 && h->h_sk_items[i1 - 1].k0 >= k_first + k_count) i1--;
  A.items = h->d_sk_items; A.item_n = h->d_sk_item_n; A.item_first = i0; A.item_count = i1 - i0;
  for (int a = 0; a < 3; a++) {
    A.kmax[a] = h->sk_kmax[a]; A.kunit[a] = h->sk_kunit[a];
    A.ebox[a] = h->P.ebox[a]; A.inv_ebox[a] = h->P.inv_ebox[a]; A.pbc[a] = h->P.pbc[a];
  }
  A.chunk = chunk; A.partial = h->d_sk_partial;
  const int ne = (h->sk_kmax[0] + 1) + (2 * h->sk_kmax[1] + 1) + (2 * h->sk_kmax[2] + 1);
  const size_t smem = sizeof(double2) * (size_t)SK_BEADS * ne;
  if (smem > 200 * 1024) { h->err = "S(k) recompute: k table does not fit shared memory"; return PG_ERR_CAPACITY; }
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    PG_CUDA(h, cudaFuncSetAttribute(k_sk_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  if (k_count > 0 && A.item_count > 0) {
    const int kt = (A.item_count + SK_ITHREADS - 1) / SK_ITHREADS;
    k_sk_block<<<dim3(kt, n_chunks), SK_ITHREADS, smem, h->stream>>>(A);
    h->launches++;
  }
  PgSkPeers Pe = h->peers;
  if (!exchange) Pe.world = 0;
  k_sk_finish<<<std::max(1, (k_count + SK_THREADS - 1) / SK_THREADS), SK_THREADS, 0, h->stream>>>(
      h->d_sk_partial, n_chunks, k_count, k_first, out, Pe, h->d_sk_counter);
  h->launches++;
  PG_CUDA(h, cudaGetLastError());
  return PG_OK;
}

int compute_totals(pg_engine* h, bool set_state, pg_totals* out) {
  int frc = flush_commit(h);
  if (frc) return frc;
  double2* S = set_state ? h->d_S : h->d_Stmp;
  if (h->P.use_ewald && h->nk > 0) {
    int src = launch_sk_slice(h, 0, h->nk, S);
    if (src) return src;
  }
  // the N^2 triangle as (row tile) x (partner slice) CTAs: slices of PG_TILE partners, narrower for small systems so
  // that the examples (a few hundred beads) still occupy the machine
  const int n_rt = (h->n + PG_TILE - 1) / PG_TILE;
  int pc = PG_TILE;
  while (pc > 8 && (long long)n_rt * ((h->n + pc - 1) / pc) < 4LL * std::max(h->n_sm, 1)) pc >>= 1;
  const int n_ct = (h->n + pc - 1) / pc;
  const int n_tiles = n_rt * n_ct;
  int rc = ensure_partials(h, n_tiles);
  if (rc) return rc;
  if (n_tiles > 0) {
    k_tot_pairs<<<dim3((unsigned)n_rt, (unsigned)n_ct), PG_TILE, 0, h->stream>>>(h->P, h->xy, h->zq, h->type, h->mol, h->n, pc, h->d_partial);
    h->launches++;
    PG_CUDA(h, cudaGetLastError());
  }
  k_tot_final<<<1, 256, 0, h->stream>>>(h->P, h->xy, h->zq, h->type, h->mol, h->n, h->d_partial, n_tiles, S,
                                        h->d_ek2, h->nk, h->d_out8, h->d_state, set_state ? 1 : 0);
  h->launches++;
  PG_CUDA(h, cudaGetLastError());
  PG_CUDA(h, cudaMemcpyAsync(h->h_out8, h->d_out8, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (out) {
    out->pair = h->h_out8[0]; out->ewald = h->h_out8[1]; out->bond = h->h_out8[2]; out->ext = h->h_out8[3];
    out->real = h->h_out8[4]; out->recip = h->h_out8[5]; out->self = h->h_out8[6]; out->dipole = h->h_out8[7];
  }
  return PG_OK;
}

void chain_free(pg_engine* h);

void free_all(pg_engine* h) {
  chain_free(h);
  cudaFree(h->xy); cudaFree(h->zq); cudaFree(h->type); cudaFree(h->mol);
  cudaFree(h->t_xy); cudaFree(h->t_zq); cudaFree(h->t_type); cudaFree(h->t_mol);
  for (int r = 0; r < SK_MAX_PEERS; r++)
    if (h->ipc_opened[r]) cudaIpcCloseMemHandle(h->ipc_opened[r]);
  cudaFree(h->d_kl); cudaFree(h->d_ek2); cudaFree(h->d_kvec); cudaFree(h->d_xchg); cudaFree(h->d_dS); cudaFree(h->d_Stmp);
  cudaFree(h->d_qidx); cudaFree(h->d_sk_counter); cudaFree(h->d_sk_items); cudaFree(h->d_sk_item_n);
  for (int s = 0; s < 2; s++) {
    if (h->h_stage2[s]) cudaFreeHost(h->h_stage2[s]);
    cudaFree(h->d_stage2[s]);
  }
  cudaFree(h->d_partial); cudaFree(h->d_partial_i); cudaFree(h->d_out8);
  if (h->h_out8) cudaFreeHost(h->h_out8);
  cudaFree(h->d_state);
  cudaFree(h->d_P);
  if (h->h_result) cudaFreeHost(h->h_result);
  if (h->h_mail) cudaFreeHost(h->h_mail);
  if (h->h_trial_in) cudaFreeHost(h->h_trial_in);
  if (h->h_trial_ct) cudaFreeHost(h->h_trial_ct);
  if (h->h_tmail) cudaFreeHost(h->h_tmail);
  cudaFree(h->d_trial_in); cudaFree(h->d_trial_ct); cudaFree(h->d_tr_partial); cudaFree(h->d_sk_partial); cudaFree(h->d_tr_mz); cudaFree(h->d_tr_counter);
  cudaFree(h->d_rp); cudaFree(h->d_rp_dE); cudaFree(h->d_rp_acc);
  cudaFree(h->d_mc_moves); cudaFree(h->d_mc_rv); cudaFree(h->d_stop);
  if (h->h_mc_pin) cudaFreeHost(h->h_mc_pin);
  if (h->h_stop) cudaFreeHost(h->h_stop);
  if (h->h_mc_dE) cudaFreeHost(h->h_mc_dE);
  if (h->h_mc_acc) cudaFreeHost(h->h_mc_acc);
  for (auto& g : h->rp_graphs) cudaGraphExecDestroy(g.exec);
  h->rp_graphs.clear();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
}

}  // namespace

// =============================================================== C ABI ====
extern "C" {

int pg_abi_version(void) { return PG_ABI_VERSION; }

const char* pg_last_error(const pg_engine* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int pg_create(const pg_params* params, int device, int capacity_beads, pg_engine** out) {
  if (!params || !out) { g_create_err = "null argument"; return PG_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    g_create_err = std::string("no usable CUDA device (there is no CPU fallback): ") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
    return PG_ERR_NO_DEVICE;
  }
  pg_engine* h = new pg_engine();
  h->device = device;
  int rc = build_params(h, params);
  if (rc) { delete h; return rc; }
  {
    // k_move<true>: central image only, one box for LJ and Ewald, three periodic axes, and every
    // LJ cutoff well inside half the cell so BBDist's fold can only act beyond it.
    const PgDev& P = h->P;
    double lmin = std::min(P.box[0], std::min(P.box[1], P.box[2]));
    bool lj_ok = true;
    if (P.pair_kind == PG_PAIR_TRUNCATED_LJ)
      for (int a = 0; a < P.n_types; a++)
        for (int b = 0; b < P.n_types; b++)
          if (!(P.lj_rcut[a * PG_MAX_TYPES + b] < 0.49 * lmin)) lj_ok = false;
    h->fast = (params->npbc == 3) && P.pair_kind != PG_PAIR_HARD_SPHERE && lj_ok && P.same_box[0] && P.same_box[1] &&
              P.same_box[2] && (!P.use_ewald || P.single_image);
    const char* nf = getenv("PLUM_B200_NO_FAST");
    if (nf && nf[0] == '1') h->fast = false;
  }
#define PG_CREATE_CUDA(call)                                                    \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      g_create_err = std::string(#call) + ": " + cudaGetErrorString(e_);        \
      free_all(h); delete h; return PG_ERR_CUDA;                                \
    }                                                                           \
  } while (0)
  PG_CREATE_CUDA(cudaSetDevice(device));
  PG_CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  // blocking sync: a host thread waiting for a replay batch sleeps instead of spinning (many replicas per GPU)
  {
    const char* es = getenv("PLUM_B200_EVENT_SPIN");
    const unsigned ef = (es && es[0] == '1') ? cudaEventDefault : cudaEventBlockingSync;
    PG_CREATE_CUDA(cudaEventCreateWithFlags(&h->ev0, ef));
    PG_CREATE_CUDA(cudaEventCreateWithFlags(&h->ev1, ef));
  }
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_state, sizeof(PgState)));
  PG_CREATE_CUDA(cudaMemset(h->d_state, 0, sizeof(PgState)));
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_P, sizeof(PgDev)));
  PG_CREATE_CUDA(cudaMemcpy(h->d_P, &h->P, sizeof(PgDev), cudaMemcpyHostToDevice));
  PG_CREATE_CUDA(cudaHostAlloc((void**)&h->h_result, sizeof(PgResult), cudaHostAllocMapped));
  memset((void*)h->h_result, 0, sizeof(PgResult));
  PG_CREATE_CUDA(cudaHostGetDevicePointer((void**)&h->d_result, (void*)h->h_result, 0));
  PG_CREATE_CUDA(cudaHostAlloc((void**)&h->h_mail, sizeof(PgMailRec) * MV_NSLOT, cudaHostAllocMapped));
  memset((void*)h->h_mail, 0, sizeof(PgMailRec) * MV_NSLOT);
  PG_CREATE_CUDA(cudaHostGetDevicePointer((void**)&h->d_mail, (void*)h->h_mail, 0));
  {
    int per_sm = 0, n_sm = 0;
    if (h->fast)
      PG_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_move<true>, MV_THREADS, 0));
    else
      PG_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_move<false>, MV_THREADS, 0));
    PG_CREATE_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    // how many of them one k_move grid takes is decided per launch (move_slots_now)
    h->move_per_sm = std::max(1, per_sm);
    if (const char* e = getenv("PLUM_B200_CTAS_PER_SM")) h->move_per_sm_fixed = std::max(1, std::min(per_sm, atoi(e)));
    h->move_slots = std::max(1, per_sm * n_sm);
    h->n_sm = n_sm;
  }
  {
    int per_sm = 0;
    PG_CREATE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trials, TR_THREADS, 0));
    h->trial_slots = std::max(1, per_sm) * std::max(1, h->n_sm);
    PG_CREATE_CUDA(cudaHostAlloc((void**)&h->h_tmail, sizeof(PgMailRec) * 3 * TR_MAXT, cudaHostAllocMapped));
    memset((void*)h->h_tmail, 0, sizeof(PgMailRec) * 3 * TR_MAXT);
    PG_CREATE_CUDA(cudaHostGetDevicePointer((void**)&h->d_tmail, (void*)h->h_tmail, 0));
    PG_CREATE_CUDA(cudaMalloc((void**)&h->d_tr_counter, sizeof(unsigned int)));
    PG_CREATE_CUDA(cudaMemset(h->d_tr_counter, 0, sizeof(unsigned int)));
  }
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_out8, sizeof(double) * 8));
  PG_CREATE_CUDA(cudaHostAlloc((void**)&h->h_out8, sizeof(double) * 8, cudaHostAllocDefault));
  int nk_alloc = std::max(h->nk, 1);
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_kl, sizeof(int) * 4 * (size_t)nk_alloc));
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_ek2, sizeof(double) * (size_t)nk_alloc));
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_kvec, sizeof(double4) * (size_t)nk_alloc));
  {
    const size_t s_bytes = sizeof(double2) * (size_t)nk_alloc;
    PG_CREATE_CUDA(cudaMalloc((void**)&h->d_xchg, s_bytes + sizeof(unsigned int) * 2 * SK_MAX_PEERS));
    PG_CREATE_CUDA(cudaMemset(h->d_xchg, 0, s_bytes + sizeof(unsigned int) * 2 * SK_MAX_PEERS));
    h->d_S = reinterpret_cast<double2*>(h->d_xchg);
    h->d_flags = reinterpret_cast<unsigned int*>(h->d_xchg + s_bytes);
    PG_CREATE_CUDA(cudaMalloc((void**)&h->d_sk_counter, sizeof(unsigned int) * 4));
    PG_CREATE_CUDA(cudaMemset(h->d_sk_counter, 0, sizeof(unsigned int) * 4));
    memset(&h->peers, 0, sizeof(h->peers));
    for (int a = 0; a < 3; a++) { h->sk_kmax[a] = 0; h->sk_kunit[a] = 1 * 2 * kPi / h->P.ebox[a]; }
    for (int k = 0; k < h->nk; k++)
      for (int a = 0; a < 3; a++) h->sk_kmax[a] = std::max(h->sk_kmax[a], std::abs(h->h_kl[4 * k + a]));
    // work items of the full recompute: runs of the k list with equal (lx, ly) and consecutive lz, at most SK_SEG long
    for (int k = 0; k < h->nk;) {
      const int lx = h->h_kl[4 * k], ly = h->h_kl[4 * k + 1], lz0 = h->h_kl[4 * k + 2];
      int n = 1;
      while (k + n < h->nk && n < SK_SEG && h->h_kl[4 * (k + n)] == lx && h->h_kl[4 * (k + n) + 1] == ly &&
             h->h_kl[4 * (k + n) + 2] == lz0 + n) n++;
      PgSkItem it; it.lx = lx; it.ly = ly; it.lz0 = lz0; it.k0 = k;
      h->h_sk_items.push_back(it);
      h->h_sk_item_n.push_back(n);
      k += n;
    }
    const size_t ni = std::max<size_t>(h->h_sk_items.size(), 1);
    PG_CREATE_CUDA(cudaMalloc((void**)&h->d_sk_items, sizeof(PgSkItem) * ni));
    PG_CREATE_CUDA(cudaMalloc((void**)&h->d_sk_item_n, sizeof(int) * ni));
    if (!h->h_sk_items.empty()) {
      PG_CREATE_CUDA(cudaMemcpy(h->d_sk_items, h->h_sk_items.data(), sizeof(PgSkItem) * h->h_sk_items.size(), cudaMemcpyHostToDevice));
      PG_CREATE_CUDA(cudaMemcpy(h->d_sk_item_n, h->h_sk_item_n.data(), sizeof(int) * h->h_sk_item_n.size(), cudaMemcpyHostToDevice));
    }
  }
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_dS, sizeof(double2) * (size_t)nk_alloc));
  PG_CREATE_CUDA(cudaMalloc((void**)&h->d_Stmp, sizeof(double2) * (size_t)nk_alloc));
  PG_CREATE_CUDA(cudaMemset(h->d_S, 0, sizeof(double2) * (size_t)nk_alloc));
  PG_CREATE_CUDA(cudaMemset(h->d_dS, 0, sizeof(double2) * (size_t)nk_alloc));
  if (h->nk > 0) {
    PG_CREATE_CUDA(cudaMemcpy(h->d_kl, h->h_kl.data(), sizeof(int) * 4 * (size_t)h->nk, cudaMemcpyHostToDevice));
    PG_CREATE_CUDA(cudaMemcpy(h->d_ek2, h->h_ek2.data(), sizeof(double) * (size_t)h->nk, cudaMemcpyHostToDevice));
    PG_CREATE_CUDA(cudaMemcpy(h->d_kvec, h->h_kvec.data(), sizeof(double4) * (size_t)h->nk, cudaMemcpyHostToDevice));
  }
#undef PG_CREATE_CUDA
  {
    const char* tm = getenv("PLUM_B200_TIMING");
    if (tm && tm[0] == '1') {
      if (cudaMalloc((void**)&h->d_timing, sizeof(unsigned long long) * 8 * 8192) != cudaSuccess) h->d_timing = nullptr;
    }
  }
  {
    const char* ng = getenv("PLUM_B200_NO_GRAPH");
    h->use_graph = !(ng && ng[0] == '1');
  }
  {
    const char* ni = getenv("PLUM_B200_NO_INLINE");
    h->use_inline = !(ni && ni[0] == '1');
  }
  const char* mb = getenv("PLUM_B200_NO_MAILBOX");
  h->use_mailbox = !(mb && mb[0] == '1');
  h->mol_first.assign(1, 0);
  rc = ensure_capacity(h, std::max(capacity_beads, 1));
  if (!rc) rc = ensure_group(h, 256);
  if (!rc) rc = ensure_partials(h, 4096);
  if (rc) { g_create_err = h->err; free_all(h); delete h; return rc; }
  if (device >= 0 && device < 64) g_live_engines[device].fetch_add(1);
  *out = h;
  return PG_OK;
}

int pg_destroy(pg_engine* h) {
  if (!h) return PG_OK;
  if (h->device >= 0 && h->device < 64) g_live_engines[h->device].fetch_sub(1);
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  free_all(h);
  delete h;
  return PG_OK;
}

int pg_get_ewald_info(const pg_engine* h, pg_ewald_info* out) {
  if (!h || !out) return PG_ERR_INVALID;
  *out = h->info;
  return PG_OK;
}

int pg_num_beads(const pg_engine* h) { return h ? h->n : PG_ERR_INVALID; }
uint64_t pg_launch_count(const pg_engine* h) { return h ? h->launches : 0; }
void* pg_stream(const pg_engine* h) { return h ? (void*)h->stream : nullptr; }

int pg_upload_system(pg_engine* h, int n_beads, const double* xyz, const double* q, const int32_t* type, int n_mol,
                     const int32_t* mol_first) {
  if (!h || n_beads < 0 || n_mol < 0 || !mol_first) return PG_ERR_INVALID;
  if (n_beads > 0 && (!xyz || !q || !type)) { h->err = "null bead arrays"; return PG_ERR_INVALID; }
  if (mol_first[0] != 0 || mol_first[n_mol] != n_beads) { h->err = "mol_first must span [0, n_beads]"; return PG_ERR_INVALID; }
  for (int i = 0; i < n_beads; i++)
    if (type[i] < 0 || type[i] >= h->P.n_types) { h->err = "bead type id out of range"; return PG_ERR_INVALID; }
  PG_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_capacity(h, std::max(n_beads, 1));
  if (rc) return rc;
  std::vector<double2> hxy(std::max(n_beads, 1)), hzq(std::max(n_beads, 1));
  std::vector<int> hmol(std::max(n_beads, 1));
  h->h_q.assign(q, q + n_beads);
  h->h_type.assign(type, type + n_beads);
  h->mol_first.assign(mol_first, mol_first + n_mol + 1);
  for (int m = 0; m < n_mol; m++) {
    if (mol_first[m + 1] < mol_first[m]) { h->err = "mol_first must be non-decreasing"; return PG_ERR_INVALID; }
    for (int i = mol_first[m]; i < mol_first[m + 1]; i++) hmol[i] = m;
  }
  for (int i = 0; i < n_beads; i++) {
    hxy[i] = make_double2(xyz[3 * i], xyz[3 * i + 1]);
    hzq[i] = make_double2(xyz[3 * i + 2], q[i]);
  }
  if (n_beads > 0) {
    PG_CUDA(h, cudaMemcpyAsync(h->xy, hxy.data(), sizeof(double2) * (size_t)n_beads, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(h->zq, hzq.data(), sizeof(double2) * (size_t)n_beads, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(h->type, type, sizeof(int) * (size_t)n_beads, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(h->mol, hmol.data(), sizeof(int) * (size_t)n_beads, cudaMemcpyHostToDevice, h->stream));
  }
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n = n_beads;
  h->n_mol = n_mol;
  h->pending = false;
  h->pc.valid = false;
  h->ch.valid = false;
  h->qidx_valid = false;
  replay_invalidate(h, true);
  return PG_OK;
}

int pg_init_energy(pg_engine* h, pg_totals* out) {
  if (!h) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->pending = false;
  return compute_totals(h, true, out);
}

int pg_recompute_totals(pg_engine* h, pg_totals* out) {
  if (!h) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  return compute_totals(h, false, out);
}

int pg_get_totals(const pg_engine* hc, pg_totals* out) {
  pg_engine* h = const_cast<pg_engine*>(hc);
  if (!h || !out) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  PgState st;
  PG_CUDA(h, cudaMemcpyAsync(&st, h->d_state, sizeof(PgState), cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  out->pair = st.E_pair; out->ewald = st.E_ewald; out->bond = st.E_bond; out->ext = st.E_ext;
  out->real = st.E_real; out->recip = st.E_recip; out->self = st.E_self; out->dipole = st.cur_dipl;
  return PG_OK;
}

int pg_download_positions(pg_engine* h, double* xyz) {
  if (!h || (!xyz && h->n > 0)) return PG_ERR_INVALID;
  if (h->n == 0) return PG_OK;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  std::vector<double2> hxy(h->n), hzq(h->n);
  PG_CUDA(h, cudaMemcpyAsync(hxy.data(), h->xy, sizeof(double2) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaMemcpyAsync(hzq.data(), h->zq, sizeof(double2) * (size_t)h->n, cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < h->n; i++) { xyz[3 * i] = hxy[i].x; xyz[3 * i + 1] = hxy[i].y; xyz[3 * i + 2] = hzq[i].x; }
  return PG_OK;
}

// ------------------------------------------------------------ per-move path
// Asynchronous half of pg_delta_e: stages the trial and launches k_move, returns without waiting.
int pg_delta_e_begin(pg_engine* h, int mol, const double* trial_xyz, const uint8_t* moved) {
  if (!h || !trial_xyz || !moved) return PG_ERR_INVALID;
  if (mol < 0 || mol >= h->n_mol) { h->err = "molecule index out of range"; return PG_ERR_INVALID; }
  if (h->pending || h->inflight) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  if (h->mc_inflight || h->ch.inflight) { h->err = "a batch is still in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  const int g0 = h->mol_first[mol], glen = h->mol_first[mol + 1] - g0;
  int rc = ensure_group(h, glen);
  if (rc) return rc;
  next_slot(h);   // the other slot still holds the previous trial (deferred commit)
  // compact layout: only stage_size(glen) bytes cross PCIe
  StageView sv = stage_view(h->h_stage, glen);
  memcpy(sv.trial, trial_xyz, sizeof(double) * 3 * (size_t)glen);
  for (int i = 0; i < glen; i++) {
    sv.gq[i] = h->h_q[g0 + i];
    sv.gtype[i] = h->h_type[g0 + i];
    sv.moved[i] = moved[i] ? 1 : 0;
  }
  const int nq = stage_fill_qidx(sv, 0, glen);
  const bool inl = (glen <= MV_INL) && h->use_inline;
  if (!inl) PG_CUDA(h, cudaMemcpyAsync(h->d_stage, h->h_stage, stage_size(glen), cudaMemcpyHostToDevice, h->stream));
  rc = launch_move(h, g0, glen, nq, h->d_stage, glen, false, 0.0, 0, true, false, inl ? h->h_stage : nullptr);
  if (rc) return rc;
  h->inflight = true;
  h->pend_mode = PG_MODE_MOVE; h->pend_g0 = g0; h->pend_glen = glen; h->pend_group = h->d_stage; h->pend_cap = glen;
  h->pend_hgroup = h->h_stage; h->pend_on_device = !inl;
  return PG_OK;
}

// Second half: 1 while the device has not answered (returns at once), 0 when `out` is filled and the
// trial is pending its pg_commit, < 0 on error.
int pg_delta_e_poll(pg_engine* h, pg_delta* out) {
  if (!h || !out) return PG_ERR_INVALID;
  if (!h->inflight) { h->err = "no trial in flight"; return PG_ERR_STATE; }
  for (int s = MV_NSLOT - 1; s >= 0; s--)
    if (__atomic_load_n(&h->h_mail[s].seq, __ATOMIC_ACQUIRE) != h->seq) return 1;
  int rc = wait_mail(h, h->seq, out);
  if (rc) return rc;
  h->inflight = false;
  h->pending = true;
  return PG_OK;
}

int pg_delta_e(pg_engine* h, int mol, const double* trial_xyz, const uint8_t* moved, pg_delta* out) {
  if (!out) return PG_ERR_INVALID;
  int rc = pg_delta_e_begin(h, mol, trial_xyz, moved);
  if (rc) return rc;
  rc = wait_mail(h, h->seq, out);
  if (rc) { h->inflight = false; return rc; }
  h->inflight = false;
  h->pending = true;
  return PG_OK;
}

// FinalizeEnergies: records the decision; the next k_move launch (or the next call that needs
// the state) applies it on the device.
int pg_commit(pg_engine* h, int accept) {
  if (!h) return PG_ERR_INVALID;
  if (!h->pending) { h->err = "no pending trial"; return PG_ERR_STATE; }
  h->pending = false;
  if (accept) h->ch.valid = false;   // coordinates change outside the chain kernel: its resident structures are stale
  h->pc.valid = true;
  h->pc.accept = accept ? 1 : 0;
  h->pc.g0 = h->pend_g0; h->pc.glen = h->pend_glen; h->pc.d_group = h->pend_group; h->pc.cap = h->pend_cap;
  h->pc.h_group = h->pend_hgroup; h->pc.on_device = h->pend_on_device;
  return PG_OK;
}

// ------------------------------------------------------------------ replay
static void replay_drop_graphs(pg_engine* h);
int pg_replay_upload(pg_engine* h, int n_moves, const pg_proposal* moves, int n_xyz_beads, const double* trial_xyz,
                     const uint8_t* moved) {
  if (!h || n_moves < 0 || (n_moves > 0 && (!moves || !trial_xyz || !moved))) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  if (h->mc_inflight) { h->err = "an MC batch is in flight"; return PG_ERR_STATE; }
  h->rp_moves.clear();
  h->rp_mc = false;
  replay_drop_graphs(h);
  cudaFree(h->d_rp); cudaFree(h->d_rp_dE); cudaFree(h->d_rp_acc);
  h->d_rp = nullptr; h->d_rp_dE = nullptr; h->d_rp_acc = nullptr;
  h->rp_bytes_cap = 0; h->rp_log_cap = 0;
  h->rp_beads = (size_t)std::max(n_xyz_beads, 1);
  std::vector<char> host(stage_size((int)h->rp_beads));
  StageView sv = stage_view(host.data(), (int)h->rp_beads);
  memcpy(sv.trial, trial_xyz, sizeof(double) * 3 * (size_t)n_xyz_beads);
  for (int i = 0; i < n_xyz_beads; i++) sv.moved[i] = moved[i] ? 1 : 0;
  for (int m = 0; m < n_moves; m++) {
    int mol = moves[m].mol;
    if (mol < 0 || mol >= h->n_mol) { h->err = "replay: molecule index out of range"; return PG_ERR_INVALID; }
    ReplayMove r;
    r.mol = mol; r.g0 = h->mol_first[mol]; r.glen = h->mol_first[mol + 1] - r.g0; r.off = moves[m].xyz_offset;
    r.u = moves[m].u;
    if (r.off < 0 || r.off + r.glen > n_xyz_beads) { h->err = "replay: xyz_offset out of range"; return PG_ERR_INVALID; }
    for (int i = 0; i < r.glen; i++) { sv.gq[r.off + i] = h->h_q[r.g0 + i]; sv.gtype[r.off + i] = h->h_type[r.g0 + i]; }
    r.nq = stage_fill_qidx(sv, r.off, r.glen);
    h->rp_moves.push_back(r);
  }
  PG_CUDA(h, cudaMalloc((void**)&h->d_rp, host.size()));
  PG_CUDA(h, cudaMemcpy(h->d_rp, host.data(), host.size(), cudaMemcpyHostToDevice));
  PG_CUDA(h, cudaMalloc((void**)&h->d_rp_dE, sizeof(double) * (size_t)std::max(n_moves, 1)));
  PG_CUDA(h, cudaMalloc((void**)&h->d_rp_acc, (size_t)std::max(n_moves, 1)));
  h->rp_bytes_cap = host.size(); h->rp_log_cap = (size_t)std::max(n_moves, 1);
  return PG_OK;
}

// Device pointers of replay move m inside the packed proposal buffer (StageView layout with
// capacity rp_beads, every sub-array shifted by the move's bead offset).
static void replay_view(pg_engine* h, int m, StageView& v) {
  StageView base = stage_view(h->d_rp, (int)h->rp_beads);
  const ReplayMove& r = h->rp_moves[m];
  v.trial = base.trial + 3 * (size_t)r.off; v.gq = base.gq + r.off; v.gtype = base.gtype + r.off;
  v.qidx = base.qidx + r.off; v.moved = base.moved + r.off;
}

static int replay_launch(pg_engine* h, int m, int prev, bool log) {
  const ReplayMove& r = h->rp_moves[m];
  PgMoveArgs A;
  memset(&A, 0, sizeof(A));
  A.xy = h->xy; A.zq = h->zq; A.type = h->type; A.n = h->n;
  A.g0 = r.g0; A.glen = r.glen;
  StageView v;
  replay_view(h, m, v);
  A.trial = v.trial; A.gq = v.gq; A.gtype = v.gtype; A.moved = v.moved;
  A.qidx = v.qidx; A.nq = h->P.use_ewald ? r.nq : 0;
  if (prev >= 0) {
    const ReplayMove& pr = h->rp_moves[prev];
    StageView pv;
    replay_view(h, prev, pv);
    A.prev_valid = 1; A.prev_accept = -1; A.pg0 = pr.g0; A.pglen = pr.glen; A.ptrial = pv.trial;
  }
  A.kvec = h->d_kvec; A.S = h->d_S; A.dS = h->d_dS; A.nk = h->P.use_ewald ? h->nk : 0;
  const int n_ctas = move_grid(h, r.glen, A.nq, A);
  if (n_ctas < 0) return PG_ERR_CAPACITY;
  int rc = ensure_partials(h, n_ctas);
  if (rc) return rc;
  A.partial = h->d_partial; A.state = h->d_state; A.mail = nullptr;
  A.decide_on_device = 1; A.u = r.u;
  A.replay_dE = log ? (h->rp_mc ? h->d_mc_dE : h->d_rp_dE) : nullptr;
  A.replay_acc = log ? (h->rp_mc ? h->d_mc_acc : h->d_rp_acc) : nullptr;
  A.replay_index = m;
  A.stop = h->rp_mc ? h->d_stop : nullptr;
  A.seq = ++h->seq;
  A.Pg = h->d_P;
  A.timing = (h->d_timing && n_ctas <= 8192) ? h->d_timing : nullptr;
  h->timing_ctas = n_ctas;
  if (h->fast)
    k_move<true><<<n_ctas, MV_THREADS, 0, h->stream>>>(A);
  else
    k_move<false><<<n_ctas, MV_THREADS, 0, h->stream>>>(A);
  h->launches++;
  return PG_OK;
}

static void replay_drop_graphs(pg_engine* h) {
  for (auto& g : h->rp_graphs) cudaGraphExecDestroy(g.exec);
  h->rp_graphs.clear();
}

// Enqueue the launches of one replay batch on the engine stream (directly, or into a capture).
static int replay_enqueue(pg_engine* h, int first, int count, bool commit) {
  for (int m = first; m < first + count; m++) {
    int rc = replay_launch(h, m, (commit && m > first) ? m - 1 : -1, commit);
    if (rc) return rc;
  }
  if (commit && count > 0) {
    // the last move's decision (taken on the device) is applied by a stand-alone commit
    const ReplayMove& r = h->rp_moves[first + count - 1];
    StageView v;
    replay_view(h, first + count - 1, v);
    int nthreads = std::max(std::max(r.glen, h->nk), 1);
    k_commit<<<(nthreads + 255) / 256, 256, 0, h->stream>>>(h->P, -1, PG_MODE_MOVE, r.g0, r.glen, v.trial, v.gq,
                                                            v.gtype, h->xy, h->zq, h->type, h->d_S, h->d_dS,
                                                            h->P.use_ewald ? h->nk : 0, h->d_state);
    h->launches++;
  }
  return PG_OK;
}

// Returns the instantiated graph of a batch (captured on first use), or nullptr when graphs are off.
static int replay_graph(pg_engine* h, int first, int count, bool commit, cudaGraphExec_t* out) {
  *out = nullptr;
  if (!h->use_graph || count <= 0) return PG_OK;
  for (auto& g : h->rp_graphs)
    if (g.first == first && g.count == count && g.commit == (int)commit) { *out = g.exec; return PG_OK; }
  int rc = ensure_partials(h, h->move_slots + h->n_sm + 8);   // no allocation inside a capture
  if (rc) return rc;
  const uint64_t launches0 = h->launches;
  PG_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  rc = replay_enqueue(h, first, count, commit);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  h->launches = launches0;   // captured, not launched
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) { h->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return PG_ERR_CUDA; }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e); return PG_ERR_CUDA; }
  h->rp_graphs.push_back({first, count, (int)commit, exec});
  *out = exec;
  return PG_OK;
}

int pg_replay_run(pg_engine* h, int first, int count, double* dE_out, uint8_t* accept_out, float* elapsed_ms) {
  if (!h || first < 0 || count < 0 || first + count > (int)h->rp_moves.size()) return PG_ERR_INVALID;
  if (h->rp_mc) { h->err = "the proposal buffers hold an MC batch (pg_mc_upload), not a replay"; return PG_ERR_STATE; }
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  int rc = flush_commit(h);
  if (rc) return rc;
  h->ch.valid = false;
  cudaGraphExec_t exec = nullptr;
  rc = replay_graph(h, first, count, true, &exec);   // built outside the timed events
  if (rc) return rc;
  PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  if (exec) {
    PG_CUDA(h, cudaGraphLaunch(exec, h->stream));
    h->launches += (uint64_t)count + 1;
  } else {
    rc = replay_enqueue(h, first, count, true);
    if (rc) return rc;
  }
  PG_CUDA(h, cudaGetLastError());
  PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  PG_CUDA(h, cudaEventSynchronize(h->ev1));
  if (elapsed_ms) PG_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  if (dE_out && count > 0)
    PG_CUDA(h, cudaMemcpy(dE_out, h->d_rp_dE + first, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost));
  if (accept_out && count > 0)
    PG_CUDA(h, cudaMemcpy(accept_out, h->d_rp_acc + first, (size_t)count, cudaMemcpyDeviceToHost));
  return PG_OK;
}

int pg_replay_prepare(pg_engine* h, int first, int count, int with_commit) {
  if (!h || first < 0 || count < 0 || first + count > (int)h->rp_moves.size()) return PG_ERR_INVALID;
  if (h->rp_mc) { h->err = "the proposal buffers hold an MC batch (pg_mc_upload), not a replay"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  cudaGraphExec_t exec = nullptr;
  return replay_graph(h, first, count, with_commit != 0, &exec);
}

int pg_replay_time_delta(pg_engine* h, int first, int count, float* elapsed_ms) {
  if (!h || !elapsed_ms || first < 0 || count < 0 || first + count > (int)h->rp_moves.size()) return PG_ERR_INVALID;
  if (h->rp_mc) { h->err = "the proposal buffers hold an MC batch (pg_mc_upload), not a replay"; return PG_ERR_STATE; }
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  int rc = flush_commit(h);
  if (rc) return rc;
  cudaGraphExec_t exec = nullptr;
  rc = replay_graph(h, first, count, false, &exec);   // no commit rides along: the state does not advance
  if (rc) return rc;
  PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  if (exec) {
    PG_CUDA(h, cudaGraphLaunch(exec, h->stream));
    h->launches += (uint64_t)count;
  } else {
    rc = replay_enqueue(h, first, count, false);
    if (rc) return rc;
  }
  PG_CUDA(h, cudaGetLastError());
  PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  PG_CUDA(h, cudaEventSynchronize(h->ev1));
  PG_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  // a trial evaluated without commit leaves trial_dipl touched: restore the lag state
  k_commit<<<1, 32, 0, h->stream>>>(h->P, 0, PG_MODE_MOVE, 0, 0, nullptr, nullptr, nullptr, h->xy, h->zq, h->type,
                                    h->d_S, h->d_dS, 0, h->d_state);
  h->launches++;
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  return PG_OK;
}


// --------------------------------------------- batched MC, device-side proposals
#define PG_MC_WAVE 64   // most steps one k_propose launch builds

#define PG_MC_RESERVE(h, ptr, cap, need, T)                                \
  do {                                                                      \
    if ((need) > (cap) || !(ptr)) {                                         \
      size_t c_ = std::max<size_t>((need), (cap) * 2);                      \
      cudaFree(ptr); (ptr) = nullptr; (cap) = 0;                            \
      PG_CUDA(h, cudaMalloc((void**)&(ptr), sizeof(T) * c_));               \
      (cap) = c_;                                                           \
    }                                                                       \
  } while (0)

int pg_mc_upload(pg_engine* h, int n_moves, const pg_move_desc* moves, int n_rvec, const double* rvec) {
  if (!h || n_moves < 0 || n_rvec < 0 || (n_moves > 0 && !moves) || (n_rvec > 0 && !rvec)) return PG_ERR_INVALID;
  if (h->mc_inflight) { h->err = "an MC batch is in flight"; return PG_ERR_STATE; }
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  // the engine stream may still be reading the previous batch's buffers
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  replay_drop_graphs(h);
  h->rp_moves.clear();
  h->mc_mol.clear();
  h->rp_mc = true;
  size_t beads = 0;
  for (int m = 0; m < n_moves; m++) {
    const pg_move_desc& d = moves[m];
    if (d.mol < 0 || d.mol >= h->n_mol) { h->err = "mc: molecule index out of range"; return PG_ERR_INVALID; }
    const int len = h->mol_first[d.mol + 1] - h->mol_first[d.mol];
    if (len > PP_MAXLEN) { h->err = "mc: molecule longer than PP_MAXLEN"; return PG_ERR_CAPACITY; }
    if (d.kind == PG_MOVE_PIVOT) {
      if (d.i0 < 0 || d.i0 >= len || d.rv_offset < 0 || d.rv_offset + (len - 1) > n_rvec) { h->err = "mc: bad pivot descriptor"; return PG_ERR_INVALID; }
    } else if (d.kind == PG_MOVE_REPTATION) {
      if (d.i0 != 1 && d.i0 != -1) { h->err = "mc: reptation direction must be +1 or -1"; return PG_ERR_INVALID; }
    } else if (d.kind == PG_MOVE_CRANKSHAFT) {
      if (d.i0 < 0 || d.i0 >= len || d.rv_offset < d.i0 || d.rv_offset > len) { h->err = "mc: bad crankshaft descriptor"; return PG_ERR_INVALID; }
    } else if (d.kind != PG_MOVE_BEAD && d.kind != PG_MOVE_COM) {
      h->err = "mc: unknown move kind"; return PG_ERR_INVALID;
    }
    beads += (size_t)len;
  }
  h->rp_beads = std::max<size_t>(beads, 1);
  const int cap = (int)h->rp_beads;
  const size_t bytes = stage_size(cap), trial_bytes = sizeof(double) * 3 * (size_t)cap, tail_bytes = bytes - trial_bytes;
  const size_t mv_bytes = sizeof(PgPropDev) * (size_t)std::max(n_moves, 1);
  if (bytes > h->rp_bytes_cap || !h->d_rp) {
    size_t c = std::max(bytes, h->rp_bytes_cap * 2);
    cudaFree(h->d_rp); h->d_rp = nullptr; h->rp_bytes_cap = 0;
    PG_CUDA(h, cudaMalloc((void**)&h->d_rp, c));
    h->rp_bytes_cap = c;
  }
  if ((size_t)std::max(n_moves, 1) > h->mc_log_cap || !h->h_mc_dE) {
    size_t c = std::max<size_t>((size_t)std::max(n_moves, 1), h->mc_log_cap * 2);
    if (h->h_mc_dE) cudaFreeHost(h->h_mc_dE);
    if (h->h_mc_acc) cudaFreeHost(h->h_mc_acc);
    h->h_mc_dE = nullptr; h->h_mc_acc = nullptr; h->mc_log_cap = 0;
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_mc_dE, sizeof(double) * c, cudaHostAllocMapped));
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_mc_acc, c, cudaHostAllocMapped));
    PG_CUDA(h, cudaHostGetDevicePointer((void**)&h->d_mc_dE, h->h_mc_dE, 0));
    PG_CUDA(h, cudaHostGetDevicePointer((void**)&h->d_mc_acc, h->h_mc_acc, 0));
    h->mc_log_cap = c;
  }
  PG_MC_RESERVE(h, h->d_mc_moves, h->mc_moves_cap, (size_t)std::max(n_moves, 1), PgPropDev);
  PG_MC_RESERVE(h, h->d_mc_rv, h->mc_rv_cap, (size_t)std::max(n_rvec, 1), double4);
  if (!h->d_stop) {
    PG_CUDA(h, cudaMalloc((void**)&h->d_stop, sizeof(int)));
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_stop, sizeof(int), cudaHostAllocDefault));
  }
  const size_t pin_need = tail_bytes + mv_bytes + 64;
  if (pin_need > h->mc_pin_cap) {
    if (h->h_mc_pin) cudaFreeHost(h->h_mc_pin);
    h->h_mc_pin = nullptr; h->mc_pin_cap = 0;
    size_t c = pin_need * 2;
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_mc_pin, c, cudaHostAllocDefault));
    h->mc_pin_cap = c;
  }
  // host image of the block behind the trial coordinates (charges, types, charged-moved lists, moved flags):
  // a view whose `trial` sits 3*cap doubles in front of the pinned buffer, so the sub-array offsets are the device's
  StageView sv = stage_view(h->h_mc_pin - trial_bytes, cap);
  PgPropDev* pm = reinterpret_cast<PgPropDev*>(h->h_mc_pin + ((tail_bytes + 15) / 16) * 16);
  int off = 0;
  for (int m = 0; m < n_moves; m++) {
    const pg_move_desc& d = moves[m];
    ReplayMove r;
    r.mol = d.mol; r.g0 = h->mol_first[d.mol]; r.glen = h->mol_first[d.mol + 1] - r.g0; r.off = off; r.u = d.u;
    // crankshaft: only the beads strictly between the axis beads are flagged (molecule.cc:251-253); when there are
    // none the whole molecule is flagged with trial == current, which evaluates to the same dE = 0
    const bool crank = (d.kind == PG_MOVE_CRANKSHAFT) && (std::min(d.rv_offset, r.glen - 1) - d.i0 > 1);
    for (int i = 0; i < r.glen; i++) {
      sv.gq[off + i] = h->h_q[r.g0 + i]; sv.gtype[off + i] = h->h_type[r.g0 + i];
      sv.moved[off + i] = (d.kind == PG_MOVE_BEAD) ? (i == 0) : (crank ? (i > d.i0 && i < std::min(d.rv_offset, r.glen - 1)) : 1);
    }
    r.nq = stage_fill_qidx(sv, off, r.glen);
    h->rp_moves.push_back(r);
    h->mc_mol.push_back(d.mol);
    PgPropDev& q = pm[m];
    q.g0 = r.g0; q.glen = r.glen; q.kind = d.kind; q.i0 = d.i0; q.off = off; q.rv_off = d.rv_offset; q.pad0 = q.pad1 = 0;
    q.s = d.s; q.vx = d.v[0]; q.vy = d.v[1]; q.vz = d.v[2]; q.vlen = d.vlen; q.pad2 = 0.0;
    off += r.glen;
  }
  if (n_moves > 0) {
    PG_CUDA(h, cudaMemcpyAsync(h->d_rp + trial_bytes, h->h_mc_pin, tail_bytes, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(h->d_mc_moves, pm, sizeof(PgPropDev) * (size_t)n_moves, cudaMemcpyHostToDevice, h->stream));
  }
  if (n_rvec > 0)   // pageable source: the copy is staged by the driver before the call returns
    PG_CUDA(h, cudaMemcpyAsync(h->d_mc_rv, rvec, sizeof(double) * 4 * (size_t)n_rvec, cudaMemcpyHostToDevice, h->stream));
  return PG_OK;
}

int pg_mc_begin(pg_engine* h, int first, int count) {
  if (!h || !h->rp_mc || first < 0 || count < 0 || first + count > (int)h->rp_moves.size()) return PG_ERR_INVALID;
  if (h->mc_inflight) { h->err = "an MC batch is in flight"; return PG_ERR_STATE; }
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  int rc = flush_commit(h);
  if (rc) return rc;
  rc = ensure_partials(h, h->move_slots + h->n_sm + 8);
  if (rc) return rc;
  h->ch.valid = false;
  PG_CUDA(h, cudaMemsetAsync(h->d_stop, 0, sizeof(int), h->stream));
  PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  StageView base = stage_view(h->d_rp, (int)h->rp_beads);
  int wave_hi = first;
  int wave_mols[PG_MC_WAVE];
  for (int m = first; m < first + count; m++) {
    if (m >= wave_hi) {
      // a wave: consecutive steps that move pairwise different molecules
      int w = 0;
      while (m + w < first + count && w < PG_MC_WAVE) {
        const int mol = h->mc_mol[m + w];
        bool seen = false;
        for (int i = 0; i < w; i++) if (wave_mols[i] == mol) { seen = true; break; }
        if (seen) break;
        wave_mols[w++] = mol;
      }
      PgProposeArgs P;
      P.xy = h->xy; P.zq = h->zq; P.moves = h->d_mc_moves; P.rv = h->d_mc_rv; P.trial = base.trial;
      P.first = m; P.prev = (m > first) ? m - 1 : -1; P.state = h->d_state; P.stop = h->d_stop;
      k_propose<<<w, PP_THREADS, 0, h->stream>>>(P);
      h->launches++;
      wave_hi = m + w;
    }
    rc = replay_launch(h, m, (m > first) ? m - 1 : -1, true);
    if (rc) {
      // kernels of this batch are already queued and the last decided step has no commit behind it: the engine's
      // state can no longer be trusted — drain the stream and make every later batch call fail until a new upload
      cudaStreamSynchronize(h->stream);
      replay_invalidate(h, true);
      return rc;
    }
  }
  if (count > 0) {
    // the last step's decision (taken on the device) is applied by a stand-alone commit
    const ReplayMove& r = h->rp_moves[first + count - 1];
    StageView v;
    replay_view(h, first + count - 1, v);
    int nthreads = std::max(std::max(r.glen, h->nk), 1);
    k_commit<<<(nthreads + 255) / 256, 256, 0, h->stream>>>(h->P, -1, PG_MODE_MOVE, r.g0, r.glen, v.trial, v.gq,
                                                            v.gtype, h->xy, h->zq, h->type, h->d_S, h->d_dS,
                                                            h->P.use_ewald ? h->nk : 0, h->d_state);
    h->launches++;
  }
  PG_CUDA(h, cudaGetLastError());
  PG_CUDA(h, cudaMemcpyAsync(h->h_stop, h->d_stop, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  h->mc_first = first; h->mc_count = count; h->mc_inflight = true;
  return PG_OK;
}

int pg_mc_end(pg_engine* h, double* dE_out, uint8_t* accept_out, int* n_done, float* elapsed_ms) {
  if (!h) return PG_ERR_INVALID;
  if (!h->mc_inflight) { h->err = "no MC batch in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  h->mc_inflight = false;
  {
    // a lone Markov chain waits by polling (a blocking-sync wake-up costs tens of microseconds per batch);
    // with many replicas per GPU the waiting threads sleep instead
    const int live = (h->device >= 0 && h->device < 64) ? g_live_engines[h->device].load(std::memory_order_relaxed) : 1;
    if (live <= 2) {
      for (unsigned long spins = 1;; spins++) {
        cudaError_t q = cudaEventQuery(h->ev1);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { h->err = std::string("mc batch failed: ") + cudaGetErrorString(q); return PG_ERR_CUDA; }
        if ((spins & 0x3ff) == 0) sched_yield();
      }
    } else {
      PG_CUDA(h, cudaEventSynchronize(h->ev1));
    }
  }
  if (elapsed_ms) PG_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  const int stop = *h->h_stop;
  const int done = stop ? stop - h->mc_first : h->mc_count;
  if (done < 0 || done > h->mc_count) { h->err = "mc: inconsistent stop index"; return PG_ERR_STATE; }
  if (n_done) *n_done = done;
  if (dE_out && done > 0) memcpy(dE_out, h->h_mc_dE + h->mc_first, sizeof(double) * (size_t)done);
  if (accept_out && done > 0) memcpy(accept_out, h->h_mc_acc + h->mc_first, (size_t)done);
  return PG_OK;
}

int pg_mc_run(pg_engine* h, int first, int count, double* dE_out, uint8_t* accept_out, int* n_done, float* elapsed_ms) {
  int rc = pg_mc_begin(h, first, count);
  if (rc) return rc;
  return pg_mc_end(h, dE_out, accept_out, n_done, elapsed_ms);
}

int pg_mc_trial_xyz(pg_engine* h, int m, double* xyz) {
  if (!h || !h->rp_mc || !xyz || m < 0 || m >= (int)h->rp_moves.size()) return PG_ERR_INVALID;
  if (h->mc_inflight) { h->err = "an MC batch is in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  StageView v;
  replay_view(h, m, v);
  PG_CUDA(h, cudaMemcpy(xyz, v.trial, sizeof(double) * 3 * (size_t)h->rp_moves[m].glen, cudaMemcpyDeviceToHost));
  return PG_OK;
}

// -------------------------------------------------------------------- CBMC
int pg_trial_energies(pg_engine* h, const pg_trial_set* set, const double* bead1_xyz, const double* bead2_xyz,
                      const double* chain_xyz, const double* chain_q, const int32_t* chain_type, double* out_energy,
                      double* out_pair, double* out_ewald) {
  if (!h || !set || !bead1_xyz || !out_energy) return PG_ERR_INVALID;
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  if (h->inflight || h->mc_inflight || h->ch.inflight) { h->err = "a trial or batch is still in flight"; return PG_ERR_STATE; }
  const int nt = set->n_trials, cl = set->current_len, use2 = set->use_bead2 ? 1 : 0;
  if (nt <= 0 || cl < 0) return PG_ERR_INVALID;
  if (use2 && !bead2_xyz) return PG_ERR_INVALID;
  const int n_chain = cl * (use2 ? 2 : 1);
  if (n_chain > 0 && (!chain_xyz || !chain_q || !chain_type)) return PG_ERR_INVALID;
  if (set->type1 < 0 || set->type1 >= h->P.n_types || set->type2 < 0 || set->type2 >= h->P.n_types) {
    h->err = "trial bead type id out of range";
    return PG_ERR_INVALID;
  }
  for (int i = 0; i < n_chain; i++)
    if (chain_type[i] < 0 || chain_type[i] >= h->P.n_types) { h->err = "chain type id out of range"; return PG_ERR_INVALID; }
  int skip_b0 = 0, skip_b1 = 0;
  if (set->skip_mol_first >= 0) {
    if (set->skip_mol_last < set->skip_mol_first || set->skip_mol_last >= h->n_mol) {
      h->err = "skip molecule range out of bounds";
      return PG_ERR_INVALID;
    }
    skip_b0 = h->mol_first[set->skip_mol_first];
    skip_b1 = h->mol_first[set->skip_mol_last + 1];
  }
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  // small growth steps ride in the kernel parameters; larger ones are staged (chain once, trials per chunk)
  const bool inl = (nt <= TR_INL_T && n_chain <= TR_INL_C);
  if (!inl && (!h->h_trial_in || n_chain > h->ccap)) {
    const int ccap = std::max(64, n_chain * 2);
    if (h->h_trial_in) cudaFreeHost(h->h_trial_in);
    if (h->h_trial_ct) cudaFreeHost(h->h_trial_ct);
    cudaFree(h->d_trial_in); cudaFree(h->d_trial_ct);
    h->h_trial_in = nullptr; h->h_trial_ct = nullptr; h->d_trial_in = nullptr; h->d_trial_ct = nullptr;
    h->ccap = 0;
    const size_t in_d = (size_t)6 * TR_MAXT + (size_t)ccap * 4;
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_trial_in, sizeof(double) * in_d, cudaHostAllocDefault));
    PG_CUDA(h, cudaMalloc((void**)&h->d_trial_in, sizeof(double) * in_d));
    PG_CUDA(h, cudaHostAlloc((void**)&h->h_trial_ct, sizeof(int) * (size_t)ccap, cudaHostAllocDefault));
    PG_CUDA(h, cudaMalloc((void**)&h->d_trial_ct, sizeof(int) * (size_t)ccap));
    h->ccap = ccap;
  }
  PgTrialArgs A;
  memset(&A, 0, sizeof(A));
  fill_move_dev(h, A.D);
  A.xy = h->xy; A.zq = h->zq; A.type = h->type; A.n = h->n;
  A.skip_b0 = skip_b0; A.skip_b1 = skip_b1;
  A.nc = n_chain; A.current_len = cl; A.use_b2 = use2; A.t1 = set->type1; A.t2 = set->type2;
  A.q1 = set->q1; A.q2 = use2 ? set->q2 : 0.0;
  A.inl = inl ? 1 : 0;
  A.kvec = h->d_kvec; A.S = h->d_S; A.nk = (h->P.use_ewald ? h->nk : 0);
  A.counter = h->d_tr_counter; A.mail = h->d_tmail; A.Pg = h->d_P;
  const size_t ccap = (size_t)h->ccap;
  if (inl) {
    for (int i = 0; i < n_chain; i++) {
      A.inl_cxyz[3 * i] = chain_xyz[3 * i]; A.inl_cxyz[3 * i + 1] = chain_xyz[3 * i + 1]; A.inl_cxyz[3 * i + 2] = chain_xyz[3 * i + 2];
      A.inl_cq[i] = chain_q[i]; A.inl_ct[i] = chain_type[i];
    }
  } else {
    double* hcx = h->h_trial_in + 6 * TR_MAXT;
    double* hcq = hcx + 3 * ccap;
    if (n_chain > 0) {
      memcpy(hcx, chain_xyz, sizeof(double) * 3 * (size_t)n_chain);
      memcpy(hcq, chain_q, sizeof(double) * (size_t)n_chain);
      memcpy(h->h_trial_ct, chain_type, sizeof(int) * (size_t)n_chain);
      PG_CUDA(h, cudaMemcpyAsync(h->d_trial_in + 6 * TR_MAXT, hcx, sizeof(double) * 4 * ccap, cudaMemcpyHostToDevice, h->stream));
      PG_CUDA(h, cudaMemcpyAsync(h->d_trial_ct, h->h_trial_ct, sizeof(int) * (size_t)n_chain, cudaMemcpyHostToDevice, h->stream));
    }
    A.b1 = h->d_trial_in; A.b2 = h->d_trial_in + 3 * TR_MAXT;
    A.cxyz = h->d_trial_in + 6 * TR_MAXT; A.cq = A.cxyz + 3 * ccap; A.ctype = h->d_trial_ct;
  }
  const int n_sm = std::max(1, h->n_sm);
  const long long slots_all = (long long)(h->n + n_chain) * A.D.img_split;
  if (slots_all >= (1LL << 31) / TR_MAXT) { h->err = "growth step too large for 32-bit work indexing"; return PG_ERR_CAPACITY; }
  A.n_tiles = (int)std::max<long long>(1, (slots_all + TR_THREADS - 1) / TR_THREADS);
  for (int t0 = 0; t0 < nt; t0 += TR_MAXT) {
    const int ntc = std::min(TR_MAXT, nt - t0);
    A.nt = ntc;
    if (inl) {
      memcpy(A.inl_b1, bead1_xyz + 3 * (size_t)t0, sizeof(double) * 3 * (size_t)ntc);
      if (use2) memcpy(A.inl_b2, bead2_xyz + 3 * (size_t)t0, sizeof(double) * 3 * (size_t)ntc);
    } else {
      memcpy(h->h_trial_in, bead1_xyz + 3 * (size_t)t0, sizeof(double) * 3 * (size_t)ntc);
      if (use2) memcpy(h->h_trial_in + 3 * TR_MAXT, bead2_xyz + 3 * (size_t)t0, sizeof(double) * 3 * (size_t)ntc);
      PG_CUDA(h, cudaMemcpyAsync(h->d_trial_in, h->h_trial_in, sizeof(double) * 6 * TR_MAXT, cudaMemcpyHostToDevice, h->stream));
    }
    // lanes per k vector: enough for the longer of the two term lists (S' terms, trials)
    int lpk = TR_MIN_LPK;
    while (lpk < 32 && lpk < std::max(skip_b1 - skip_b0 + n_chain, ntc)) lpk <<= 1;
    int helpers = A.nk > 0 ? (int)std::min<long long>(n_sm, ((long long)A.nk * lpk + TR_THREADS - 1) / TR_THREADS) : 0;
    A.lpk = lpk; A.n_helpers = helpers;
    const long long units = (long long)A.n_tiles * ntc;
    A.n_pair_ctas = (int)std::max(1LL, std::min(units, (long long)std::max(1, h->trial_slots - helpers)));
    const size_t need = (size_t)(helpers + A.n_tiles) * (size_t)ntc;
    if (need > h->tr_partial_cap) {
      cudaFree(h->d_tr_partial); h->d_tr_partial = nullptr; h->tr_partial_cap = 0;
      PG_CUDA(h, cudaMalloc((void**)&h->d_tr_partial, sizeof(double2) * need * 2));
      h->tr_partial_cap = need * 2;
    }
    if ((size_t)A.n_tiles > h->tr_mz_cap) {
      cudaFree(h->d_tr_mz); h->d_tr_mz = nullptr; h->tr_mz_cap = 0;
      PG_CUDA(h, cudaMalloc((void**)&h->d_tr_mz, sizeof(double) * (size_t)A.n_tiles * 2));
      h->tr_mz_cap = (size_t)A.n_tiles * 2;
    }
    A.partial = h->d_tr_partial; A.mz_partial = h->d_tr_mz;
    A.seq = ++h->seq;
    k_trials<<<helpers + A.n_pair_ctas, TR_THREADS, 0, h->stream>>>(A);
    h->launches++;
    PG_CUDA(h, cudaGetLastError());
    // wait for the 3 ntc self-validating records
    volatile PgMailRec* m = h->h_tmail;
    SpinWait w;
    for (int r = 0; r < 3 * ntc; r++) {
      double v;
      for (;;) {
        if (__atomic_load_n(&h->h_tmail[r].seq, __ATOMIC_ACQUIRE) == A.seq) {
          v = m[r].value;
          __atomic_thread_fence(__ATOMIC_ACQUIRE);
          if (m[r].seq == A.seq) break;
        }
        int wrc = w.step(h);
        if (wrc) return wrc;
      }
      const int t = t0 + r / 3;
      if (r % 3 == 0) out_energy[t] = v;
      else if (r % 3 == 1) { if (out_pair) out_pair[t] = v; }
      else if (out_ewald) out_ewald[t] = v;
    }
  }
  return PG_OK;
}

// ------------------------------------------------------------ GC bookkeeping
static void result_to_totals(const PgResult* r, double sign, pg_totals* t) {
  t->pair = sign * r->pair; t->ewald = sign * r->ewald; t->bond = sign * r->bond; t->ext = sign * r->ext;
  t->real = sign * r->real; t->recip = sign * r->recip; t->self = r->self_e; t->dipole = sign * r->dipole;
}

int pg_insert_molecules(pg_engine* h, int n_new_mol, const int32_t* mol_len, const double* xyz, const double* q,
                        const int32_t* type, pg_totals* added) {
  if (!h || n_new_mol <= 0 || !mol_len || !xyz || !q || !type) return PG_ERR_INVALID;
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  if (h->inflight || h->mc_inflight || h->ch.inflight) { h->err = "a trial or batch is still in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  int n_add = 0;
  for (int m = 0; m < n_new_mol; m++) {
    if (mol_len[m] <= 0) { h->err = "molecule length must be positive"; return PG_ERR_INVALID; }
    n_add += mol_len[m];
  }
  for (int i = 0; i < n_add; i++)
    if (type[i] < 0 || type[i] >= h->P.n_types) { h->err = "bead type id out of range"; return PG_ERR_INVALID; }
  int rc = ensure_capacity(h, h->n + n_add);
  if (!rc) rc = ensure_group(h, n_add);
  if (rc) return rc;
  StageView sv = stage_view(h->h_stage, h->gcap);
  memcpy(sv.trial, xyz, sizeof(double) * 3 * (size_t)n_add);
  for (int i = 0; i < n_add; i++) { sv.gq[i] = q[i]; sv.gtype[i] = type[i]; sv.moved[i] = 1; }
  PG_CUDA(h, cudaMemcpyAsync(h->d_stage, h->h_stage, h->stage_bytes, cudaMemcpyHostToDevice, h->stream));
  // potential_bond.cc:29-34: only the LAST appended molecule's bond energy is added
  const int last_len = mol_len[n_new_mol - 1];
  rc = launch_delta(h, PG_MODE_INSERT, h->n, n_add, mol_len[0], n_add - last_len, last_len, h->d_stage, h->gcap, false,
                    0.0, 0, true);
  if (rc) return rc;
  rc = wait_result(h, h->seq);
  if (rc) return rc;
  if (added) result_to_totals(h->h_result, 1.0, added);
  rc = launch_commit(h, 1, PG_MODE_INSERT, h->n, n_add, h->d_stage, h->gcap);
  if (rc) return rc;
  std::vector<int> hmol(n_add);
  int b = 0;
  for (int m = 0; m < n_new_mol; m++) {
    for (int i = 0; i < mol_len[m]; i++) hmol[b++] = h->n_mol + m;
    h->mol_first.push_back(h->mol_first.back() + mol_len[m]);
  }
  PG_CUDA(h, cudaMemcpyAsync(h->mol + h->n, hmol.data(), sizeof(int) * (size_t)n_add, cudaMemcpyHostToDevice, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->h_q.insert(h->h_q.end(), q, q + n_add);
  h->h_type.insert(h->h_type.end(), type, type + n_add);
  h->n += n_add;
  h->n_mol += n_new_mol;
  h->ch.valid = false;
  h->qidx_valid = false;
  replay_invalidate(h, true);
  return PG_OK;
}

int pg_delete_molecules(pg_engine* h, int mf, int ml, pg_totals* removed) {
  if (!h || mf < 0 || ml < mf || ml >= h->n_mol) return PG_ERR_INVALID;
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  if (h->inflight || h->mc_inflight || h->ch.inflight) { h->err = "a trial or batch is still in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  const int b0 = h->mol_first[mf], b1 = h->mol_first[ml + 1], glen = b1 - b0;
  int rc = ensure_group(h, glen);
  if (rc) return rc;
  StageView sv = stage_view(h->h_stage, h->gcap);
  for (int i = 0; i < glen; i++) {
    sv.trial[3 * i] = sv.trial[3 * i + 1] = sv.trial[3 * i + 2] = 0.0;
    sv.gq[i] = h->h_q[b0 + i]; sv.gtype[i] = h->h_type[b0 + i]; sv.moved[i] = 1;
  }
  PG_CUDA(h, cudaMemcpyAsync(h->d_stage, h->h_stage, h->stage_bytes, cudaMemcpyHostToDevice, h->stream));
  const int first_len = h->mol_first[mf + 1] - b0;
  rc = launch_delta(h, PG_MODE_DELETE, b0, glen, first_len, 0, first_len, h->d_stage, h->gcap, false, 0.0, 0, true);
  if (rc) return rc;
  rc = wait_result(h, h->seq);
  if (rc) return rc;
  if (removed) result_to_totals(h->h_result, -1.0, removed);
  rc = launch_commit(h, 1, PG_MODE_DELETE, b0, glen, h->d_stage, h->gcap);
  if (rc) return rc;
  const int tail = h->n - b1, nm = ml - mf + 1;
  const int nb = std::max(1, (tail + 255) / 256);
  if (tail > 0) {
    k_gather_tail<<<nb, 256, 0, h->stream>>>(h->xy, h->zq, h->type, h->mol, b1, h->n, h->t_xy, h->t_zq, h->t_type,
                                             h->t_mol);
    h->launches++;
  }
  k_scatter_tail<<<nb, 256, 0, h->stream>>>(h->xy, h->zq, h->type, h->mol, b0, tail, nm, h->t_xy, h->t_zq, h->t_type,
                                            h->t_mol, h->d_state, h->n - glen);
  h->launches++;
  PG_CUDA(h, cudaGetLastError());
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->h_q.erase(h->h_q.begin() + b0, h->h_q.begin() + b1);
  h->h_type.erase(h->h_type.begin() + b0, h->h_type.begin() + b1);
  std::vector<int> mfirst;
  for (int m = 0; m <= mf; m++) mfirst.push_back(h->mol_first[m]);
  for (int m = ml + 2; m <= h->n_mol; m++) mfirst.push_back(h->mol_first[m] - glen);
  h->mol_first.swap(mfirst);
  h->n -= glen;
  h->n_mol -= nm;
  h->ch.valid = false;
  h->qidx_valid = false;
  replay_invalidate(h, true);
  return PG_OK;
}

// --------------------------------------------------- wall-force pressure
int pg_wall_force(pg_engine* h, int n_phantom, double* out6) {
  if (!h || !out6 || n_phantom < 0 || n_phantom > h->n_mol) return PG_ERR_INVALID;
  for (int m = 0; m < n_phantom; m++)
    if (h->mol_first[m + 1] - h->mol_first[m] != 1) { h->err = "wall sites must be single-bead molecules"; return PG_ERR_INVALID; }
  for (int k = 0; k < 6; k++) out6[k] = 0.0;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  const int n = h->n;
  if (n == 0) return PG_OK;
  std::vector<unsigned char> klass((size_t)n);
  for (int m = 0; m < h->n_mol; m++) {
    const int len = h->mol_first[m + 1] - h->mol_first[m];
    const unsigned char c = (m < n_phantom) ? 2 : (len > 1 ? 1 : 0);
    for (int b = h->mol_first[m]; b < h->mol_first[m + 1]; b++) klass[b] = c;
  }
  PgWallForceArgs A;
  A.xy = h->xy; A.zq = h->zq; A.type = h->type; A.n = n; A.phantom = n_phantom;
  A.b_first = h->mol_first[n_phantom / 2];
  A.kvec = h->d_kvec; A.nk = h->P.use_ewald ? h->nk : 0;
  const long long work = (long long)n_phantom * (n - A.b_first) + (n - n_phantom);
  if (work <= 0) return PG_OK;
  const long long blocks = (work + WF_THREADS - 1) / WF_THREADS;
  if (blocks > (1LL << 30)) { h->err = "wall-force sample too large"; return PG_ERR_CAPACITY; }
  unsigned char* d_klass = nullptr;
  double* d_part = nullptr;
  std::vector<double> part((size_t)blocks * 6);
  cudaError_t e = cudaMalloc((void**)&d_klass, (size_t)n);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_part, sizeof(double) * part.size());
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_klass, klass.data(), (size_t)n, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    A.klass = d_klass; A.partial = d_part;
    k_wall_force<<<(unsigned)blocks, WF_THREADS, 0, h->stream>>>(h->P, A);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_klass); cudaFree(d_part);
  if (e != cudaSuccess) { h->err = std::string("pg_wall_force: ") + cudaGetErrorString(e); return PG_ERR_CUDA; }
  for (long long b = 0; b < blocks; b++)
    for (int k = 0; k < 6; k++) out6[k] += part[(size_t)b * 6 + k];
  return PG_OK;
}

// ------------------------------------------- volume-perturbation pressure
int pg_vol_scaling_sample(pg_engine* h, int n_phantom, double dz, pg_vol_sample* out) {
  if (!h || !out || n_phantom < 0 || n_phantom > h->n_mol || !(dz > 0)) return PG_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  const int n = h->n, n_mol = h->n_mol;
  out->n_free = n_mol - n_phantom;
  if (n == 0) return PG_OK;
  // parameter block of the box stretched along z (pressure.cc:195, potential_ewald_coul.cc:51-54, :98-100)
  PgDev Pz = h->P;
  Pz.box[2] = h->P.box[2] + dz;
  Pz.inv_box[2] = 1.0 / Pz.box[2];
  Pz.half_box[2] = 0.5 * Pz.box[2];
  std::vector<int> kl;
  std::vector<double> kw;
  if (h->P.use_ewald) {
    Pz.ebox[2] = h->P.ebox[2] + dz;
    Pz.inv_ebox[2] = 1.0 / Pz.ebox[2];
    Pz.recip_pref = 2 * kPi * h->P.lB / (h->P.ebox[0] * h->P.ebox[1] * (h->P.ebox[2] + dz));
    bool single = (h->P.pbc[2] != 0);
    for (int i = 0; i < 3; i++)
      if (!(Pz.rc_relaxed < 0.5 * Pz.ebox[i] * (1.0 - 1e-9))) single = false;
    Pz.single_image = single ? 1 : 0;
    Pz.real_self_unit = 0.5 * pg_pair_real_d(Pz, 0.0, 0.0, 0.0, 1.0);
    // k vectors inside the cutoff sphere of either geometry (half space), with ek2 of each
    const int* rc = h->info.repl_cell;
    const double cut = h->info.repl_cutoff, alpha = h->alpha;
    for (int ax = -rc[0]; ax <= rc[0]; ax++)
      for (int ay = -rc[1]; ay <= rc[1]; ay++)
        for (int az = -rc[2]; az <= rc[2]; az++) {
          const bool pos = (ax > 0) || (ax == 0 && ay > 0) || (ax == 0 && ay == 0 && az > 0);
          if (!pos) continue;
          const double kx = ax * 2 * kPi / h->P.ebox[0], ky = ay * 2 * kPi / h->P.ebox[1];
          const double kz_o = az * 2 * kPi / h->P.ebox[2], kz_n = az * 2 * kPi / (h->P.ebox[2] + dz);
          const double k2_o = kx * kx + ky * ky + kz_o * kz_o, k2_n = kx * kx + ky * ky + kz_n * kz_n;
          const bool in_o = (k2_o > 0 && k2_o <= cut), in_n = (k2_n > 0 && k2_n <= cut);
          if (!in_o && !in_n) continue;
          kl.push_back(ax); kl.push_back(ay); kl.push_back(az); kl.push_back(0);
          kw.push_back(in_o ? exp(-k2_o / (4 * alpha)) / k2_o : 0.0);
          kw.push_back(in_n ? exp(-k2_n / (4 * alpha)) / k2_n : 0.0);
        }
  }
  for (int i = 0; i < 3; i++) Pz.same_box[i] = (Pz.ebox[i] == Pz.box[i]) ? 1 : 0;
  const int nku = (int)(kw.size() / 2);
  PgVolArgs A;
  memset(&A, 0, sizeof(A));
  A.xy = h->xy; A.zq = h->zq; A.type = h->type; A.mol = h->mol;
  A.n = n; A.n_mol = n_mol; A.phantom = n_phantom; A.dz = dz; A.nku = nku;
  const int n_rt = (n + PG_TILE - 1) / PG_TILE;
  int pc = PG_TILE;
  while (pc > 8 && (long long)n_rt * ((n + pc - 1) / pc) < 4LL * std::max(h->n_sm, 1)) pc >>= 1;
  const int n_ct = (n + pc - 1) / pc;
  A.pc = pc; A.n_pair_part = n_rt * n_ct;
  // one scratch block: Pz | mol_first | z_new | mol_part | pair_part | k_part | kw | kl | out | grp
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_P = take(sizeof(PgDev)), o_mf = take(sizeof(int) * (size_t)(n_mol + 1)), o_zn = take(sizeof(double) * (size_t)n),
               o_mp = take(sizeof(double) * 4 * (size_t)n_mol), o_rp = take(sizeof(double) * 32 * (size_t)A.n_pair_part),
               o_kp = take(sizeof(double) * 16 * (size_t)(nku > 0 ? nku : 1)), o_kw = take(sizeof(double) * 2 * (size_t)(nku > 0 ? nku : 1)),
               o_kl = take(sizeof(int) * 4 * (size_t)(nku > 0 ? nku : 1)), o_out = take(sizeof(double) * 36), o_g = take((size_t)n);
  char* d = nullptr;
  double res[36];
  cudaError_t e = cudaMalloc((void**)&d, off);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_P, &Pz, sizeof(PgDev), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + o_mf, h->mol_first.data(), sizeof(int) * (size_t)(n_mol + 1), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && nku > 0) e = cudaMemcpyAsync(d + o_kw, kw.data(), sizeof(double) * kw.size(), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess && nku > 0) e = cudaMemcpyAsync(d + o_kl, kl.data(), sizeof(int) * kl.size(), cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    A.Pz = reinterpret_cast<const PgDev*>(d + o_P); A.mol_first = reinterpret_cast<const int*>(d + o_mf);
    A.z_new = reinterpret_cast<double*>(d + o_zn); A.mol_part = reinterpret_cast<double*>(d + o_mp);
    A.pair_part = reinterpret_cast<double*>(d + o_rp); A.k_part = reinterpret_cast<double*>(d + o_kp);
    A.kw = reinterpret_cast<const double*>(d + o_kw); A.kl = reinterpret_cast<const int*>(d + o_kl);
    A.out = reinterpret_cast<double*>(d + o_out); A.grp = reinterpret_cast<unsigned char*>(d + o_g);
    k_vol_prep<<<(n_mol + 127) / 128, 128, 0, h->stream>>>(h->P, A);
    k_vol_pairs<<<dim3((unsigned)n_rt, (unsigned)n_ct), PG_TILE, 0, h->stream>>>(h->P, A);
    h->launches += 2;
    if (nku > 0) { k_vol_recip<<<(nku + 63) / 64, 64, 0, h->stream>>>(h->P, A); h->launches++; }
    k_vol_final<<<1, VS_FINAL_THREADS, 0, h->stream>>>(h->P, A);
    h->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(res, d + o_out, sizeof(res), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) { h->err = std::string("pg_vol_scaling_sample: ") + cudaGetErrorString(e); return PG_ERR_CUDA; }
  for (int c = 0; c < 16; c++) { out->el[c] = res[c]; out->hs[c] = res[16 + c]; }
  out->bond = res[32]; out->dipole = res[33]; out->dU = res[34];
  return PG_OK;
}

// ------------------------------------------------------- reciprocal space
int pg_sk_compute_slice(pg_engine* h, int k_first, int k_count, double* sk_dev) {
  if (!h || k_first < 0 || k_count < 0 || k_first + k_count > h->nk) return PG_ERR_INVALID;
  if (k_count == 0) return PG_OK;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  double2* out = sk_dev ? reinterpret_cast<double2*>(sk_dev) : (h->d_S + k_first);
  { int src_ = launch_sk_slice(h, k_first, k_count, out); if (src_) return src_; }
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  return PG_OK;
}

int pg_sk_set(pg_engine* h, const double* sk_dev) {
  if (!h || !sk_dev) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  if (h->nk > 0)
    PG_CUDA(h, cudaMemcpyAsync(h->d_S, sk_dev, sizeof(double2) * (size_t)h->nk, cudaMemcpyDeviceToDevice, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  return PG_OK;
}

int pg_sk_energy(pg_engine* h, int k_first, int k_count, double* e_out) {
  if (!h || !e_out || k_first < 0 || k_count < 0 || k_first + k_count > h->nk) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  k_sk_energy<<<1, 256, 0, h->stream>>>(h->P, h->d_S, h->d_ek2, k_first, k_count, h->d_out8);
  h->launches++;
  PG_CUDA(h, cudaGetLastError());
  PG_CUDA(h, cudaMemcpyAsync(h->h_out8, h->d_out8, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  *e_out = h->h_out8[0];
  return PG_OK;
}

int pg_sk_download(pg_engine* h, double* sk_host) {
  if (!h || (!sk_host && h->nk > 0)) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  { int frc_ = flush_commit(h); if (frc_) return frc_; }
  if (h->nk > 0)
    PG_CUDA(h, cudaMemcpyAsync(sk_host, h->d_S, sizeof(double2) * (size_t)h->nk, cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  return PG_OK;
}

// ---- full S(k) recompute: drift reset, k-sharded over GPUs through peer-mapped memory (pg_sk.cu)
int pg_sk_export(pg_engine* h, void* handle64) {
  if (!h || !handle64) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t hd;
  PG_CUDA(h, cudaIpcGetMemHandle(&hd, h->d_xchg));
  memcpy(handle64, &hd, 64);
  return PG_OK;
}

static int sk_set_peers(pg_engine* h, int rank, int world, char* const* blocks) {
  const size_t s_bytes = sizeof(double2) * (size_t)std::max(h->nk, 1);
  {
    // nothing may allocate (an implicit device synchronisation) while a peer's kernel spins on a flag: size the
    // scratch of a recompute now, and make the driver load the kernels of the recompute (with lazy module loading the
    // FIRST launch of a kernel waits for the device to drain — behind a spinning kernel of the same process: forever)
    cudaFuncAttributes fa;
    PG_CUDA(h, cudaFuncGetAttributes(&fa, k_sk_ready));
    PG_CUDA(h, cudaFuncGetAttributes(&fa, k_sk_block));
    PG_CUDA(h, cudaFuncGetAttributes(&fa, k_sk_finish));
    PG_CUDA(h, cudaFuncGetAttributes(&fa, k_sk_energy));
    int rc = ensure_qidx(h);
    if (rc) return rc;
    const int chunk = SK_BEADS * std::max(1, (h->nq_tot + SK_BEADS * 256 - 1) / (SK_BEADS * 256));
    const size_t need = (size_t)std::max(1, (h->nq_tot + chunk - 1) / chunk) * (size_t)std::max(h->nk, 1);
    if (need > h->sk_partial_cap) {
      cudaFree(h->d_sk_partial); h->d_sk_partial = nullptr; h->sk_partial_cap = 0;
      PG_CUDA(h, cudaMalloc((void**)&h->d_sk_partial, sizeof(double2) * need));
      h->sk_partial_cap = need;
    }
  }
  memset(&h->peers, 0, sizeof(h->peers));
  h->peers.world = world; h->peers.rank = rank; h->peers.seq = 0;
  for (int r = 0; r < world; r++) {
    h->peers.S[r] = reinterpret_cast<double2*>(blocks[r]);
    h->peers.flags[r] = reinterpret_cast<unsigned int*>(blocks[r] + s_bytes);
  }
  return PG_OK;
}

int pg_sk_detach(pg_engine* h) {
  if (!h) return PG_ERR_INVALID;
  if (h->sk_inflight) { h->err = "a recompute is in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int r = 0; r < SK_MAX_PEERS; r++)
    if (h->ipc_opened[r]) { cudaIpcCloseMemHandle(h->ipc_opened[r]); h->ipc_opened[r] = nullptr; }
  memset(&h->peers, 0, sizeof(h->peers));
  PG_CUDA(h, cudaMemset(h->d_flags, 0, sizeof(unsigned int) * 2 * SK_MAX_PEERS));
  return PG_OK;
}

int pg_sk_attach(pg_engine* h, int rank, int world, const void* handles) {
  if (!h || !handles || world < 1 || world > SK_MAX_PEERS || rank < 0 || rank >= world) return PG_ERR_INVALID;
  int rc = pg_sk_detach(h);
  if (rc) return rc;
  char* blocks[SK_MAX_PEERS];
  for (int r = 0; r < world; r++) {
    if (r == rank) { blocks[r] = h->d_xchg; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
    void* p = nullptr;
    PG_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened[r] = p;
    blocks[r] = static_cast<char*>(p);
  }
  return sk_set_peers(h, rank, world, blocks);
}

int pg_sk_attach_local(pg_engine* h, int rank, int world, pg_engine* const* peers) {
  if (!h || !peers || world < 1 || world > SK_MAX_PEERS || rank < 0 || rank >= world) return PG_ERR_INVALID;
  int rc = pg_sk_detach(h);
  if (rc) return rc;
  char* blocks[SK_MAX_PEERS];
  for (int r = 0; r < world; r++) {
    if (!peers[r] || peers[r]->nk != h->nk) { h->err = "pg_sk_attach_local: peers must share the k list"; return PG_ERR_INVALID; }
    if (peers[r]->device != h->device) {
      int can = 0;
      PG_CUDA(h, cudaDeviceCanAccessPeer(&can, h->device, peers[r]->device));
      if (!can) { h->err = "pg_sk_attach_local: no peer access between the devices"; return PG_ERR_INVALID; }
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->err = cudaGetErrorString(e); return PG_ERR_CUDA; }
      cudaGetLastError();
    }
    blocks[r] = peers[r]->d_xchg;
  }
  return sk_set_peers(h, rank, world, blocks);
}

int pg_recompute_sk_begin(pg_engine* h) {
  if (!h) return PG_ERR_INVALID;
  if (h->sk_inflight) { h->err = "a recompute is in flight"; return PG_ERR_STATE; }
  if (h->pending || h->inflight || h->mc_inflight || h->ch.inflight) { h->err = "a trial or batch is still in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  int rc = flush_commit(h);
  if (rc) return rc;
  PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  if (h->P.use_ewald && h->nk > 0) {
    if (h->peers.world > 1) {
      h->peers.seq++;
      k_sk_ready<<<1, 32, 0, h->stream>>>(h->peers);
      h->launches++;
      const int q = h->nk / h->peers.world, r = h->nk % h->peers.world;
      const int first = q * h->peers.rank + std::min(h->peers.rank, r), count = q + (h->peers.rank < r ? 1 : 0);
      rc = launch_sk_slice(h, first, count, h->d_S + first, true);
    } else {
      rc = launch_sk_slice(h, 0, h->nk, h->d_S, false);
    }
    if (rc) return rc;
    k_sk_energy<<<1, 256, 0, h->stream>>>(h->P, h->d_S, h->d_ek2, 0, h->nk, h->d_out8);
    h->launches++;
    PG_CUDA(h, cudaGetLastError());
    PG_CUDA(h, cudaMemcpyAsync(h->h_out8, h->d_out8, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(h->h_out8 + 1, h->d_sk_counter, sizeof(unsigned int) * 2, cudaMemcpyDeviceToHost, h->stream));
  } else {
    h->h_out8[0] = 0.0;
    h->h_out8[1] = 0.0;
  }
  PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  h->sk_inflight = true;
  return PG_OK;
}

int pg_recompute_sk_end(pg_engine* h, double* recip_energy, float* elapsed_ms) {
  if (!h) return PG_ERR_INVALID;
  if (!h->sk_inflight) { h->err = "no recompute in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  h->sk_inflight = false;
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (recip_energy) *recip_energy = h->h_out8[0];
  if (elapsed_ms) PG_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  {
    unsigned int w[2];
    memcpy(w, h->h_out8 + 1, sizeof(w));
    if (w[1] != 0u) {
      h->err = std::string("k-sharded recompute: rank ") + std::to_string(w[1] > 100u ? w[1] - 101u : w[1] - 1u) +
               (w[1] > 100u ? " never finished its slice" : " never entered the recompute") + " (every attached rank must call it)";
      cudaMemsetAsync(h->d_sk_counter, 0, sizeof(unsigned int) * 4, h->stream);
      return PG_ERR_TIMEOUT;
    }
  }
  return PG_OK;
}

int pg_recompute_sk(pg_engine* h, double* recip_energy) {
  int rc = pg_recompute_sk_begin(h);
  if (rc) return rc;
  return pg_recompute_sk_end(h, recip_energy, nullptr);
}

// Debug only (not part of the ABI header): the flag block of the k-sharded recompute ([0..16) ready, [16..32) done) + seq.
int pgx_sk_flags(pg_engine* h, unsigned int* out33) {
  if (!h || !out33) return PG_ERR_INVALID;
  cudaSetDevice(h->device);
  cudaMemcpy(out33, h->d_flags, sizeof(unsigned int) * 2 * SK_MAX_PEERS, cudaMemcpyDeviceToHost);
  out33[32] = h->peers.seq;
  return PG_OK;
}

// Debug only (not part of the ABI header): per-CTA %globaltimer stamps of the last k_move launch.
int pgx_read_timing(pg_engine* h, unsigned long long* out, int max_ctas) {
  if (!h || !h->d_timing) return 0;
  cudaStreamSynchronize(h->stream);
  int n = std::min(h->timing_ctas, max_ctas);
  cudaMemcpy(out, h->d_timing, sizeof(unsigned long long) * 8 * (size_t)n, cudaMemcpyDeviceToHost);
  return n;
}

// --------------------------------------------------------- instrumentation
int pg_measure_fp64_peak(pg_engine* h, double* gflops) {
  if (!h || !gflops) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  cudaDeviceProp prop;
  PG_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    k_fp64_peak<<<blocks, threads, 0, h->stream>>>(h->d_out8, iters, 1.0000001, 1e-9);
    h->launches++;
    PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    PG_CUDA(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    PG_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    double fl = (double)blocks * threads * (double)iters * 8.0 * 2.0;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e9);
  }
  *gflops = best;
  return PG_OK;
}

// L2 streaming-read bandwidth (GB/s): 32 MiB buffer (a quarter of the 126 MB L2, resident after the first sweep),
// best of 5 timed launches of 16 sweeps each.
int pg_measure_l2_peak(pg_engine* h, double* gbs) {
  if (!h || !gbs) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = (size_t)32 << 20, n16 = bytes / 16;
  int4* buf = nullptr;
  PG_CUDA(h, cudaMalloc((void**)&buf, bytes));
  PG_CUDA(h, cudaMemsetAsync(buf, 1, bytes, h->stream));
  cudaDeviceProp prop;
  PG_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
  const int blocks = prop.multiProcessorCount * 8, reps = 16;
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    k_l2_read<<<blocks, 256, 0, h->stream>>>(buf, n16, reps, reinterpret_cast<int*>(h->d_out8));
    h->launches++;
    PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    PG_CUDA(h, cudaEventSynchronize(h->ev1));
    float ms = 0;
    PG_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (rep > 0) best = std::max(best, (double)bytes * reps / (ms * 1e-3) / 1e9);
  }
  cudaFree(buf);
  *gbs = best;
  return PG_OK;
}

}  // extern "C"
#include "pg_chain_host.cu"
