// plum_b200 — k_chain: a whole Markov chain of translational steps resident on the device.
//
// One thread-block CLUSTER (1..16 CTAs, distributed shared memory) owns one system replica and runs
// Simulation::Run's loop body (src/simulation/simulation.cc:216-355 of the reference) for up to max_steps steps
// without the host: it walks the caller's std::mt19937 stream itself (pg_chain_gen.h: same state layout, same draw
// order, same separately rounded arithmetic), builds the trial coordinates of the move generators
// (src/molecules/molecule.cc:103-312), evaluates ForceField::EnergyDifference (src/force_field/force_field.cc:407-434),
// takes the Metropolis decision (simulation.cc:327-332 — no variate is drawn when dE >= 1e8: the generator is on the
// device, so nothing has to stop or rewind) and applies FinalizeEnergies (force_field.cc:436-451).  One launch may carry
// many independent replicas (one cluster each), which is how a single host thread keeps a whole GPU busy.
//
// It serves the systems k_move<true> serves (three periodic axes, one box for LJ and Ewald, only the central image
// within the real-space cutoff — the 22 000-bead benchmark system S) and replaces its stream-everything loop by spatial
// structure that stays resident and is updated incrementally on accept:
//
//   LJ / WCA      a uniform cell grid (edge >= the largest LJ cutoff, <= 64 cells per axis, 4..32 bead slots per cell
//                 chosen when the grid is built + a chunked overflow list): a moved bead looks at 27 cells instead of
//                 at all N partners
//                 (pair set of potential_pair.cc:181-197, predicate r < rcut of potential_truncated_lj.cc:49-85)
//   real space    a compact list of the charged beads (FP64 record + FP32 box fractions): FP32 pre-filter against the
//                 real-space cutoff, exact FP64 separation for what passes, in-range configurations compacted into
//                 per-warp queues so that erfc runs in full warps (potential_ewald_coul.cc:134-164)
//   reciprocal    dS(k) from per-axis phase tables e^{i l theta} built by one sincos + complex recurrences per moved
//                 charged bead and axis: 2 complex products per (k, bead) instead of a sincos
//                 (potential_ewald_coul.cc:198-225 in structure-factor form; never |S_new|^2 - |S_old|^2)
//
// Work is split over the cluster's CTAs (k slice, charged-partner slice, cell units, intra-molecular pairs); every CTA
// keeps its own copy of the random stream and of the trial coordinates, so the only exchange per step is eight partial
// sums written to every CTA's shared memory (DSMEM) behind one cluster barrier, plus one barrier behind an accepted
// commit.  Sums are taken in a fixed order: a chain is reproducible run to run for a given cluster size.
//
// No tensor cores (FP64 pair arithmetic), no TMA (gathers of 16..32-byte records from L2).
#include <cooperative_groups.h>

#include "pg_chain_gen.h"

namespace cgx = cooperative_groups;

#define CH_WARPS (CH_THREADS / 32)
#define CH_MAXLEN 256       // longest molecule the kernel moves (shared-memory staging)
#define CH_CELL_CAP_MAX 32  // bead slots per cell: 4, 8, 16 or 32, chosen when the grid is built (16-byte loads)
#define CH_OVF_CAP 4096     // overflow list (beads whose cell is full)
#define CH_OVF_CHUNK 8      // overflow entries one work unit scans
#define CH_NC_MAX 64        // cells per axis at most
#define CH_CELLS_MAX (1 << 18)
#define CH_QCAP 128         // per-warp queue of (partner, configuration) pairs that passed the FP32 filter
#define CH_NEAR 256         // per-warp list of charged partners near the moved beads' bounding box
#define CH_TAB 2048         // phase-table entries (double2)
#define CH_GMAX 16          // largest cluster
#define CH_NACC 8
#define CH_KPT 8            // k vectors one thread owns at most
#define CH_NPHASE 16

enum { CH_ERR_NONE = 0, CH_ERR_RNG = 1, CH_ERR_LEN = 2, CH_ERR_KIND = 3, CH_ERR_OVERFLOW = 4, CH_ERR_K = 5 };

// One step of the chain as the host sees it (16 bytes).
struct __align__(16) PgChainRec {
  double dE;
  int mol;
  int info;   // kind (low byte, signed) | accept << 8 | stage << 16
};

struct PgChainArgs {
  PgMoveDev D;
  const PgDev* Pg;
  double2* xy; double2* zq; const int* type; int n;
  const int* mol_first; const int* chains; const int* ions;
  CgConfig cfg;
  // reciprocal space (half-space list)
  const int4* kl; const double* ek2; double2* S; double2* dS; int nk;
  int kmax[3]; double kunit[3];
  // charged beads
  int nq_tot; double2* qpos; float4* qfrac; const int* qslot;
  // cell grid
  int nc[3]; int cell_cap; int* cell_slots; int* ovf; int* ovf_n; int* bead_cell; int* bead_slot;
  // in / out
  uint32_t* mt_io;          // [625] state words + position
  PgChainRec* log;          // [max_steps]
  double* trial_log;        // optional (tests): [max_steps][trial_stride][3]
  int trial_stride;
  PgState* state;
  int* out;                 // [4] steps done, stop kind, error, overflow high-water
  int max_steps;
  int pivot_mode;           // 0 exact arms, 1 prefix-sum arms, 2 exact arms for the state + prefix-sum arms for the energies
  int specialize;           // clusters: run the four phases of the energy change on disjoint warp groups
  int pad2_;
  // instrumentation (tools/chain_probe.py): per-phase clock64 sums of thread 0 of rank 0 [5 kinds][CH_NPHASE], and a
  // mask of phases to leave out (timing experiments only: the energies are then wrong)
  unsigned long long* prof;
  int dbg_skip;
  int pad_;
  // work counters, accumulated over the launches of this engine: [0] pair configurations evaluated inside a cutoff
  // (LJ + real space + intra-molecular), [1] steps that evaluated an energy change, [2] accepted steps
  unsigned long long* counters;
};


// Cell coordinate of a position along one axis: the SAME function builds the grid on the host, files an accepted bead
// on the device and finds the cell of a trial position (separately rounded product: no contraction on either side).
PP_HD int ch_cell1(double x, double invL, int nc) {
  double s = PP_MUL(x, invL);
  s = s - floor(s);
  int i = (int)PP_MUL(s, (double)nc);
  return i >= nc ? nc - 1 : (i < 0 ? 0 : i);
}

#ifdef __CUDACC__

// Cooperative twist of the next-generation buffer when the position has crossed into it (all threads call this; every
// caller has a barrier between the thread that advanced the position and this call).  The barrier behind the read also
// orders every shared flag read just before the call (stop / err / row) against the next write of thread 0.
__device__ __forceinline__ void ch_mt_fix(PgMt& m, int tid) {
  const int need = m.need;
  __syncthreads();
  if (need) {
    const uint32_t* o = m.x[m.cur];
    uint32_t* nw = m.x[m.cur ^ 1];
    if (tid < 227) nw[tid] = cg_twist_word(o, nw, tid);
    __syncthreads();
    if (tid < 227) nw[227 + tid] = cg_twist_word(o, nw, 227 + tid);
    __syncthreads();
    if (tid < 170) nw[454 + tid] = cg_twist_word(o, nw, 454 + tid);
    if (tid == 0) m.need = 0;
    __syncthreads();
  }
}

__device__ __forceinline__ double ch_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// phase stamp (instrumentation): every warp notes when it has finished phase `ph`; at the end of the step thread 0 of
// the reporting rank charges each phase with the time from the previous phase's last warp to this phase's last warp
#define CH_STAMP(ph)                                                            \
  do {                                                                          \
    if (A.prof && lane == 0) sm.wst[ph][warp] = clock64();                      \
  } while (0)

// Exact FP64 separation and the reference's real-space term for this warp's queued (partner, configuration) pairs, one
// per lane: erfc runs in full warps.  Entries are in loop order and lane e takes entries e, e + 32, ...: deterministic.
#define CH_QUEUE_FLUSH()                                                                  \
  do {                                                                                    \
    __syncwarp();                                                                         \
    for (int e_ = lane; e_ < qn; e_ += 32) {                                              \
      const int2 h_ = sm.q_hit[warp][e_];                                                 \
      const double2 a_ = __ldcg(&A.qpos[2 * h_.x]), c_ = __ldcg(&A.qpos[2 * h_.x + 1]);   \
      const int g_ = sm.qidx[h_.y >> 1];                                                  \
      double dx_, dy_, dz_;                                                               \
      if (h_.y & 1) { dx_ = a_.x - sm.cur[0][g_]; dy_ = a_.y - sm.cur[1][g_]; dz_ = c_.x - sm.cur[2][g_]; } \
      else { dx_ = a_.x - sm.trl[0][g_]; dy_ = a_.y - sm.trl[1][g_]; dz_ = c_.x - sm.trl[2][g_]; } \
      dx_ -= Lx * mv_rint(dx_ * iLx); dy_ -= Ly * mv_rint(dy_ * iLy); dz_ -= Lz * mv_rint(dz_ * iLz); \
      const double r2_ = dx_ * dx_ + dy_ * dy_ + dz_ * dz_;                               \
      if (r2_ <= rc2) {                                                                   \
        const double r_ = sqrt(r2_);                                                      \
        if (r_ > 0 && r_ <= P.real_cutoff) { acc_real += P.lB * (sm.sq[h_.y] * c_.y) * erfc(P.sqrt_alpha * r_) / r_; acc_cnt += 1.0; } \
      }                                                                                   \
    }                                                                                     \
    __syncwarp();                                                                         \
    qn = 0;                                                                               \
  } while (0)

// FP32 pre-filter of this warp's near-partner list against every charged moved bead configuration; what passes goes
// into the (partner, configuration) queue.
#define CH_NEAR_PROCESS()                                                                 \
  do {                                                                                    \
    __syncwarp();                                                                         \
    for (int b_ = 0; b_ < nn; b_ += 32) {                                                 \
      const bool has_ = b_ + lane < nn;                                                   \
      const int j_ = has_ ? sm.q_near[warp][b_ + lane] : 0;                               \
      float4 pf_ = make_float4(0.f, 0.f, 0.f, 0.f);                                       \
      if (has_) pf_ = __ldcg(&A.qfrac[j_]);                                               \
      for (int e = 0; e < 2 * nq; e++) {                                                  \
        const float4 f = sm.fe[e];                                                        \
        float ax = pf_.x - f.x, ay = pf_.y - f.y, az = pf_.z - f.z;                       \
        ax -= mv_rintf(ax); ay -= mv_rintf(ay); az -= mv_rintf(az);                       \
        ax *= fLx; ay *= fLy; az *= fLz;                                                  \
        const bool fhit = has_ && fmaf(ax, ax, fmaf(ay, ay, az * az)) <= cutf;            \
        const unsigned mh = __ballot_sync(0xffffffffu, fhit);                             \
        if (mh) {                                                                         \
          if (fhit) sm.q_hit[warp][qn + __popc(mh & lt)] = make_int2(j_, e);              \
          qn += __popc(mh);                                                               \
          if (qn > CH_QCAP - 32) CH_QUEUE_FLUSH();                                        \
        }                                                                                 \
      }                                                                                   \
    }                                                                                     \
    __syncwarp();                                                                         \
    nn = 0;                                                                               \
  } while (0)

// k_chain is compiled twice from pg_chain_body.cuh:
//   512 threads, 1 CTA per SM   one chain per SM or a cluster of CTAs per chain (the warp-role split of the cluster mode
//                               assumes 16 warps): the shape with the shortest step
//   448 threads, 2 CTAs per SM  fleets of more chains than SMs (pg_chain_run_multi*): 72 registers per thread, 28 instead
//                               of 16 resident warps per SM hide more of the dependent-latency stalls of a step
//                               (measured, 296 chains of S: 3.06 M moves/s against 2.47 M; 2 x 384: 2.90 M, 2 x 512 at 64
//                               registers: 2.93 M — profiles/r02_chain_threads_ab.jsonl)
#define CH_THREADS 512
#define CH_MIN_CTAS 1
#define ChSmem ChSmem512
#define k_chain k_chain512
#include "pg_chain_body.cuh"
#undef CH_THREADS
#undef CH_MIN_CTAS
#undef ChSmem
#undef k_chain
#define CH_THREADS 448
#define CH_MIN_CTAS 2
#define ChSmem ChSmem448
#define k_chain k_chain448
#include "pg_chain_body.cuh"
#undef CH_THREADS
#undef CH_MIN_CTAS
#undef ChSmem
#undef k_chain
#define CH_THREADS_MAX 512

#endif  // __CUDACC__

// Host-side state of the chain path of one engine (pg_chain_host.cu).
struct PgChainHost {
  bool configured = false;    // pg_chain_configure accepted the settings
  bool valid = false;         // the resident spatial structures describe the current coordinates and molecule table
  bool inflight = false;
  int cluster = 1;            // CTAs per chain
  int phantom = 0;
  CgConfig cfg;
  int max_len = 1;
  bool want_trials = false;   // keep every step's trial coordinates (tests)
  int pivot_mode = 0;
  // device
  int *d_mol_first = nullptr, *d_chains = nullptr, *d_ions = nullptr;
  size_t mol_cap = 0;
  int nc[3] = {1, 1, 1};
  int cell_cap = 4;
  int *d_cell_slots = nullptr, *d_ovf = nullptr, *d_ovf_n = nullptr, *d_bead_cell = nullptr, *d_bead_slot = nullptr, *d_qslot = nullptr;
  size_t cell_cap_words = 0, bead_cap = 0;
  int nq_tot = 0;
  double2* d_qpos = nullptr;
  float4* d_qfrac = nullptr;
  size_t q_cap = 0;
  uint32_t* d_mt = nullptr;
  PgChainRec* d_log = nullptr;
  size_t log_cap = 0;
  double* d_trial_log = nullptr;
  size_t tl_cap = 0;
  int* d_out = nullptr;
  PgChainArgs* d_args = nullptr;
  size_t args_cap = 0;
  int max_steps = 0;
  int kmax[3] = {0, 0, 0};
  double kunit[3] = {0, 0, 0};
  unsigned long long* d_prof = nullptr;   // instrumentation
  unsigned long long* d_counters = nullptr;
  char* h_pin = nullptr;                  // pinned staging of pg_chain_run_multi_io
  char* d_pack = nullptr;                 // ... and its device image (k_chain_pack)
  size_t pin_cap = 0;
  int dbg_skip = 0;
};
