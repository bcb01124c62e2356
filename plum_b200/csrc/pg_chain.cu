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
//   LJ / WCA      a uniform cell grid (edge >= the largest LJ cutoff, <= 32^3 cells, 8 bead slots per cell + an overflow
//                 list): a moved bead looks at 27 cells instead of at all N partners
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

#define CH_THREADS 512
#define CH_WARPS (CH_THREADS / 32)
#define CH_MAXLEN 256       // longest molecule the kernel moves (shared-memory staging)
#define CH_CELL_CAP 8       // bead slots per cell (two 16-byte loads)
#define CH_OVF_CAP 4096     // overflow list (cells that are full)
#define CH_NC_MAX 32        // cells per axis at most
#define CH_QCAP 128         // per-warp queue of in-range real-space configurations
#define CH_TAB 2048         // phase-table entries (double2)
#define CH_GMAX 16          // largest cluster
#define CH_NACC 8
#define CH_KPT 8            // k vectors one thread owns at most

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
  int nc[3]; int* cell_slots; int* ovf; int* ovf_n; int* bead_cell; int* bead_slot;
  // in / out
  uint32_t* mt_io;          // [625] state words + position
  PgChainRec* log;          // [max_steps]
  double* trial_log;        // optional (tests): [max_steps][trial_stride][3]
  int trial_stride;
  PgState* state;
  int* out;                 // [4] steps done, stop kind, error, overflow high-water
  int max_steps;
  int exact_pivot;
};

struct ChSmem {
  PgChainArgs A;
  PgMt mt;
  CgStep step;
  int g0, glen, nq, row, last_pair, accept, err, stop;
  double dE;
  double cur[3][CH_MAXLEN], trl[3][CH_MAXLEN], gq[CH_MAXLEN];
  double4 rv[CH_MAXLEN];
  int gtype[CH_MAXLEN], qidx[CH_MAXLEN];
  float4 fe[2 * CH_MAXLEN];
  double sq[2 * CH_MAXLEN];
  double2 tab[CH_TAB];
  double q_r2[CH_WARPS][CH_QCAP], q_qq[CH_WARPS][CH_QCAP];
  double red[CH_WARPS][CH_NACC];
  double part[2][CH_GMAX][CH_NACC];
  double tot[CH_NACC];
  int scan[CH_WARPS];
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

#define CH_QUEUE_FLUSH()                                                                  \
  do {                                                                                    \
    __syncwarp();                                                                         \
    for (int e_ = lane; e_ < qn; e_ += 32) {                                              \
      const double r_ = sqrt(sm.q_r2[warp][e_]);                                          \
      if (r_ > 0 && r_ <= P.real_cutoff) acc_real += P.lB * sm.q_qq[warp][e_] * erfc(P.sqrt_alpha * r_) / r_; \
    }                                                                                     \
    __syncwarp();                                                                         \
    qn = 0;                                                                               \
  } while (0)

__global__ void __launch_bounds__(CH_THREADS, 1) k_chain(const PgChainArgs* __restrict__ all) {
  cgx::cluster_group cluster = cgx::this_cluster();
  const int G = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gt = rank * CH_THREADS + tid, GT = G * CH_THREADS;
  const unsigned lt = (1u << lane) - 1u;

  extern __shared__ __align__(16) unsigned char ch_raw[];
  ChSmem& sm = *reinterpret_cast<ChSmem*>(ch_raw);
  {
    const int* src = reinterpret_cast<const int*>(all + blockIdx.x / G);
    int* dst = reinterpret_cast<int*>(&sm.A);
    for (int i = tid; i < (int)(sizeof(PgChainArgs) / 4); i += CH_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const PgChainArgs& A = sm.A;
  const PgMoveDev& P = A.D;
  for (int i = tid; i < CG_N; i += CH_THREADS) sm.mt.x[0][i] = A.mt_io[i];
  if (tid == 0) {
    sm.mt.cur = 0; sm.mt.p = (int)A.mt_io[CG_N]; sm.mt.need = 1;
    sm.err = 0; sm.stop = 0; sm.accept = 0;
  }
  __syncthreads();
  ch_mt_fix(sm.mt, tid);
  if (tid == 0 && sm.mt.p >= CG_N) cg_advance(sm.mt, 0);   // position 624: the first draw twists
  __syncthreads();
  ch_mt_fix(sm.mt, tid);

  // running totals (rank 0, thread 0): the reference's E_tot members
  double E_pair = 0, E_ewald = 0, E_bond = 0, E_ext = 0, E_real = 0, E_recip = 0;
  if (rank == 0 && tid == 0) {
    E_pair = A.state->E_pair; E_ewald = A.state->E_ewald; E_bond = A.state->E_bond; E_ext = A.state->E_ext;
    E_real = A.state->E_real; E_recip = A.state->E_recip;
  }
  int ovf_hw = __ldcg(A.ovf_n);   // overflow list high-water mark (changes only in rank 0 / warp 0, re-read behind a commit)

  const double Lx = P.box[0], Ly = P.box[1], Lz = P.box[2];
  const double iLx = P.inv_box[0], iLy = P.inv_box[1], iLz = P.inv_box[2];
  const float fLx = P.fbox[0], fLy = P.fbox[1], fLz = P.fbox[2];
  const double fdelta = 1e-6 * fmax(Lx, fmax(Ly, Lz));
  const float cutf = P.use_ewald ? mv_relax_f(P.rc2_relaxed, fdelta) : -1.0f;
  const int nk = P.use_ewald ? A.nk : 0;
  unsigned k_lo, k_hi, q_lo, q_hi;
  mv_share((unsigned)nk, (unsigned)G, (unsigned)rank, k_lo, k_hi);
  mv_share((unsigned)(P.use_ewald ? A.nq_tot : 0), (unsigned)G, (unsigned)rank, q_lo, q_hi);
  const int nkt = ((int)(k_hi - k_lo) + CH_THREADS - 1) / CH_THREADS;
  if (((nk + G - 1) / G + CH_THREADS - 1) / CH_THREADS > CH_KPT && tid == 0) sm.err = CH_ERR_K;   // (the host checks first)
  const int ne0 = A.kmax[0] + 1, ne1 = A.kmax[1] + 1, ne2 = A.kmax[2] + 1, ne = ne0 + ne1 + ne2;
  const int EC = CH_TAB / ne;
  if (EC < 1 && tid == 0) sm.err = CH_ERR_K;
  const int nbx = A.nc[0] >= 3 ? 3 : 1, nby = A.nc[1] >= 3 ? 3 : 1, nbz = A.nc[2] >= 3 ? 3 : 1;
  const int nnb = nbx * nby * nbz, per = nnb + 1;
  __syncthreads();

  int step_i = 0;
  int par = 0;
  for (; step_i < A.max_steps; step_i++) {
    // ------------------------------------------------------------------ (1) the step's head: which move, which molecule
    if (tid == 0) {
      CgStep d;
      const int* mf = A.mol_first;
      const int used = cg_step_header(sm.mt, A.cfg, A.chains, A.ions, [mf](int mol) { return mf[mol + 1] - mf[mol]; }, d);
      if (used > 560 && !sm.err) sm.err = CH_ERR_RNG;
      if (d.kind == CG_CRANK) sm.err = CH_ERR_KIND;
      if (d.kind == CG_STOP_GC) sm.stop = 1;
      else cg_advance(sm.mt, used);
      if (d.kind >= 0) {
        sm.g0 = mf[d.mol]; sm.glen = mf[d.mol + 1] - mf[d.mol];
        if (sm.glen > CH_MAXLEN || sm.glen < 1 || (d.kind == CG_BEAD && sm.glen != 1)) sm.err = CH_ERR_LEN;
      }
      sm.step = d;
      sm.row = 0;
    }
    __syncthreads();
    if (sm.stop || sm.err) break;
    ch_mt_fix(sm.mt, tid);
    const int kind = sm.step.kind;
    if (kind == CG_NONE) {
      if (rank == 0 && tid == 0) { PgChainRec r; r.dE = 0.0; r.mol = -1; r.info = 0xff; A.log[step_i] = r; }
      __syncthreads();
      continue;
    }
    const int g0 = sm.g0, glen = sm.glen;

    // ------------------------------------------------------------------ (2) the molecule's current coordinates
    for (int i = tid; i < glen; i += CH_THREADS) {
      const double2 a = __ldcg(&A.xy[g0 + i]), c = __ldcg(&A.zq[g0 + i]);
      sm.cur[0][i] = a.x; sm.cur[1][i] = a.y; sm.cur[2][i] = c.x;
      sm.trl[0][i] = a.x; sm.trl[1][i] = a.y; sm.trl[2][i] = c.x;
      sm.gq[i] = c.y;
      sm.gtype[i] = A.type[g0 + i];
    }
    // ------------------------------------------------------------------ (3) pivot rows: len - 1 x (randSphere, bond length)
    if (kind == CG_PIVOT) {
      const int n_rows = sm.step.n_rows;
      for (;;) {
        const int row0 = sm.row;
        __syncthreads();   // everybody has read the row counter (and err) before thread 0 moves on
        if (row0 >= n_rows) break;
        if (A.cfg.vary_bond) {
          // the bond draw behind every accepted pair shifts the pairing of everything after it: one thread walks
          if (tid == 0) {
            int next = row0;
            const int used = cg_pivot_rows_serial(sm.mt, A.cfg, row0, n_rows, 400, reinterpret_cast<double*>(sm.rv), &next);
            if (used > 600) sm.err = CH_ERR_RNG;
            sm.row = next;
            cg_advance(sm.mt, used);
          }
          __syncthreads();
        } else {
          // rigid bonds: the stream is a sequence of candidate pairs, row r is the r-th accepted one: 256 pairs per pass
          bool acc = false;
          double v[3] = {0.0, 0.0, 0.0};
          if (tid < 256) acc = cg_sphere_pair(cg_uniform_of(cg_raw(sm.mt, 2 * tid)), cg_uniform_of(cg_raw(sm.mt, 2 * tid + 1)), v);
          const unsigned bal = __ballot_sync(0xffffffffu, acc);
          if (lane == 0) sm.scan[warp] = __popc(bal);
          __syncthreads();
          int before = 0, total = 0;
          for (int w = 0; w < 8; w++) { const int c = sm.scan[w]; if (w < warp) before += c; total += c; }
          const int need = n_rows - row0, mine = before + __popc(bal & lt);
          if (acc && mine < need) {
            sm.rv[row0 + mine] = make_double4(v[0], v[1], v[2], A.cfg.bond_len);
            if (mine == need - 1) sm.last_pair = tid;
          }
          __syncthreads();
          if (tid == 0) {
            if (total >= need) { cg_advance(sm.mt, 2 * (sm.last_pair + 1)); sm.row = n_rows; }
            else { cg_advance(sm.mt, 512); sm.row = row0 + total; }
          }
          __syncthreads();
        }
        ch_mt_fix(sm.mt, tid);
        if (sm.err) break;
      }
    }
    __syncthreads();
    if (sm.err) break;

    // ------------------------------------------------------------------ (4) trial coordinates (pg_propose_math.h)
    {
      const CgStep& d = sm.step;
      if (kind == CG_BEAD) {
        if (tid < 3) sm.trl[tid][0] = pp_bead_translate(sm.cur[tid][0], d.s, d.v[tid]);
      } else if (kind == CG_COM) {
        for (int i = tid; i < 3 * glen; i += CH_THREADS) {
          const int a = i / glen, g = i - a * glen;
          sm.trl[a][g] = pp_com_translate(sm.cur[a][g], d.v[a]);
        }
      } else if (kind == CG_REPT) {
        const int dir = d.i0, end = (dir > 0) ? glen - 1 : 0;
        for (int i = tid; i < 3 * glen; i += CH_THREADS) {
          const int a = i / glen, g = i - a * glen;
          sm.trl[a][g] = (g == end) ? pp_reptation_end(sm.cur[a][g], d.s, d.v[a], d.vlen) : sm.cur[a][g + dir];
        }
      } else {   // CG_PIVOT: the two arms in two warps; an arm is sequential by construction (molecule.cc:170-231)
        const int p = d.i0;
        const double msr = d.s;
        double* sx = sm.trl[0]; double* sy = sm.trl[1]; double* sz = sm.trl[2];
        if (warp == 0 && p + 1 < glen) {
          double a[3] = {sx[p], sy[p], sz[p]};
          double b[3] = {sx[p + 1], sy[p + 1], sz[p + 1]};
          for (int i = p + 1; i < glen; i++) {
            const double4 r = sm.rv[i - (p + 1)];
            double nb[3] = {0.0, 0.0, 0.0};
            if (i + 1 < glen) { nb[0] = sx[i + 1]; nb[1] = sy[i + 1]; nb[2] = sz[i + 1]; }
            __syncwarp();
            const double v[3] = {r.x, r.y, r.z};
            double m[3];
            pp_pivot_step(a, b, msr, v, r.w, m);
            for (int j = i + lane; j < glen; j += 32) {
              sx[j] = PP_ADD(sx[j], m[0]); sy[j] = PP_ADD(sy[j], m[1]); sz[j] = PP_ADD(sz[j], m[2]);
            }
            __syncwarp();
            a[0] = PP_ADD(b[0], m[0]); a[1] = PP_ADD(b[1], m[1]); a[2] = PP_ADD(b[2], m[2]);
            b[0] = PP_ADD(nb[0], m[0]); b[1] = PP_ADD(nb[1], m[1]); b[2] = PP_ADD(nb[2], m[2]);
          }
        } else if (warp == 1 && p > 0) {
          const int row0 = glen - 1 - p;
          double a[3] = {sx[p], sy[p], sz[p]};
          double b[3] = {sx[p - 1], sy[p - 1], sz[p - 1]};
          for (int i = p - 1; i >= 0; i--) {
            const double4 r = sm.rv[row0 + (p - 1 - i)];
            double nb[3] = {0.0, 0.0, 0.0};
            if (i > 0) { nb[0] = sx[i - 1]; nb[1] = sy[i - 1]; nb[2] = sz[i - 1]; }
            __syncwarp();
            const double v[3] = {r.x, r.y, r.z};
            double m[3];
            pp_pivot_step(a, b, msr, v, r.w, m);
            for (int j = i - lane; j >= 0; j -= 32) {
              sx[j] = PP_ADD(sx[j], m[0]); sy[j] = PP_ADD(sy[j], m[1]); sz[j] = PP_ADD(sz[j], m[2]);
            }
            __syncwarp();
            a[0] = PP_ADD(b[0], m[0]); a[1] = PP_ADD(b[1], m[1]); a[2] = PP_ADD(b[2], m[2]);
            b[0] = PP_ADD(nb[0], m[0]); b[1] = PP_ADD(nb[1], m[1]); b[2] = PP_ADD(nb[2], m[2]);
          }
        }
      }
    }
    // charged moved beads (group-relative), in bead order
    if (warp == 2) {
      int cnt = 0;
      for (int base = 0; base < glen; base += 32) {
        const int i = base + lane;
        const bool c = P.use_ewald && i < glen && (kind != CG_BEAD || i == 0) && sm.gq[i] != 0.0;
        const unsigned b = __ballot_sync(0xffffffffu, c);
        if (c) sm.qidx[cnt + __popc(b & lt)] = i;
        cnt += __popc(b);
      }
      if (lane == 0) sm.nq = cnt;
    }
    __syncthreads();
    if (A.trial_log && rank == 0)
      for (int i = tid; i < 3 * glen; i += CH_THREADS) {
        const int g = i / 3, a = i - 3 * g;
        A.trial_log[((size_t)step_i * A.trial_stride + g) * 3 + a] = sm.trl[a][g];
      }
    const int nq = sm.nq;
    for (int e = tid; e < 2 * nq; e += CH_THREADS) {
      const int g = sm.qidx[e >> 1];
      const double(*c)[CH_MAXLEN] = (e & 1) ? sm.cur : sm.trl;
      sm.fe[e] = make_float4(mv_frac(c[0][g], iLx), mv_frac(c[1][g], iLy), mv_frac(c[2][g], iLz), 0.0f);
      sm.sq[e] = (e & 1) ? -sm.gq[g] : sm.gq[g];
    }

    double acc_pair = 0.0, acc_real = 0.0, acc_rec = 0.0, acc_ov = 0.0, w_sum = 0.0, b_sum = 0.0, w_out = 0.0;

    // ------------------------------------------------------------------ (5) reciprocal space: this CTA's k slice
    if (nq > 0 && nk > 0) {
      double dre[CH_KPT], dim[CH_KPT];
#pragma unroll
      for (int i = 0; i < CH_KPT; i++) { dre[i] = 0.0; dim[i] = 0.0; }
      for (int c0 = 0; c0 < 2 * nq; c0 += EC) {
        const int ec = min(EC, 2 * nq - c0);
        __syncthreads();   // the previous chunk's tables are still being read (and sm.fe / sm.sq are now written)
        for (int t = tid; t < 3 * ec; t += CH_THREADS) {
          const int el = t / 3, ax = t - 3 * el, e = c0 + el;
          const int g = sm.qidx[e >> 1];
          double x = (e & 1) ? sm.cur[ax][g] : sm.trl[ax][g];
          x = pg_wrap_pos(x, P.ebox[ax], P.inv_ebox[ax], 1);
          double s1, c1;
          sincos(A.kunit[ax] * x, &s1, &c1);
          double2* row = sm.tab + el * ne + (ax == 0 ? 0 : (ax == 1 ? ne0 : ne0 + ne1));
          const int km = A.kmax[ax];
          row[0] = make_double2(1.0, 0.0);
          double cr = c1, sr = s1;
          if (km >= 1) row[1] = make_double2(c1, s1);
          for (int l = 2; l <= km; l++) {
            const double cn = cr * c1 - sr * s1, sn = sr * c1 + cr * s1;
            cr = cn; sr = sn;
            row[l] = make_double2(cr, sr);
          }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < CH_KPT; i++) {
          const int k = (int)k_lo + tid + i * CH_THREADS;
          if (i < nkt && k < (int)k_hi) {
            const int4 l = __ldg(&A.kl[k]);
            const int aly = abs(l.y), alz = abs(l.z);
            double ar = 0.0, ai = 0.0;
            for (int el = 0; el < ec; el++) {
              const double2* row = sm.tab + el * ne;
              const double2 a = row[l.x];
              double2 b = row[ne0 + aly], c = row[ne0 + ne1 + alz];
              if (l.y < 0) b.y = -b.y;
              if (l.z < 0) c.y = -c.y;
              const double abr = a.x * b.x - a.y * b.y, abi = a.x * b.y + a.y * b.x;
              const double w = sm.sq[c0 + el];
              ar += w * (abr * c.x - abi * c.y);
              ai += w * (abr * c.y + abi * c.x);
            }
            dre[i] += ar; dim[i] += ai;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < CH_KPT; i++) {
        const int k = (int)k_lo + tid + i * CH_THREADS;
        if (i < nkt && k < (int)k_hi) {
          const double2 S = __ldcg(&A.S[k]);
          const double ek2 = __ldg(&A.ek2[k]);
          // never |S_new|^2 - |S_old|^2: 2 Re(conj(S) dS) + |dS|^2; x2 for the -k half
          acc_rec += 2.0 * ek2 * (2.0 * (S.x * dre[i] + S.y * dim[i]) + (dre[i] * dre[i] + dim[i] * dim[i]));
          __stcg(&A.dS[k], make_double2(dre[i], dim[i]));
        }
      }
    } else {
      __syncthreads();   // sm.fe / sm.sq written above
    }

    // ------------------------------------------------------------------ (6) real space: charged partners of this CTA's slice
    if (nq > 0) {
      int qn = 0;
      const double rc2 = P.rc2_relaxed;
      for (unsigned jb = q_lo; jb < q_hi; jb += CH_THREADS) {
        const unsigned j = jb + tid;
        bool valid = j < q_hi;
        float4 pf = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          pf = __ldcg(&A.qfrac[j]);
          const int bead = __float_as_int(pf.w);
          valid = !(bead >= g0 && bead < g0 + glen);   // beads of the moved molecule itself: (8)
        }
        if (!__any_sync(0xffffffffu, valid)) continue;
        bool have = false;
        double px = 0, py = 0, pz = 0, pq = 0;
        for (int e = 0; e < 2 * nq; e++) {
          const float4 f = sm.fe[e];
          float ax = pf.x - f.x, ay = pf.y - f.y, az = pf.z - f.z;
          ax -= mv_rintf(ax); ay -= mv_rintf(ay); az -= mv_rintf(az);
          ax *= fLx; ay *= fLy; az *= fLz;
          const float f2 = fmaf(ax, ax, fmaf(ay, ay, az * az));
          const bool fhit = valid && f2 <= cutf;
          if (!__any_sync(0xffffffffu, fhit)) continue;
          bool hit = false;
          double r2 = 0.0;
          if (fhit) {
            if (!have) {
              const double2 a = __ldcg(&A.qpos[2 * j]), c = __ldcg(&A.qpos[2 * j + 1]);
              px = a.x; py = a.y; pz = c.x; pq = c.y;
              have = true;
            }
            const int g = sm.qidx[e >> 1];
            double dx, dy, dz;
            if (e & 1) { dx = px - sm.cur[0][g]; dy = py - sm.cur[1][g]; dz = pz - sm.cur[2][g]; }
            else { dx = px - sm.trl[0][g]; dy = py - sm.trl[1][g]; dz = pz - sm.trl[2][g]; }
            dx -= Lx * mv_rint(dx * iLx); dy -= Ly * mv_rint(dy * iLy); dz -= Lz * mv_rint(dz * iLz);
            r2 = dx * dx + dy * dy + dz * dz;
            hit = r2 <= rc2;
          }
          const unsigned mh = __ballot_sync(0xffffffffu, hit);
          if (mh) {
            if (hit) {
              const int pos = qn + __popc(mh & lt);
              sm.q_r2[warp][pos] = r2; sm.q_qq[warp][pos] = sm.sq[e] * pq;
            }
            qn += __popc(mh);
            if (qn > CH_QCAP - 32) CH_QUEUE_FLUSH();
          }
        }
      }
      if (qn > 0) CH_QUEUE_FLUSH();
    }

    // ------------------------------------------------------------------ (7) LJ / WCA through the cell grid
    if (P.pair_kind == 1) {
      const int n_mv = (kind == CG_BEAD) ? 1 : glen;
      const int U = n_mv * 2 * per;
      const double ljc2max = P.ljc2max;
      const PgDev* __restrict__ Pg = A.Pg;
      for (int u = gt; u < U; u += GT) {
        const int mc = u / per, cc = u - mc * per;
        const int m = mc >> 1, old = mc & 1;
        const double x = old ? sm.cur[0][m] : sm.trl[0][m], y = old ? sm.cur[1][m] : sm.trl[1][m],
                     z = old ? sm.cur[2][m] : sm.trl[2][m];
        const int tm = sm.gtype[m];
        // one candidate partner: exact FP64 minimum-image separation, the reference's r < rcut predicate
        auto lj_eval = [&](int j) {
          if (j < 0 || (j >= g0 && j < g0 + glen)) return;
          const double2 a = __ldcg(&A.xy[j]), c = __ldcg(&A.zq[j]);
          double dx = a.x - x, dy = a.y - y, dz = c.x - z;
          dx -= Lx * mv_rint(dx * iLx); dy -= Ly * mv_rint(dy * iLy); dz -= Lz * mv_rint(dz * iLz);
          const double r2 = dx * dx + dy * dy + dz * dz;
          if (r2 > ljc2max) return;
          const int tp = tm * PG_MAX_TYPES + A.type[j];
          if (r2 > Pg->lj_rcut2_relaxed[tp]) return;
          const double e = pg_pair_energy_r(*Pg, sqrt(r2), tp);
          if (old) acc_pair -= e;
          else { acc_pair += e; if (e >= PG_VLE) acc_ov += 1.0; }
        };
        if (cc < nnb) {
          int ix = ch_cell1(x, iLx, A.nc[0]), iy = ch_cell1(y, iLy, A.nc[1]), iz = ch_cell1(z, iLz, A.nc[2]);
          const int ox = (nbx == 3) ? (cc % 3) - 1 : 0;
          const int r1 = (nbx == 3) ? cc / 3 : cc;
          const int oy = (nby == 3) ? (r1 % 3) - 1 : 0;
          const int r2_ = (nby == 3) ? r1 / 3 : r1;
          const int oz = (nbz == 3) ? (r2_ % 3) - 1 : 0;
          ix += ox; iy += oy; iz += oz;
          if (ix < 0) ix += A.nc[0]; else if (ix >= A.nc[0]) ix -= A.nc[0];
          if (iy < 0) iy += A.nc[1]; else if (iy >= A.nc[1]) iy -= A.nc[1];
          if (iz < 0) iz += A.nc[2]; else if (iz >= A.nc[2]) iz -= A.nc[2];
          const int cidx = (ix * A.nc[1] + iy) * A.nc[2] + iz;
          const int4* cp = reinterpret_cast<const int4*>(A.cell_slots + (size_t)cidx * CH_CELL_CAP);
          const int4 s0 = __ldcg(cp), s1 = __ldcg(cp + 1);
          lj_eval(s0.x); lj_eval(s0.y); lj_eval(s0.z); lj_eval(s0.w);
          lj_eval(s1.x); lj_eval(s1.y); lj_eval(s1.z); lj_eval(s1.w);
        } else {
          // last unit of a bead configuration: the overflow list (beads whose cell was full), normally empty
          for (int o = 0; o < ovf_hw; o++) lj_eval(__ldcg(&A.ovf[o]));
        }
      }
    }

    // ------------------------------------------------------------------ (8) intra-molecular pairs (potential_pair.cc:157-178,
    // potential_ewald.cc:436-477): every pair of a chain move has a moved bead
    if (kind != CG_BEAD && glen > 1) {
      const int npairs = glen * (glen - 1) / 2;
      for (int p = gt; p < npairs; p += GT) {
        int jj = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
        while (jj * (jj - 1) / 2 > p) jj--;
        while ((jj + 1) * jj / 2 <= p) jj++;
        const int g = p - jj * (jj - 1) / 2;
        double dxn = sm.trl[0][jj] - sm.trl[0][g], dyn = sm.trl[1][jj] - sm.trl[1][g], dzn = sm.trl[2][jj] - sm.trl[2][g];
        double dxo = sm.cur[0][jj] - sm.cur[0][g], dyo = sm.cur[1][jj] - sm.cur[1][g], dzo = sm.cur[2][jj] - sm.cur[2][g];
        dxn -= Lx * mv_rint(dxn * iLx); dyn -= Ly * mv_rint(dyn * iLy); dzn -= Lz * mv_rint(dzn * iLz);
        dxo -= Lx * mv_rint(dxo * iLx); dyo -= Ly * mv_rint(dyo * iLy); dzo -= Lz * mv_rint(dzo * iLz);
        const double r2n = dxn * dxn + dyn * dyn + dzn * dzn, r2o = dxo * dxo + dyo * dyo + dzo * dzo;
        const int tp = sm.gtype[g] * PG_MAX_TYPES + sm.gtype[jj];
        const double qq = P.use_ewald ? sm.gq[g] * sm.gq[jj] : 0.0;
        const double lim = fmax(P.pair_kind == 1 ? P.ljc2max : -1.0, (qq != 0.0) ? P.rc2_relaxed : -1.0);
        if (r2n <= lim) {
          const double2 en = mv_pair_inrange(A.Pg, r2n, qq, tp);
          if (en.x >= PG_VLE) acc_ov += 1.0;
          acc_pair += en.x; acc_real += en.y;
        }
        if (r2o <= lim) {
          const double2 eo = mv_pair_inrange(A.Pg, r2o, qq, tp);
          acc_pair -= eo.x; acc_real -= eo.y;
        }
      }
    }

    // ------------------------------------------------------------------ (9) walls and bonds of the moved molecule (rank 0)
    if (rank == 0 && (P.ext_kind != 0 || P.bond_kind != 0)) {
      for (int g = tid; g < glen; g += CH_THREADS) {
        if (P.ext_kind != 0 && (kind != CG_BEAD || g == 0)) {
          const int t = sm.gtype[g];
          const double en = pg_wall_energy(*A.Pg, sm.trl[2][g], t);
          if (en >= PG_VLE) w_out = 1.0;
          w_sum += en - pg_wall_energy(*A.Pg, sm.cur[2][g], t);
        }
        if (P.bond_kind != 0 && g + 1 < glen)
          b_sum += pg_bond_energy(*A.Pg, sm.trl[0][g], sm.trl[1][g], sm.trl[2][g], sm.trl[0][g + 1], sm.trl[1][g + 1], sm.trl[2][g + 1]) -
                   pg_bond_energy(*A.Pg, sm.cur[0][g], sm.cur[1][g], sm.cur[2][g], sm.cur[0][g + 1], sm.cur[1][g + 1], sm.cur[2][g + 1]);
      }
    }

    // ------------------------------------------------------------------ (10) sums: warp -> CTA -> cluster, fixed order
    {
      double v[CH_NACC] = {acc_pair, acc_real, acc_rec, acc_ov, w_sum, b_sum, w_out, 0.0};
#pragma unroll
      for (int i = 0; i < 7; i++) v[i] = ch_warp_sum(v[i]);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < CH_NACC; i++) sm.red[warp][i] = v[i];
      }
      __syncthreads();
      if (tid < CH_NACC) {
        double s = 0.0;
        for (int w = 0; w < CH_WARPS; w++) s += sm.red[w][tid];
        if (G > 1) {
          for (int r = 0; r < G; r++) *cluster.map_shared_rank(&sm.part[par][rank][tid], r) = s;
        } else {
          sm.part[par][0][tid] = s;
        }
      }
      if (G > 1) cluster.sync(); else __syncthreads();
      if (tid < CH_NACC) {
        double s = 0.0;
        for (int r = 0; r < G; r++) s += sm.part[par][r][tid];
        sm.tot[tid] = s;
      }
      par ^= 1;
      __syncthreads();
    }

    // ------------------------------------------------------------------ (11) ForceField::EnergyDifference's orchestration and
    // the Metropolis test, in every CTA alike (same sums, same stream)
    double d_pair = 0, d_real = 0, d_recip = 0, d_ext = 0, d_bond = 0, d_ewald = 0, dE = 0;
    int stage = 0;
    if (tid == 0) {
      d_pair = sm.tot[0]; d_real = sm.tot[1];
      d_recip = P.use_ewald ? P.recip_pref * sm.tot[2] : 0.0;
      d_ext = sm.tot[4]; d_bond = sm.tot[5];
      if (sm.tot[6] > 0.0) d_ext = PG_VLE;   // potential_external.cc:107-110
      bool done = false;
      if (P.pair_kind != 0) { dE += d_pair; if (dE >= PG_VLE) { stage = 1; done = true; } }
      if (!done && P.ext_kind != 0) { dE += d_ext; if (dE >= PG_VLE) { stage = 2; done = true; } }
      if (!done) {
        if (P.use_ewald) { d_ewald = d_real + d_recip; dE += d_ewald; }
        if (P.bond_kind != 0) dE += d_bond;
      }
      int accept = 0;
      if (dE < PG_VLE) {                     // simulation.cc:327-332: the variate is drawn only here
        const double uacc = cg_uniform_of(cg_raw(sm.mt, 0));
        cg_advance(sm.mt, 1);
        accept = uacc < exp(-P.beta * dE);
      }
      sm.accept = accept;
      sm.dE = dE;
      if (rank == 0) {
        PgChainRec r; r.dE = dE; r.mol = sm.step.mol; r.info = (kind & 0xff) | (accept << 8) | (stage << 16);
        A.log[step_i] = r;
        if (accept) {
          E_pair += d_pair; E_ewald += d_ewald; E_bond += d_bond; E_ext += d_ext; E_real += d_real; E_recip += d_recip;
        }
      }
    }
    __syncthreads();
    ch_mt_fix(sm.mt, tid);

    // ------------------------------------------------------------------ (12) FinalizeEnergies: an accepted move becomes the state
    if (sm.accept) {
      if (nq > 0 && nk > 0) {
#pragma unroll
        for (int i = 0; i < CH_KPT; i++) {
          const int k = (int)k_lo + tid + i * CH_THREADS;
          if (i < nkt && k < (int)k_hi) {
            double2 S = __ldcg(&A.S[k]);
            const double2 d = __ldcg(&A.dS[k]);
            S.x += d.x; S.y += d.y;
            __stcg(&A.S[k], S);
          }
        }
      }
      if (rank == 0) {
        for (int i = tid; i < glen; i += CH_THREADS) {
          if (kind == CG_BEAD && i != 0) continue;
          const int jg = g0 + i;
          const double x = sm.trl[0][i], y = sm.trl[1][i], z = sm.trl[2][i], q = sm.gq[i];
          __stcg(&A.xy[jg], make_double2(x, y));
          __stcg(&A.zq[jg], make_double2(z, q));
          if (q != 0.0 && P.use_ewald) {
            const int s = A.qslot[jg];
            __stcg(&A.qpos[2 * s], make_double2(x, y));
            __stcg(&A.qpos[2 * s + 1], make_double2(z, q));
            __stcg(&A.qfrac[s], make_float4(mv_frac(x, iLx), mv_frac(y, iLy), mv_frac(z, iLz), __int_as_float(jg)));
          }
        }
        // cell grid: one warp, 32 beads at a time; beads that enter the same cell take its free slots in bead order
        if (warp == 0 && P.pair_kind == 1) {
          const int n_mv = (kind == CG_BEAD) ? 1 : glen;
          int hw = ovf_hw;
          for (int base = 0; base < n_mv; base += 32) {
            const int i = base + lane;
            const bool active = i < n_mv;
            const int bead = g0 + i;
            int newc = -1, oldc = -1;
            if (active) {
              newc = (ch_cell1(sm.trl[0][i], iLx, A.nc[0]) * A.nc[1] + ch_cell1(sm.trl[1][i], iLy, A.nc[1])) * A.nc[2] +
                     ch_cell1(sm.trl[2][i], iLz, A.nc[2]);
              oldc = __ldcg(&A.bead_cell[bead]);
            }
            const bool changed = active && newc != oldc;
            if (changed) {
              const int os = __ldcg(&A.bead_slot[bead]);
              if (os < CH_CELL_CAP) __stcg(&A.cell_slots[(size_t)oldc * CH_CELL_CAP + os], -1);
              else __stcg(&A.ovf[os - CH_CELL_CAP], -1);
            }
            __syncwarp();
            unsigned todo = __ballot_sync(0xffffffffu, changed);
            while (todo) {
              const int leader = __ffs(todo) - 1;
              const int c = __shfl_sync(0xffffffffu, newc, leader);
              const unsigned grp = __ballot_sync(0xffffffffu, changed && newc == c) & todo;
              int sv = 0;
              if (lane < CH_CELL_CAP) sv = __ldcg(&A.cell_slots[(size_t)c * CH_CELL_CAP + lane]);
              const unsigned empty = __ballot_sync(0xffffffffu, lane < CH_CELL_CAP && sv < 0);
              const int n_empty = __popc(empty), n_grp = __popc(grp);
              if ((grp >> lane) & 1u) {
                const int r = __popc(grp & lt);
                if (r < n_empty) {
                  const int slot = __fns(empty, 0, r + 1);
                  __stcg(&A.cell_slots[(size_t)c * CH_CELL_CAP + slot], bead);
                  __stcg(&A.bead_slot[bead], slot);
                } else {
                  const int idx = hw + (r - n_empty);
                  if (idx < CH_OVF_CAP) { __stcg(&A.ovf[idx], bead); __stcg(&A.bead_slot[bead], CH_CELL_CAP + idx); }
                  else __stcg(&A.out[2], (int)CH_ERR_OVERFLOW);
                }
                __stcg(&A.bead_cell[bead], c);
              }
              if (n_grp > n_empty) hw = min(hw + (n_grp - n_empty), CH_OVF_CAP);
              todo &= ~grp;
              __syncwarp();
            }
          }
          if (lane == 0 && hw != ovf_hw) __stcg(A.ovf_n, hw);
        }
      }
      __threadfence();
      if (G > 1) cluster.sync(); else __syncthreads();
      ovf_hw = __ldcg(A.ovf_n);
      if (tid == 0 && __ldcg(&A.out[2]) != 0) sm.err = CH_ERR_OVERFLOW;   // raised by rank 0, seen by every CTA alike
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  __syncthreads();
  if (rank == 0) {
    for (int i = tid; i < CG_N; i += CH_THREADS) A.mt_io[i] = sm.mt.x[sm.mt.cur][i];
    if (tid == 0) {
      A.mt_io[CG_N] = (uint32_t)sm.mt.p;
      PgState* st = A.state;
      st->E_pair = E_pair; st->E_ewald = E_ewald; st->E_bond = E_bond; st->E_ext = E_ext; st->E_real = E_real; st->E_recip = E_recip;
      A.out[0] = step_i; A.out[1] = sm.stop; A.out[3] = ovf_hw;
      if (sm.err) A.out[2] = sm.err;
    }
  }
  if (G > 1) cluster.sync();
}

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
  // device
  int *d_mol_first = nullptr, *d_chains = nullptr, *d_ions = nullptr;
  size_t mol_cap = 0;
  int nc[3] = {1, 1, 1};
  int *d_cell_slots = nullptr, *d_ovf = nullptr, *d_ovf_n = nullptr, *d_bead_cell = nullptr, *d_bead_slot = nullptr, *d_qslot = nullptr;
  size_t cell_cap = 0, bead_cap = 0;
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
};
