// plum_b200 — k_move: the per-trial-move energy change as ONE launch per move.
//
// Same contract as k_delta in MOVE mode (ForceField::EnergyDifference,
// src/force_field/force_field.cc:407-434 of the reference) plus what takes the fixed costs
// off the Markov chain's critical path:
//
//  * deferred commit — the decision on the PREVIOUS trial (FinalizeEnergies,
//    force_field.cc:436-451) rides along as an overlay: every CTA reads the previous
//    trial's coordinates in place of the resident ones if it was accepted, each k-thread
//    folds dS_prev(k) into its own S(k) before using it, and the last CTA to finish writes
//    the coordinates back and updates the running totals.  One launch per move instead of
//    two, and pg_commit() costs nothing on the host.
//  * a slim hot loop — for systems where only the central image can reach the real-space
//    cutoff (rc < L/2, the 22k-bead benchmark system) the moved-bead x partner loop is a
//    branch-light FP64 distance filter (wrap by add-magic rounding, r^2, one compare);
//    the erfc / LJ arithmetic sits behind it in a non-inlined function that only the few
//    in-range pairs call.  Everything else (multi-image sums, hard spheres, padded slab
//    boxes, intra-molecular pairs) goes through the exact generic routines of pg_math.cuh.
//  * one wave, two roles — the grid is at most the number of CTAs resident at once on the
//    148 SMs.  For a big move the first 148 CTAs (one per SM, block indices are dealt round-robin
//    over the SMs) are "helpers": each takes an equal slice of the k vectors (several lanes per k,
//    so no thread carries a long serial sincos chain while its neighbours saturate the FP64 pipe)
//    and of the intra-molecular pairs (one pair per thread).  All other CTAs split the
//    (partner tile x moved bead) units evenly, so every SM carries the same load; there is no tail
//    wave and no straggler CTA.
//  * result first — the last CTA writes dE to the host mailbox as self-validating 16-byte
//    records (value + sequence number in one store, no system-wide fence) before it does
//    the bookkeeping the host is not waiting for.
//
// No tensor cores (FP64 pair arithmetic with data-dependent branches), no TMA: the whole
// working set (40 B/bead) is L2-resident and each thread streams exactly one partner.
#include "pg_kernels.cuh"

#ifndef MV_THREADS
#define MV_THREADS 128   // measured: 128-thread CTAs beat 256 (less waiting for the slowest warp at the CTA's barriers) and 64
#endif
#define MV_WARPS (MV_THREADS / 32)
#define MV_GCHUNK 32      // max group beads per CTA chunk (shared-memory staging)
#define MV_NSLOT 8        // mailbox records
#define MV_INL 8          // largest group whose trial data ride in the kernel parameters
#ifndef MV_FAST_OCC
#define MV_FAST_OCC (1024 / MV_THREADS)   // 64 registers per thread
#endif
#define MV_GEN_OCC (512 / MV_THREADS)
#define MV_QCAP 192       // per-warp queue capacity (flushed when fewer than 64 slots remain)

// One self-validating mailbox record: written with a single 16-byte store, so a reader that
// sees `seq` sees `value` and `aux` of the same launch.
struct __align__(16) PgMailRec {
  double value;
  unsigned int seq;
  int aux;
};
// slot: 0 dE (aux = stage | accept << 8), 1 pair (aux = n_overlap), 2 ext, 3 ewald, 4 bond,
//       5 real, 6 recip, 7 mz_current

// The scalars k_move needs on its critical path, passed by value (one or two constant-bank lines)
// instead of the 3.7 KB PgDev block, whose scattered first-touch misses cost a microsecond per CTA.
struct PgMoveDev {
  double box[3], inv_box[3], ebox[3], inv_ebox[3];
  double rc2_relaxed, ljc2max, recip_pref, dipole_pref, beta;
  float fbox[3];        // box lengths in FP32 for the pre-filter
  // generic path (multi-image real space, padded slab box, hard spheres)
  double half_box[3], real_cutoff, rc_relaxed, lB, sqrt_alpha;
  int same_box[3];
  int img_cx, img_cy, img_cz;   // periodic images per axis on each side (0, 0, 0: minimum image only)
  int img_split, img_ny;        // threads per partner = x-images * y-images, each runs the z column of its (ix, iy)
  int pbc[3];
  int pair_kind, use_ewald, dipole, bond_kind, ext_kind;
};

struct PgMoveArgs {
  PgMoveDev D;
  double2* xy; double2* zq; int* type; int n;
  // current trial (device pointers into the staging block)
  int g0, glen;
  const double* trial; const double* gq; const int* gtype; const uint8_t* moved;
  const int* qidx; int nq;   // group-relative indices of the moved AND charged beads (built by the host)
  int lpk;                   // lanes per k vector in the reciprocal part: power of two, 2..32
  // small groups (glen <= MV_INL, e.g. every single-ion move) travel inside the kernel parameters:
  // no H2D copy in front of the launch.  inl_n = glen or 0; pinl_n likewise for the previous trial.
  int inl_n, pinl_n;
  double inl_trial[3 * MV_INL], inl_gq[MV_INL], pinl_trial[3 * MV_INL];
  int inl_gtype[MV_INL], inl_qidx[MV_INL];
  unsigned char inl_moved[MV_INL];
  // previous trial whose commit is still pending
  int prev_valid;
  int prev_accept;      // 0/1: decided by the host; -1: decided on the device (state->accept)
  int pg0, pglen;
  const double* ptrial;
  // reciprocal space
  const double4* kvec;  // [nk] (kx, ky, kz, ek2), half space, built on the host with the reference's expressions
  double2* S; double2* dS; int nk;
  // tiling
  int n_tiles;          // partner tiles of MV_THREADS beads
  int n_helpers, n_pair_ctas;   // CTA roles in block-index order: helpers (k slice + intra slice), then pair (move_grid)
  double* partial;      // [n_ctas][8]
  PgState* state;
  PgMailRec* mail;      // mapped host memory (NULL in replay)
  int decide_on_device; double u; double* replay_dE; uint8_t* replay_acc; int replay_index;
  unsigned int seq;
  int* stop;            // batched MC (pg_mc_run): non-zero once a step of the batch returned dE >= 1e8 — the reference draws
                        // no acceptance variate then (simulation.cc:327-332), so the host's pre-drawn stream no longer lines
                        // up and every later kernel of the batch is a no-op; NULL otherwise
  unsigned long long* timing;   // debug only (PLUM_B200_TIMING=1): per-CTA %globaltimer stamps [n_ctas][8]
  const PgDev* Pg;      // copy of the parameter block in global memory, for the non-inlined rare paths
                        // (passing the by-value kernel parameter by reference would spill all of it to local memory)
};

// Balanced partition of n items over G workers in 32-bit arithmetic (64-bit integer division is
// emulated with hundreds of instructions): worker w gets [lo, hi).
__device__ __forceinline__ void mv_share(unsigned n, unsigned G, unsigned w, unsigned& lo, unsigned& hi) {
  const unsigned q = n / G, r = n - q * G;
  lo = q * w + min(w, r);
  hi = lo + q + (w < r ? 1u : 0u);
}

__device__ __forceinline__ unsigned long long mv_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define MV_STAMP(slot) do { if (A.timing && threadIdx.x == 0) A.timing[8 * blockIdx.x + (slot)] = mv_now(); } while (0)

// add-magic round-to-nearest-even, valid for |x| < 2^51 (two FP64 adds at full FP64 rate)
__device__ __forceinline__ double mv_rint(double x) {
  const double M = 6755399441055744.0;
  return (x + M) - M;
}

// FP32 helpers of the pre-filter: box fraction of a coordinate in [-0.5, 0.5], float rint by add-magic,
// and the relaxed squared radius (sqrt(c) + delta)^2 that absorbs the FP32 error (negative: never).
__device__ __forceinline__ float mv_frac(double x, double invL) {
  const double s = x * invL;
  return (float)(s - mv_rint(s));
}
__device__ __forceinline__ float mv_rintf(float x) {
  const float M = 12582912.0f;   // 1.5 * 2^23
  return (x + M) - M;
}
__device__ __forceinline__ float mv_relax_f(double c2, double delta) {
  if (c2 < 0.0) return -1.0f;
  const double r = sqrt(c2) + delta;
  return (float)(r * r * (1.0 + 1e-6));
}

__device__ __forceinline__ void mv_mail(PgMailRec* m, int slot, double value, unsigned int seq, int aux) {
  // one 16-byte store
  asm volatile("st.volatile.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(&m[slot]), "r"(__double2loint(value)),
               "r"(__double2hiint(value)), "r"((int)seq), "r"(aux)
               : "memory");
}

// The rare in-range pair: exact LJ / real-space arithmetic for one configuration whose squared
// separation r2 passed one of the relaxed filters.  Single central image, no fold needed.
// Returns (e_lj, e_real) in registers (reference parameters would go through local memory).
__device__ __noinline__ double2 mv_pair_inrange(const PgDev* __restrict__ Pg, double r2, double qq, int tp) {
  const PgDev& P = *Pg;
  double e_lj = 0.0, e_real = 0.0;
  const double r = sqrt(r2);
  if (P.pair_kind == 1 && r2 <= P.lj_rcut2_relaxed[tp]) e_lj = pg_pair_energy_r(P, r, tp);
  if (qq != 0 && r2 <= P.rc2_relaxed) {
    if (r > 0 && r <= P.real_cutoff) e_real = P.lB * qq * erfc(P.sqrt_alpha * r) / r;
  }
  return make_double2(e_lj, e_real);
}

// LJ / hard-sphere energy from the (folded) squared separation, tables from the global parameter copy.
__device__ __noinline__ double mv_lj_r2(const PgDev* __restrict__ Pg, double l2, int tp) {
  return pg_pair_energy_r(*Pg, sqrt(l2), tp);
}

// Generic exact path: one configuration of a pair (a = group bead, p = partner), restricted to ONE
// column of periodic images (ix, iy fixed by `sub`, all iz) for the Ewald real-space term
// (potential_ewald_coul.cc real-space image loops, pruned to the images that can reach the cutoff; the
// reference's own `r <= real_cutoff` test decides).  Thread sub == 0 also owns the LJ / hard-sphere term
// (BBDist with the |d| > L/2 fold).  Returns (pair energy, real-space energy).
__device__ __forceinline__ double2 mv_generic_column(const PgMoveDev& P, const PgDev* __restrict__ Pg, double ax, double ay,
                                                     double az, double px, double py, double pz, double qq, int tp,
                                                     int sub, bool do_lj) {
  const bool wx = P.pbc[0] != 0, wy = P.pbc[1] != 0, wz = P.pbc[2] != 0;
  double e_lj = 0.0, e_re = 0.0;
  // GetDistVector: partner - group bead, wrapped in the (padded) Ewald box
  double dx = px - ax, dy = py - ay, dz = pz - az;
  if (wx) dx -= P.ebox[0] * rint(dx * P.inv_ebox[0]);
  if (wy) dy -= P.ebox[1] * rint(dy * P.inv_ebox[1]);
  if (wz) dz -= P.ebox[2] * rint(dz * P.inv_ebox[2]);
  if (sub == 0 && do_lj) {
    double lx = dx, ly = dy, lz = dz;
    if (!P.same_box[0]) lx = wx ? pg_wrap(ax - px, P.box[0], P.inv_box[0]) : ax - px;
    if (!P.same_box[1]) ly = wy ? pg_wrap(ay - py, P.box[1], P.inv_box[1]) : ay - py;
    if (!P.same_box[2]) lz = wz ? pg_wrap(az - pz, P.box[2], P.inv_box[2]) : az - pz;
    if (wx && fabs(lx) > P.half_box[0]) lx = P.box[0] - fabs(lx);
    if (wy && fabs(ly) > P.half_box[1]) ly = P.box[1] - fabs(ly);
    if (wz && fabs(lz) > P.half_box[2]) lz = P.box[2] - fabs(lz);
    const double l2 = lx * lx + ly * ly + lz * lz;
    if (P.pair_kind == 2 || l2 <= P.ljc2max) e_lj = mv_lj_r2(Pg, l2, tp);
  }
  if (P.use_ewald && qq != 0.0) {
    const int ix = sub / P.img_ny - P.img_cx, iy = sub % P.img_ny - P.img_cy;
    const double rx = dx + ix * P.ebox[0], ry = dy + iy * P.ebox[1];
    const double rxy2 = rx * rx + ry * ry;
    const double rc2r = P.rc2_relaxed;
    if (rxy2 <= rc2r) {
      const double rcr = P.rc_relaxed, rc = P.real_cutoff, pref = P.lB * qq, sa = P.sqrt_alpha;
      int k0 = (int)ceil((-rcr - dz) * P.inv_ebox[2]), k1 = (int)floor((rcr - dz) * P.inv_ebox[2]);
      k0 = max(k0, -P.img_cz); k1 = min(k1, P.img_cz);
      for (int k = k0; k <= k1; k++) {
        const double rz = dz + k * P.ebox[2];
        const double r2 = rxy2 + rz * rz;
        if (r2 > rc2r) continue;
        const double r = sqrt(r2);
        if (r > 0 && r <= rc) e_re += pref * erfc(sa * r) / r;
      }
    }
  }
  return make_double2(e_lj, e_re);
}

// Evaluate this warp's queued in-range configurations, one per lane, and add them to the lane's
// accumulators (entries are in loop order and lane e takes entries e, e+32, ...: deterministic).
// A macro so that the queues are addressed as shared memory (LDS), not through generic pointers.
#define MV_QUEUE_FLUSH()                                                              \
  do {                                                                                \
    __syncwarp();                                                                     \
    for (int e_ = lane; e_ < qn; e_ += 32) {                                          \
      const int t_ = q_tp[warp][e_];                                                  \
      const double2 en_ = mv_pair_inrange(A.Pg, q_r2[warp][e_], q_qq[warp][e_], t_ & 0xffff); \
      if (t_ & 0x10000) {                                                             \
        acc_pair -= en_.x; acc_real -= en_.y;                                         \
      } else {                                                                        \
        if (en_.x >= PG_VLE) acc_ov += 1.0;                                           \
        acc_pair += en_.x; acc_real += en_.y;                                         \
      }                                                                               \
    }                                                                                 \
    __syncwarp();                                                                     \
    qn = 0;                                                                           \
  } while (0)

// Sum 5 values over the CTA with one barrier; result valid in thread 0.
__device__ __forceinline__ void mv_block_sum5(double (&v)[5], double* smem /* [MV_WARPS][5] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Most warps of a move have nothing but their partners' dipole moments (v[2]) to contribute: no pair of theirs was in
  // range.  A warp whose other four accumulators are zero in every lane skips their shuffle trees (the sums are zero).
  const bool any_nz = __any_sync(0xffffffffu, (v[0] != 0.0) | (v[1] != 0.0) | (v[3] != 0.0) | (v[4] != 0.0));
  v[2] = warp_sum(v[2]);
  if (any_nz) {
    v[0] = warp_sum(v[0]); v[1] = warp_sum(v[1]); v[3] = warp_sum(v[3]); v[4] = warp_sum(v[4]);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 5; i++) smem[warp * 5 + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < MV_WARPS; w++) s += smem[w * 5 + i];
      v[i] = s;
    }
  }
}

template <bool FAST>
__global__ void __launch_bounds__(MV_THREADS, FAST ? MV_FAST_OCC : MV_GEN_OCC) k_move(const PgMoveArgs A) {
  const PgMoveDev& P = A.D;
  __shared__ double s_n[3][MV_GCHUNK], s_o[3][MV_GCHUNK], s_q[MV_GCHUNK];
  __shared__ int s_t[MV_GCHUNK], s_mv[MV_GCHUNK];
  __shared__ double s_cut[2][MV_GCHUNK];
  // FP32 pre-filter copies: box-fraction coordinates of the new / old position + relaxed radii^2
  // (.w of s_fn: against a charged partner, .w of s_fo: against a neutral partner)
  __shared__ float4 s_fn[MV_GCHUNK], s_fo[MV_GCHUNK];
  // per-warp queues of in-range configurations (FAST path): r^2, q1*q2, type pair | old flag
  __shared__ double q_r2[FAST ? MV_WARPS : 1][FAST ? MV_QCAP : 1], q_qq[FAST ? MV_WARPS : 1][FAST ? MV_QCAP : 1];
  __shared__ int q_tp[FAST ? MV_WARPS : 1][FAST ? MV_QCAP : 1];
  __shared__ double s_red[MV_WARPS * 5];
  __shared__ int s_last;

  __shared__ double s_it[3 * MV_INL], s_iq[MV_INL], s_pit[3 * MV_INL];
  __shared__ int s_ity[MV_INL], s_iqi[MV_INL];
  __shared__ unsigned char s_imv[MV_INL];

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  if (A.stop && __ldcg(A.stop)) return;
  MV_STAMP(0);
  const double* __restrict__ trial = A.trial;
  const double* __restrict__ gqv = A.gq;
  const int* __restrict__ gtypev = A.gtype;
  const int* __restrict__ qidxv = A.qidx;
  const unsigned char* __restrict__ movedv = A.moved;
  const double* __restrict__ ptrial = A.ptrial;
  if (A.inl_n | A.pinl_n) {
    if (tid < 3 * MV_INL) { s_it[tid] = A.inl_trial[tid]; s_pit[tid] = A.pinl_trial[tid]; }
    if (tid < MV_INL) {
      s_iq[tid] = A.inl_gq[tid]; s_ity[tid] = A.inl_gtype[tid]; s_iqi[tid] = A.inl_qidx[tid];
      s_imv[tid] = A.inl_moved[tid];
    }
    __syncthreads();
    if (A.inl_n) { trial = s_it; gqv = s_iq; gtypev = s_ity; qidxv = s_iqi; movedv = s_imv; }
    if (A.pinl_n) ptrial = s_pit;
  }
  // decision on the previous trial (uniform over the grid)
  int apply_prev = 0;
  if (A.prev_valid) apply_prev = (A.prev_accept >= 0) ? A.prev_accept : __ldcg(&A.state->accept);
  const int pg0 = A.pg0, pg1 = apply_prev ? A.pg0 + A.pglen : A.pg0;   // empty range when not applied

  double acc_pair = 0.0, acc_real = 0.0, acc_mz = 0.0, acc_rec = 0.0, acc_ov = 0.0;

  // ---------------- (a) reciprocal space: k vectors [k0, k1) of this CTA, LPK lanes per k.
  // Lane e of a k group handles the terms e, e+LPK, ... of the 2*nq list (bead, new|old); a shuffle
  // tree adds them up; lane 0 owns S(k): it folds the previous trial's dS in first (deferred commit).
  if (b < A.n_helpers && A.nk > 0) {
    unsigned k0u, k1u;
    mv_share((unsigned)A.nk, (unsigned)A.n_helpers, (unsigned)b, k0u, k1u);
    const int k0 = (int)k0u, k1 = (int)k1u;
    const int LPK = A.lpk, KPP = MV_THREADS / LPK;   // k vectors per pass
    const int kl = tid / LPK, e0 = tid % LPK;
    for (int kb = k0; kb < k1; kb += KPP) {
      const int k = kb + kl;
      const bool active = k < k1;
      double dre = 0.0, dim = 0.0;
      double4 kv = make_double4(0.0, 0.0, 0.0, 0.0);
      if (active) {
        kv = A.kvec[k];
        for (int e = e0; e < 2 * A.nq; e += LPK) {
          const int g = qidxv[e >> 1];
          const double q = gqv[g];
          double x, y, z;
          if (e & 1) {   // old configuration, through the overlay
            const int jg = A.g0 + g;
            if (jg >= pg0 && jg < pg1) {
              x = ptrial[3 * (jg - pg0)]; y = ptrial[3 * (jg - pg0) + 1]; z = ptrial[3 * (jg - pg0) + 2];
            } else {
              const double2 a = A.xy[jg], c = A.zq[jg];
              x = a.x; y = a.y; z = c.x;
            }
          } else {
            x = trial[3 * g]; y = trial[3 * g + 1]; z = trial[3 * g + 2];
          }
          x = pg_wrap_pos(x, P.ebox[0], P.inv_ebox[0], P.pbc[0]);
          y = pg_wrap_pos(y, P.ebox[1], P.inv_ebox[1], P.pbc[1]);
          z = pg_wrap_pos(z, P.ebox[2], P.inv_ebox[2], P.pbc[2]);
          double sn, cs;
          sincos(kv.x * x + kv.y * y + kv.z * z, &sn, &cs);
          const double w = (e & 1) ? -q : q;
          dre += w * cs; dim += w * sn;
        }
      }
      for (int o = LPK >> 1; o > 0; o >>= 1) {
        dre += __shfl_xor_sync(0xffffffffu, dre, o);
        dim += __shfl_xor_sync(0xffffffffu, dim, o);
      }
      if (active && e0 == 0) {
        double2 S = A.S[k];
        if (apply_prev) {
          const double2 d = A.dS[k];
          S.x += d.x; S.y += d.y;
          A.S[k] = S;
        }
        // never |S_new|^2 - |S_old|^2: 2 Re(conj(S) dS) + |dS|^2; x2 for the -k half
        acc_rec += 2.0 * kv.w * (2.0 * (S.x * dre + S.y * dim) + (dre * dre + dim * dim));
        A.dS[k] = make_double2(dre, dim);
      }
    }
  }

  // ---------------- (b) intra-molecular pairs (g, jj), g < jj, of the moved molecule: this CTA's
  // slice [p0, p1), one pair per thread (potential_pair.cc:157-178, potential_ewald.cc:436-477;
  // the g == jj self-image term is identical before and after a move)
  if (b < A.n_helpers) {
    unsigned p0, p1;
    // generic path: img_split threads per pair, one image column each
    const int isplit = FAST ? 1 : P.img_split;
    mv_share((unsigned)(A.glen * (A.glen - 1) / 2 * isplit), (unsigned)A.n_helpers, (unsigned)b, p0, p1);
    for (unsigned pp = p0 + tid; pp < p1; pp += MV_THREADS) {
      const int p = FAST ? (int)pp : (int)pp / isplit;
      const int sub = FAST ? 0 : (int)pp - p * isplit;
      int jj = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
      while (jj * (jj - 1) / 2 > p) jj--;
      while ((jj + 1) * jj / 2 <= p) jj++;
      const int g = p - jj * (jj - 1) / 2;
      if (!(movedv[g] || movedv[jj])) continue;
      const int ja = A.g0 + g, jb = A.g0 + jj;
      double aox, aoy, aoz, box_, boy, boz;
      if (ja >= pg0 && ja < pg1) {
        aox = ptrial[3 * (ja - pg0)]; aoy = ptrial[3 * (ja - pg0) + 1]; aoz = ptrial[3 * (ja - pg0) + 2];
      } else {
        double2 a = A.xy[ja], c = A.zq[ja];
        aox = a.x; aoy = a.y; aoz = c.x;
      }
      if (jb >= pg0 && jb < pg1) {
        box_ = ptrial[3 * (jb - pg0)]; boy = ptrial[3 * (jb - pg0) + 1]; boz = ptrial[3 * (jb - pg0) + 2];
      } else {
        double2 a = A.xy[jb], c = A.zq[jb];
        box_ = a.x; boy = a.y; boz = c.x;
      }
      const double anx = trial[3 * g], any_ = trial[3 * g + 1], anz = trial[3 * g + 2];
      const double bnx = trial[3 * jj], bny = trial[3 * jj + 1], bnz = trial[3 * jj + 2];
      const double qa = gqv[g], qb = gqv[jj];
      const int ta = gtypev[g], tb = gtypev[jj];
      double lj_n = 0.0, re_n = 0.0, lj_o = 0.0, re_o = 0.0;
      if (FAST) {
        const double Lx = P.box[0], Ly = P.box[1], Lz = P.box[2];
        const double iLx = P.inv_box[0], iLy = P.inv_box[1], iLz = P.inv_box[2];
        double dxn = bnx - anx, dyn = bny - any_, dzn = bnz - anz;
        double dxo = box_ - aox, dyo = boy - aoy, dzo = boz - aoz;
        dxn -= Lx * mv_rint(dxn * iLx); dyn -= Ly * mv_rint(dyn * iLy); dzn -= Lz * mv_rint(dzn * iLz);
        dxo -= Lx * mv_rint(dxo * iLx); dyo -= Ly * mv_rint(dyo * iLy); dzo -= Lz * mv_rint(dzo * iLz);
        const double r2n = dxn * dxn + dyn * dyn + dzn * dzn;
        const double r2o = dxo * dxo + dyo * dyo + dzo * dzo;
        const int tp = ta * PG_MAX_TYPES + tb;
        const double qq = qa * qb;
        const double2 en = mv_pair_inrange(A.Pg, r2n, qq, tp), eo = mv_pair_inrange(A.Pg, r2o, qq, tp);
        lj_n = en.x; re_n = en.y; lj_o = eo.x; re_o = eo.y;
      } else {
        int lj_here = (P.pair_kind != 0);
        if (P.pair_kind == 2 && jj == g + 1) lj_here = 0;   // bonded hard spheres, potential_pair.cc:165-169
        const int tp = ta * PG_MAX_TYPES + tb;
        const double2 en = mv_generic_column(P, A.Pg, anx, any_, anz, bnx, bny, bnz, qa * qb, tp, sub, lj_here != 0);
        const double2 eo = mv_generic_column(P, A.Pg, aox, aoy, aoz, box_, boy, boz, qa * qb, tp, sub, lj_here != 0);
        lj_n = en.x; re_n = en.y; lj_o = eo.x; re_o = eo.y;
      }
      if (lj_n >= PG_VLE) acc_ov += 1.0;
      acc_pair += (lj_n - lj_o);
      acc_real += (re_n - re_o);
    }
  }

  // ---------------- (c) moved beads x partners.  The work is n_tiles x glen equal "units" (one
  // partner tile against one moved bead); this CTA owns the contiguous unit range [u0, u1) in
  // tile-major order: one tile segment, or the tail of one tile and the head of the next.
  else {
    unsigned u, u1;
    mv_share((unsigned)(A.n_tiles * A.glen), (unsigned)A.n_pair_ctas, (unsigned)(b - A.n_helpers), u, u1);
    const double Lx = P.box[0], Ly = P.box[1], Lz = P.box[2];
    const double iLx = P.inv_box[0], iLy = P.inv_box[1], iLz = P.inv_box[2];
    const double rc2 = P.use_ewald ? P.rc2_relaxed : -1.0;
    const int do_lj = (P.pair_kind != 0);
    const double ljc2max = (P.pair_kind == 1) ? P.ljc2max : -1.0;
    const float fLx = P.fbox[0], fLy = P.fbox[1], fLz = P.fbox[2];
    const double fdelta = 1e-6 * fmax(Lx, fmax(Ly, Lz));   // >= 3x the worst-case FP32 error of a separation
    const int lane = tid & 31, warp = tid >> 5;
    int qn = 0;   // entries in this warp's queue (warp-uniform)
    bool first = true;
    // Partner-side state: a function of the tile only, so consecutive segments of one tile (a CTA that owns a whole
    // tile walks all the moved beads in chunks of MV_GCHUNK) load the partner, convert it to box fractions and
    // measure the warp's extent once.
    int cur_tile = -1, j = 0, sub = 0, pt = 0;
    double px = 0, py = 0, pz = 0, pq = 0;
    bool valid = false, pchg = false, anyq = false;
    float psx = 0, psy = 0, psz = 0, cx = 0, cy = 0, cz = 0, hwx = 0, hwy = 0, hwz = 0;
    unsigned vmask = 0u;
    while (u < u1) {
      const int tile = (int)(u / (unsigned)A.glen);
      const int gbeg = (int)(u - (unsigned)tile * (unsigned)A.glen);
      const int gcnt = min(min(A.glen - gbeg, MV_GCHUNK), (int)(u1 - u));
      if (tile != cur_tile) {
        cur_tile = tile;
        // partner first: its loads are the longest dependency chain of the segment
        // FAST: one partner per thread.  Generic: img_split threads per partner, one (ix, iy) image column each.
        const int slot = tile * MV_THREADS + tid;
        j = FAST ? slot : slot / P.img_split;
        sub = FAST ? 0 : slot - j * P.img_split;
        px = 0; py = 0; pz = 0; pq = 0; pt = 0;
        if (j < A.n) {
          const double2 c = A.zq[j];
          const double2 a = A.xy[j];
          pt = A.type[j];
          pq = c.y;
          px = a.x; py = a.y; pz = c.x;
          if (j >= pg0 && j < pg1) {
            px = ptrial[3 * (j - pg0)]; py = ptrial[3 * (j - pg0) + 1]; pz = ptrial[3 * (j - pg0) + 2];
          }
        }
        // partners that are beads of the moved molecule itself were handled in (b)
        valid = (j < A.n) && !((j >= A.g0) && (j < A.g0 + A.glen));
        if (FAST) {
          pchg = (pq != 0.0);
          psx = mv_frac(px, iLx); psy = mv_frac(py, iLy); psz = mv_frac(pz, iLz);
          // Warp-level culling, partner side.  The 32 partners of a warp are consecutive beads — a chain segment,
          // i.e. a compact cloud.  Its extent around one reference partner (per-axis half widths, minimum image)
          // gives a lower bound on the separation of ANY of them from a moved bead:
          //   |wrap(g - p)| >= |wrap(g - c)| - hw   (triangle inequality on the circle, per axis).
          vmask = __ballot_sync(0xffffffffu, valid);
          if (vmask) {
            const int ref = __ffs(vmask) - 1;
            cx = __shfl_sync(0xffffffffu, psx, ref); cy = __shfl_sync(0xffffffffu, psy, ref);
            cz = __shfl_sync(0xffffffffu, psz, ref);
            float ex = psx - cx, ey = psy - cy, ez = psz - cz;
            ex = valid ? fabsf(ex - mv_rintf(ex)) : 0.0f;
            ey = valid ? fabsf(ey - mv_rintf(ey)) : 0.0f;
            ez = valid ? fabsf(ez - mv_rintf(ez)) : 0.0f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              ex = fmaxf(ex, __shfl_xor_sync(0xffffffffu, ex, o));
              ey = fmaxf(ey, __shfl_xor_sync(0xffffffffu, ey, o));
              ez = fmaxf(ez, __shfl_xor_sync(0xffffffffu, ez, o));
            }
            hwx = ex + 1e-6f; hwy = ey + 1e-6f; hwz = ez + 1e-6f;   // + FP32 slack in box fractions
            anyq = __ballot_sync(0xffffffffu, valid && pchg) != 0u;
          }
        }
      }
      // each partner's dipole moment is counted once: by the segment that starts its tile
      if (gbeg == 0 && sub == 0 && j < A.n) acc_mz += pq * pz;
      if (!first) __syncthreads();   // previous segment's shared staging is still being read
      first = false;
      if (tid < gcnt) {
        const int g = gbeg + tid, jg = A.g0 + g;
        double ox, oy, oz;
        if (jg >= pg0 && jg < pg1) {
          ox = ptrial[3 * (jg - pg0)]; oy = ptrial[3 * (jg - pg0) + 1]; oz = ptrial[3 * (jg - pg0) + 2];
        } else {
          double2 a = A.xy[jg], c = A.zq[jg];
          ox = a.x; oy = a.y; oz = c.x;
        }
        s_o[0][tid] = ox; s_o[1][tid] = oy; s_o[2][tid] = oz;
        s_n[0][tid] = trial[3 * g]; s_n[1][tid] = trial[3 * g + 1]; s_n[2][tid] = trial[3 * g + 2];
        const double gq = gqv[g];
        s_q[tid] = gq;
        s_t[tid] = gtypev[g];
        s_mv[tid] = movedv[g];
        // conservative filter radii^2 of this moved bead: against a charged partner / a neutral one
        // (the exact per-pair cutoffs are re-applied by mv_pair_inrange)
        const double c0 = ljc2max, c1 = (gq != 0.0) ? fmax(rc2, ljc2max) : ljc2max;
        s_cut[0][tid] = c0;
        s_cut[1][tid] = c1;
        if (FAST) {
          const double nx = s_n[0][tid], ny = s_n[1][tid], nz = s_n[2][tid];
          s_fn[tid] = make_float4(mv_frac(nx, iLx), mv_frac(ny, iLy), mv_frac(nz, iLz), mv_relax_f(c1, fdelta));
          s_fo[tid] = make_float4(mv_frac(ox, iLx), mv_frac(oy, iLy), mv_frac(oz, iLz), mv_relax_f(c0, fdelta));
        }
      }
      __syncthreads();
      MV_STAMP(1);
      if (FAST) {
        // Every lane runs the filter (invalid lanes never hit) so the warp can vote: in-range
        // configurations are rare, so instead of evaluating erfc / LJ in a nearly empty diverged warp
        // they are compacted into this warp's queue and evaluated 32 at a time (mv_queue_flush).
        const int csel = pchg ? 1 : 0;
        // Warp-level culling, moved-bead side: lane i applies the bound to moved bead i (a chunk has <= 32 beads);
        // the warp then runs the per-partner filter only for the beads that survive.  It can only over-accept.
        unsigned keep = 0u;
        if (vmask) {
          bool k = false;
          if (lane < gcnt && s_mv[lane]) {
            const float4 fn = s_fn[lane], fo = s_fo[lane];
            const float cutf = anyq ? fn.w : fo.w;
            float bxn = fn.x - cx, byn = fn.y - cy, bzn = fn.z - cz;
            float bxo = fo.x - cx, byo = fo.y - cy, bzo = fo.z - cz;
            bxn = fmaxf(fabsf(bxn - mv_rintf(bxn)) - hwx, 0.0f) * fLx;
            byn = fmaxf(fabsf(byn - mv_rintf(byn)) - hwy, 0.0f) * fLy;
            bzn = fmaxf(fabsf(bzn - mv_rintf(bzn)) - hwz, 0.0f) * fLz;
            bxo = fmaxf(fabsf(bxo - mv_rintf(bxo)) - hwx, 0.0f) * fLx;
            byo = fmaxf(fabsf(byo - mv_rintf(byo)) - hwy, 0.0f) * fLy;
            bzo = fmaxf(fabsf(bzo - mv_rintf(bzo)) - hwz, 0.0f) * fLz;
            const float lbn = fmaf(bxn, bxn, fmaf(byn, byn, bzn * bzn));
            const float lbo = fmaf(bxo, bxo, fmaf(byo, byo, bzo * bzo));
            k = fminf(lbn, lbo) <= cutf;
          }
          keep = __ballot_sync(0xffffffffu, k);
        }
        while (keep) {
          const int i = __ffs(keep) - 1;
          keep &= keep - 1u;
          // FP32 pre-filter in box fractions (full-rate pipe); it can only over-accept
          const float4 fn = s_fn[i], fo = s_fo[i];
          float axn = psx - fn.x, ayn = psy - fn.y, azn = psz - fn.z;
          float axo = psx - fo.x, ayo = psy - fo.y, azo = psz - fo.z;
          axn -= mv_rintf(axn); ayn -= mv_rintf(ayn); azn -= mv_rintf(azn);
          axo -= mv_rintf(axo); ayo -= mv_rintf(ayo); azo -= mv_rintf(azo);
          axn *= fLx; ayn *= fLy; azn *= fLz;
          axo *= fLx; ayo *= fLy; azo *= fLz;
          const float f2n = fmaf(axn, axn, fmaf(ayn, ayn, azn * azn));
          const float f2o = fmaf(axo, axo, fmaf(ayo, ayo, azo * azo));
          const float fcut = pchg ? fn.w : fo.w;
          const bool fhit = valid && (fminf(f2n, f2o) <= fcut);
          if (__ballot_sync(0xffffffffu, fhit) == 0u) continue;
          // exact FP64 separations for the lanes the filter let through
          bool hn = false, ho = false;
          double r2n = 0.0, r2o = 0.0;
          if (fhit) {
            double dxn = px - s_n[0][i], dyn = py - s_n[1][i], dzn = pz - s_n[2][i];
            double dxo = px - s_o[0][i], dyo = py - s_o[1][i], dzo = pz - s_o[2][i];
            dxn -= Lx * mv_rint(dxn * iLx); dyn -= Ly * mv_rint(dyn * iLy); dzn -= Lz * mv_rint(dzn * iLz);
            dxo -= Lx * mv_rint(dxo * iLx); dyo -= Ly * mv_rint(dyo * iLy); dzo -= Lz * mv_rint(dzo * iLz);
            r2n = dxn * dxn + dyn * dyn + dzn * dzn;
            r2o = dxo * dxo + dyo * dyo + dzo * dzo;
            const double cut = s_cut[csel][i];
            hn = r2n <= cut;
            ho = r2o <= cut;
          }
          const unsigned mn = __ballot_sync(0xffffffffu, hn), mo = __ballot_sync(0xffffffffu, ho);
          if (mn | mo) {
            const int tp = s_t[i] * PG_MAX_TYPES + pt;
            const double qq = s_q[i] * pq;
            const unsigned lt = (1u << lane) - 1u;
            if (hn) {
              const int pos = qn + __popc(mn & lt);
              q_r2[warp][pos] = r2n; q_qq[warp][pos] = qq; q_tp[warp][pos] = tp;
            }
            qn += __popc(mn);
            if (ho) {
              const int pos = qn + __popc(mo & lt);
              q_r2[warp][pos] = r2o; q_qq[warp][pos] = qq; q_tp[warp][pos] = tp | 0x10000;   // old: subtract
            }
            qn += __popc(mo);
            if (qn > MV_QCAP - 64) MV_QUEUE_FLUSH();
          }
        }
      } else if (valid) {
        // Generic exact path (all four reference examples): this thread owns one column of periodic
        // images of its partner, thread sub == 0 also the LJ / hard-sphere term.
        for (int i = 0; i < gcnt; i++) {
          if (!s_mv[i]) continue;
          const double qq = s_q[i] * pq;
          const int tp = s_t[i] * PG_MAX_TYPES + pt;
          const double2 en = mv_generic_column(P, A.Pg, s_n[0][i], s_n[1][i], s_n[2][i], px, py, pz, qq, tp, sub, do_lj != 0);
          const double2 eo = mv_generic_column(P, A.Pg, s_o[0][i], s_o[1][i], s_o[2][i], px, py, pz, qq, tp, sub, do_lj != 0);
          if (en.x >= PG_VLE) acc_ov += 1.0;
          acc_pair += (en.x - eo.x);
          acc_real += (en.y - eo.y);
        }
      }
      u += gcnt;
    }
    if (FAST && qn > 0) MV_QUEUE_FLUSH();
  }

  // ---------------- per-CTA partials; the last CTA to finish finalises
  MV_STAMP(2);
  {
    double v[5] = {acc_pair, acc_real, acc_mz, acc_rec, acc_ov};
    mv_block_sum5(v, s_red);
    if (tid == 0) {
      double2* pp = reinterpret_cast<double2*>(A.partial + 8 * (size_t)b);
      pp[0] = make_double2(v[0], v[1]);
      pp[1] = make_double2(v[2], v[3]);
      pp[2] = make_double2(v[4], 0.0);
      if (A.timing) {
        A.timing[8 * b + 3] = mv_now();
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        A.timing[8 * b + 6] = smid;
      }
      // release our partials / acquire everybody else's in one atomic
      unsigned int done;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(done) : "l"(&A.state->done_counter) : "memory");
      s_last = (done == gridDim.x - 1);
      if (A.timing) A.timing[8 * b + 4] = mv_now();
    }
    __syncthreads();
    if (!s_last) return;
  }

  // ======================= finalisation (one CTA) =======================
  MV_STAMP(5);
  PgState* st = A.state;
  // state words the decision needs, fetched while the partials are being summed
  double st_cur_dipl = 0.0, st_trial_dipl = 0.0;
  if (tid == 0 && P.dipole) { st_cur_dipl = __ldcg(&st->cur_dipl); st_trial_dipl = __ldcg(&st->trial_dipl); }
  const int n_ctas = gridDim.x;
  double t_pair = 0.0, t_real = 0.0, t_mz = 0.0, t_rec = 0.0, t_ov = 0.0;
  for (int c = tid; c < n_ctas; c += MV_THREADS) {
    const double2* pp = reinterpret_cast<const double2*>(A.partial + 8 * (size_t)c);
    const double2 a = __ldcg(pp), b2 = __ldcg(pp + 1), c2 = __ldcg(pp + 2);
    t_pair += a.x; t_real += a.y; t_mz += b2.x; t_rec += b2.y; t_ov += c2.x;
  }
  double v[5] = {t_pair, t_real, t_mz, t_rec, t_ov};
  mv_block_sum5(v, s_red);
  double w_sum = 0.0, b_sum = 0.0, w_out = 0.0;
  if (P.ext_kind != 0 || P.bond_kind != 0) {
    for (int g = tid; g < A.glen; g += MV_THREADS) {
      const int jg = A.g0 + g;
      double ox, oy, oz;
      if (jg >= pg0 && jg < pg1) {
        ox = ptrial[3 * (jg - pg0)]; oy = ptrial[3 * (jg - pg0) + 1]; oz = ptrial[3 * (jg - pg0) + 2];
      } else {
        double2 a = A.xy[jg], c = A.zq[jg];
        ox = a.x; oy = a.y; oz = c.x;
      }
      if (P.ext_kind != 0 && movedv[g]) {
        const int t = gtypev[g];
        const double en = pg_wall_energy(*A.Pg, trial[3 * g + 2], t);
        if (en >= PG_VLE) w_out = 1.0;
        w_sum += en - pg_wall_energy(*A.Pg, oz, t);
      }
      if (P.bond_kind != 0 && g + 1 < A.glen) {
        const int jg2 = jg + 1;
        double ox2, oy2, oz2;
        if (jg2 >= pg0 && jg2 < pg1) {
          ox2 = ptrial[3 * (jg2 - pg0)]; oy2 = ptrial[3 * (jg2 - pg0) + 1]; oz2 = ptrial[3 * (jg2 - pg0) + 2];
        } else {
          double2 a = A.xy[jg2], c = A.zq[jg2];
          ox2 = a.x; oy2 = a.y; oz2 = c.x;
        }
        b_sum += pg_bond_energy(*A.Pg, trial[3 * g], trial[3 * g + 1], trial[3 * g + 2], trial[3 * g + 3],
                                trial[3 * g + 4], trial[3 * g + 5]) -
                 pg_bond_energy(*A.Pg, ox, oy, oz, ox2, oy2, oz2);
      }
    }
    __syncthreads();   // s_red is reused
    double v2[5] = {w_sum, b_sum, w_out, 0.0, 0.0};
    mv_block_sum5(v2, s_red);
    w_sum = v2[0]; b_sum = v2[1]; w_out = v2[2];
  }
  if (tid == 0) {
    MV_STAMP(6);
    // the dipole lag state as it will be once the previous trial is finalised
    if (A.prev_valid && P.dipole) {
      if (apply_prev) st_cur_dipl = st_trial_dipl; else st_trial_dipl = st_cur_dipl;
    }
    const double d_pair = v[0], d_real = v[1], mz_cur = v[2];
    const double d_recip = P.use_ewald ? P.recip_pref * v[3] : 0.0;
    const int n_overlap = (int)v[4];
    double d_ext = w_sum;
    const double d_bond = b_sum;
    double dE = 0.0, d_ewald = 0.0, d_dip = 0.0;
    int stage = 0;
    if (w_out > 0.0) d_ext = PG_VLE;   // potential_external.cc:107-110
    bool done = false;
    if (P.pair_kind != 0) {
      dE += d_pair;
      if (dE >= PG_VLE) { stage = 1; done = true; }
    }
    if (!done && P.ext_kind != 0) {
      dE += d_ext;
      if (dE >= PG_VLE) { stage = 2; done = true; }
    }
    if (!done) {
      if (P.use_ewald) {
        d_ewald = d_real + d_recip;
        if (P.dipole) {
          st_trial_dipl = P.dipole_pref * mz_cur * mz_cur;   // lags one accepted move
          d_dip = st_trial_dipl - st_cur_dipl;
          d_ewald += d_dip;
        }
        dE += d_ewald;
      }
      if (P.bond_kind != 0) dE += d_bond;
    }
    int accept = 0;
    if (A.decide_on_device) accept = (dE < PG_VLE) && (A.u < exp(-P.beta * dE));   // simulation.cc:327-332
    // ---- result first: the host is spinning on record 0
    if (A.mail) {
      mv_mail(A.mail, 0, dE, A.seq, stage | (accept << 8));
      mv_mail(A.mail, 1, d_pair, A.seq, n_overlap);
      mv_mail(A.mail, 2, d_ext, A.seq, 0);
      mv_mail(A.mail, 3, d_ewald, A.seq, 0);
      mv_mail(A.mail, 4, d_bond, A.seq, 0);
      mv_mail(A.mail, 5, d_real, A.seq, 0);
      mv_mail(A.mail, 6, d_recip, A.seq, 0);
      mv_mail(A.mail, 7, mz_cur, A.seq, 0);
    }
    // ---- bookkeeping nobody waits for: FinalizeEnergies of the previous trial, new pending trial
    if (A.prev_valid && apply_prev) {
      st->E_pair += st->d_pair; st->E_ewald += st->d_ewald; st->E_bond += st->d_bond; st->E_ext += st->d_ext;
      st->E_real += st->d_real; st->E_recip += st->d_recip;
    }
    if (P.dipole) { st->cur_dipl = st_cur_dipl; st->trial_dipl = st_trial_dipl; }
    if (A.replay_dE) A.replay_dE[A.replay_index] = dE;
    if (A.replay_acc) A.replay_acc[A.replay_index] = (uint8_t)accept;
    if (A.stop && dE >= PG_VLE) *A.stop = A.replay_index + 1;
    st->dE = dE; st->d_pair = d_pair; st->d_ext = d_ext; st->d_ewald = d_ewald; st->d_bond = d_bond;
    st->d_real = d_real; st->d_recip = d_recip; st->d_self = 0.0; st->d_dipole = d_dip;
    st->mz_current = mz_cur;
    st->stage = stage; st->n_overlap = n_overlap; st->accept = accept; st->mode = PG_MODE_MOVE;
    st->done_counter = 0;
  }
  // write the previous trial's coordinates back: every other CTA has finished reading them
  if (apply_prev) {
    for (int i = tid; i < A.pglen; i += MV_THREADS) {
      A.xy[A.pg0 + i] = make_double2(ptrial[3 * i], ptrial[3 * i + 1]);
      double2 c = A.zq[A.pg0 + i];
      c.x = ptrial[3 * i + 2];
      A.zq[A.pg0 + i] = c;
    }
  }
  MV_STAMP(7);
}
