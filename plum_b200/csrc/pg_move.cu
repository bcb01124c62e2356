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
//  * one wave — the host sizes (partner tile x group chunk) so the grid fits the resident
//    CTA slots of the 148 SMs once (no tail wave), see move_tiling() in pg_engine.cu.
//  * result first — the last CTA writes dE to the host mailbox as self-validating 16-byte
//    records (value + sequence number in one store, no system-wide fence) before it does
//    the bookkeeping the host is not waiting for.
//
// No tensor cores (FP64 pair arithmetic with data-dependent branches), no TMA: the whole
// working set (40 B/bead) is L2-resident and each thread streams exactly one partner.
#include "pg_kernels.cuh"

#define MV_THREADS 256
#define MV_WARPS (MV_THREADS / 32)
#define MV_GCHUNK 32      // max group beads per CTA chunk (shared-memory staging)
#define MV_NSLOT 8        // mailbox records

// One self-validating mailbox record: written with a single 16-byte store, so a reader that
// sees `seq` sees `value` and `aux` of the same launch.
struct __align__(16) PgMailRec {
  double value;
  unsigned int seq;
  int aux;
};
// slot: 0 dE (aux = stage | accept << 8), 1 pair (aux = n_overlap), 2 ext, 3 ewald, 4 bond,
//       5 real, 6 recip, 7 mz_current

struct PgMoveArgs {
  double2* xy; double2* zq; int* type; int n;
  // current trial (device pointers into the staging block)
  int g0, glen;
  const double* trial; const double* gq; const int* gtype; const uint8_t* moved;
  // previous trial whose commit is still pending
  int prev_valid;
  int prev_accept;      // 0/1: decided by the host; -1: decided on the device (state->accept)
  int pg0, pglen;
  const double* ptrial;
  // reciprocal space
  const double4* kvec;  // [nk] (kx, ky, kz, ek2), half space, built on the host with the reference's expressions
  double2* S; double2* dS; int nk;
  // tiling
  int n_tiles, n_chunks, chunk_size, n_pair_ctas, n_k_ctas, n_intra_ctas;
  double* partial;      // [n_ctas][8]
  PgState* state;
  PgMailRec* mail;      // mapped host memory (NULL in replay)
  int decide_on_device; double u; double* replay_dE; uint8_t* replay_acc; int replay_index;
  unsigned int seq;
  unsigned long long* timing;   // debug only (PLUM_B200_TIMING=1): per-CTA %globaltimer stamps [n_ctas][8]
  const PgDev* Pg;      // copy of the parameter block in global memory, for the non-inlined rare paths
                        // (passing the by-value kernel parameter by reference would spill all of it to local memory)
};

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

__device__ __forceinline__ void mv_mail(PgMailRec* m, int slot, double value, unsigned int seq, int aux) {
  // one 16-byte store
  asm volatile("st.volatile.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(&m[slot]), "r"(__double2loint(value)),
               "r"(__double2hiint(value)), "r"((int)seq), "r"(aux)
               : "memory");
}

// The rare in-range pair: exact LJ / real-space arithmetic for one configuration whose squared
// separation r2 passed one of the relaxed filters.  Single central image, no fold needed.
__device__ __noinline__ void mv_pair_inrange(const PgDev* __restrict__ Pg, double r2, double qq, int tp,
                                             double& e_lj, double& e_real) {
  const PgDev& P = *Pg;
  e_lj = 0.0;
  e_real = 0.0;
  const double r = sqrt(r2);
  if (P.pair_kind == 1 && r2 <= P.lj_rcut2_relaxed[tp]) e_lj = pg_pair_energy_r(P, r, tp);
  if (qq != 0 && r2 <= P.rc2_relaxed) {
    if (r > 0 && r <= P.real_cutoff) e_real = P.lB * qq * erfc(P.sqrt_alpha * r) / r;
  }
}

__device__ __noinline__ void mv_pair_generic(const PgDev* __restrict__ Pg, double ax, double ay, double az, double qa,
                                             int ta, double bx, double by, double bz, double qb, int tb, int do_lj,
                                             double& e_lj, double& e_real) {
  pg_pair_both(*Pg, ax, ay, az, qa, ta, bx, by, bz, qb, tb, do_lj, e_lj, e_real);
}

// Sum 5 values over the CTA with one barrier; result valid in thread 0.
__device__ __forceinline__ void mv_block_sum5(double (&v)[5], double* smem /* [MV_WARPS][5] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 5; i++) v[i] = warp_sum(v[i]);
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
__global__ void __launch_bounds__(MV_THREADS, 3) k_move(const PgDev P, const PgMoveArgs A) {
  // pair CTAs stage one chunk of <= MV_GCHUNK group beads; k CTAs stage up to MV_THREADS beads per round
  __shared__ double s_n[3][MV_THREADS], s_o[3][MV_THREADS], s_q[MV_THREADS];
  __shared__ int s_t[MV_GCHUNK], s_mv[MV_GCHUNK];
  __shared__ double s_ljc2[PG_MAX_TYPES * PG_MAX_TYPES];
  __shared__ double s_red[MV_WARPS * 5];
  __shared__ int s_last;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  MV_STAMP(0);
  // decision on the previous trial (uniform over the grid)
  int apply_prev = 0;
  if (A.prev_valid) apply_prev = (A.prev_accept >= 0) ? A.prev_accept : __ldcg(&A.state->accept);
  const int pg0 = A.pg0, pg1 = apply_prev ? A.pg0 + A.pglen : A.pg0;   // empty range when not applied

  double acc_pair = 0.0, acc_real = 0.0, acc_mz = 0.0, acc_rec = 0.0, acc_ov = 0.0;

  if (b < A.n_pair_ctas) {
    const int tile = b % A.n_tiles, chunk = b / A.n_tiles;
    const int gbeg = chunk * A.chunk_size;
    const int gcnt = min(A.glen, gbeg + A.chunk_size) - gbeg;
    // partner first: its loads are the longest dependency chain of the CTA
    const int j = tile * MV_THREADS + tid;
    double px = 0, py = 0, pz = 0, pq = 0;
    int pt = 0;
    if (j < A.n) {
      const double2 c = A.zq[j];
      const double2 a = A.xy[j];
      pt = A.type[j];
      pq = c.y;
      px = a.x; py = a.y; pz = c.x;
      if (j >= pg0 && j < pg1) {
        px = A.ptrial[3 * (j - pg0)]; py = A.ptrial[3 * (j - pg0) + 1]; pz = A.ptrial[3 * (j - pg0) + 2];
      }
      if (chunk == 0) acc_mz = pq * pz;
    }
    if (tid < gcnt) {
      const int g = gbeg + tid, jg = A.g0 + g;
      double ox, oy, oz;
      if (jg >= pg0 && jg < pg1) {
        ox = A.ptrial[3 * (jg - pg0)]; oy = A.ptrial[3 * (jg - pg0) + 1]; oz = A.ptrial[3 * (jg - pg0) + 2];
      } else {
        double2 a = A.xy[jg], c = A.zq[jg];
        ox = a.x; oy = a.y; oz = c.x;
      }
      s_o[0][tid] = ox; s_o[1][tid] = oy; s_o[2][tid] = oz;
      s_n[0][tid] = A.trial[3 * g]; s_n[1][tid] = A.trial[3 * g + 1]; s_n[2][tid] = A.trial[3 * g + 2];
      s_q[tid] = A.gq[g];
      s_t[tid] = A.gtype[g];
      s_mv[tid] = A.moved[g];
    } else if (FAST && tid >= 64 && tid < 64 + PG_MAX_TYPES * PG_MAX_TYPES) {
      const int e = tid - 64;
      s_ljc2[e] = (P.pair_kind == 1) ? A.Pg->lj_rcut2_relaxed[e] : -1.0;   // coalesced global read, not 64 LDCs
    }
    __syncthreads();
    MV_STAMP(1);
    if (j < A.n) {
      const bool in_group = (j >= A.g0) && (j < A.g0 + A.glen);
      const int do_lj = (P.pair_kind != 0);
      if (!in_group) {
        if (FAST) {
          const double Lx = P.box[0], Ly = P.box[1], Lz = P.box[2];
          const double iLx = P.inv_box[0], iLy = P.inv_box[1], iLz = P.inv_box[2];
          const double rc2 = P.use_ewald ? P.rc2_relaxed : -1.0;
#pragma unroll 2
          for (int i = 0; i < gcnt; i++) {
            if (!s_mv[i]) continue;
            const double gq = s_q[i];
            const int tp = s_t[i] * PG_MAX_TYPES + pt;
            double dxn = px - s_n[0][i], dyn = py - s_n[1][i], dzn = pz - s_n[2][i];
            double dxo = px - s_o[0][i], dyo = py - s_o[1][i], dzo = pz - s_o[2][i];
            dxn -= Lx * mv_rint(dxn * iLx); dyn -= Ly * mv_rint(dyn * iLy); dzn -= Lz * mv_rint(dzn * iLz);
            dxo -= Lx * mv_rint(dxo * iLx); dyo -= Ly * mv_rint(dyo * iLy); dzo -= Lz * mv_rint(dzo * iLz);
            const double r2n = dxn * dxn + dyn * dyn + dzn * dzn;
            const double r2o = dxo * dxo + dyo * dyo + dzo * dzo;
            const double ljc2 = s_ljc2[tp];
            const double qq = gq * pq;
            const double cut = (qq != 0.0) ? fmax(rc2, ljc2) : ljc2;
            if (fmin(r2n, r2o) <= cut) {
              double lj_n = 0.0, re_n = 0.0, lj_o = 0.0, re_o = 0.0;
              if (r2n <= cut) mv_pair_inrange(A.Pg, r2n, qq, tp, lj_n, re_n);
              if (r2o <= cut) mv_pair_inrange(A.Pg, r2o, qq, tp, lj_o, re_o);
              if (lj_n >= PG_VLE) acc_ov += 1.0;
              acc_pair += (lj_n - lj_o);
              acc_real += (re_n - re_o);
            }
          }
        } else {
          for (int i = 0; i < gcnt; i++) {
            if (!s_mv[i]) continue;
            double lj_n, re_n, lj_o, re_o;
            mv_pair_generic(A.Pg, s_n[0][i], s_n[1][i], s_n[2][i], s_q[i], s_t[i], px, py, pz, pq, pt, do_lj, lj_n, re_n);
            mv_pair_generic(A.Pg, s_o[0][i], s_o[1][i], s_o[2][i], s_q[i], s_t[i], px, py, pz, pq, pt, do_lj, lj_o, re_o);
            if (lj_n >= PG_VLE) acc_ov += 1.0;
            acc_pair += (lj_n - lj_o);
            acc_real += (re_n - re_o);
          }
        }
      }
      // partners that are beads of the moved molecule itself are handled by the intra CTAs below
    }
  } else if (b < A.n_pair_ctas + A.n_k_ctas) {
    // reciprocal part: one k per thread.  The previous trial's dS is folded into S first.
    const int k = (b - A.n_pair_ctas) * MV_THREADS + tid;
    double4 kv = make_double4(0.0, 0.0, 0.0, 0.0);
    double2 S = make_double2(0.0, 0.0);
    if (k < A.nk) {
      kv = A.kvec[k];
      S = A.S[k];
      if (apply_prev) {
        const double2 d = A.dS[k];
        S.x += d.x; S.y += d.y;
        A.S[k] = S;
      }
    }
    double dre = 0.0, dim = 0.0;
    for (int gbeg = 0; gbeg < A.glen; gbeg += MV_THREADS) {
      const int gcnt = min(A.glen - gbeg, MV_THREADS);
      if (gbeg > 0) __syncthreads();
      if (tid < gcnt) {
        const int g = gbeg + tid, jg = A.g0 + g;
        const double q = A.moved[g] ? A.gq[g] : 0.0;   // unmoved or neutral beads drop out of dS
        s_q[tid] = q;
        if (q != 0.0) {
          double ox, oy, oz;
          if (jg >= pg0 && jg < pg1) {
            ox = A.ptrial[3 * (jg - pg0)]; oy = A.ptrial[3 * (jg - pg0) + 1]; oz = A.ptrial[3 * (jg - pg0) + 2];
          } else {
            double2 a = A.xy[jg], c = A.zq[jg];
            ox = a.x; oy = a.y; oz = c.x;
          }
          s_o[0][tid] = pg_wrap_pos(ox, P.ebox[0], P.inv_ebox[0], P.pbc[0]);
          s_o[1][tid] = pg_wrap_pos(oy, P.ebox[1], P.inv_ebox[1], P.pbc[1]);
          s_o[2][tid] = pg_wrap_pos(oz, P.ebox[2], P.inv_ebox[2], P.pbc[2]);
          s_n[0][tid] = pg_wrap_pos(A.trial[3 * g], P.ebox[0], P.inv_ebox[0], P.pbc[0]);
          s_n[1][tid] = pg_wrap_pos(A.trial[3 * g + 1], P.ebox[1], P.inv_ebox[1], P.pbc[1]);
          s_n[2][tid] = pg_wrap_pos(A.trial[3 * g + 2], P.ebox[2], P.inv_ebox[2], P.pbc[2]);
        }
      }
      __syncthreads();
      if (k < A.nk) {
        for (int i = 0; i < gcnt; i++) {
          const double q = s_q[i];
          if (q == 0) continue;
          double sn, cn, so, co;
          sincos(kv.x * s_n[0][i] + kv.y * s_n[1][i] + kv.z * s_n[2][i], &sn, &cn);
          sincos(kv.x * s_o[0][i] + kv.y * s_o[1][i] + kv.z * s_o[2][i], &so, &co);
          dre += q * cn; dim += q * sn;
          dre -= q * co; dim -= q * so;
        }
      }
    }
    if (k < A.nk) {
      // never |S_new|^2 - |S_old|^2: 2 Re(conj(S) dS) + |dS|^2; x2 for the -k half
      acc_rec = 2.0 * kv.w * (2.0 * (S.x * dre + S.y * dim) + (dre * dre + dim * dim));
      A.dS[k] = make_double2(dre, dim);
    }
  }

  else if (b < A.n_pair_ctas + A.n_k_ctas + A.n_intra_ctas) {
    // intra-molecular pairs (g, jj), g < jj, of the moved molecule, one pair per thread
    // (potential_pair.cc:157-178, potential_ewald.cc:436-477; the g == jj self-image term is
    // identical before and after a move).  Sized by the host: glen*(glen-1)/2 threads.
    const int p = (b - A.n_pair_ctas - A.n_k_ctas) * MV_THREADS + tid;
    const int npairs = A.glen * (A.glen - 1) / 2;
    if (p < npairs) {
      int jj = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
      while (jj * (jj - 1) / 2 > p) jj--;
      while ((jj + 1) * jj / 2 <= p) jj++;
      const int g = p - jj * (jj - 1) / 2;
      if (A.moved[g] || A.moved[jj]) {
        const int ja = A.g0 + g, jb = A.g0 + jj;
        double aox, aoy, aoz, box_, boy, boz;
        if (ja >= pg0 && ja < pg1) {
          aox = A.ptrial[3 * (ja - pg0)]; aoy = A.ptrial[3 * (ja - pg0) + 1]; aoz = A.ptrial[3 * (ja - pg0) + 2];
        } else {
          double2 a = A.xy[ja], c = A.zq[ja];
          aox = a.x; aoy = a.y; aoz = c.x;
        }
        if (jb >= pg0 && jb < pg1) {
          box_ = A.ptrial[3 * (jb - pg0)]; boy = A.ptrial[3 * (jb - pg0) + 1]; boz = A.ptrial[3 * (jb - pg0) + 2];
        } else {
          double2 a = A.xy[jb], c = A.zq[jb];
          box_ = a.x; boy = a.y; boz = c.x;
        }
        const double anx = A.trial[3 * g], any_ = A.trial[3 * g + 1], anz = A.trial[3 * g + 2];
        const double bnx = A.trial[3 * jj], bny = A.trial[3 * jj + 1], bnz = A.trial[3 * jj + 2];
        const double qa = A.gq[g], qb = A.gq[jj];
        const int ta = A.gtype[g], tb = A.gtype[jj];
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
          mv_pair_inrange(A.Pg, r2n, qq, tp, lj_n, re_n);
          mv_pair_inrange(A.Pg, r2o, qq, tp, lj_o, re_o);
        } else {
          int lj_here = (P.pair_kind != 0);
          if (P.pair_kind == 2 && jj == g + 1) lj_here = 0;   // bonded hard spheres, potential_pair.cc:165-169
          mv_pair_generic(A.Pg, anx, any_, anz, qa, ta, bnx, bny, bnz, qb, tb, lj_here, lj_n, re_n);
          mv_pair_generic(A.Pg, aox, aoy, aoz, qa, ta, box_, boy, boz, qb, tb, lj_here, lj_o, re_o);
        }
        if (lj_n >= PG_VLE) acc_ov += 1.0;
        acc_pair += (lj_n - lj_o);
        acc_real += (re_n - re_o);
      }
    }
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
      if (A.timing) A.timing[8 * b + 3] = mv_now();
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
        ox = A.ptrial[3 * (jg - pg0)]; oy = A.ptrial[3 * (jg - pg0) + 1]; oz = A.ptrial[3 * (jg - pg0) + 2];
      } else {
        double2 a = A.xy[jg], c = A.zq[jg];
        ox = a.x; oy = a.y; oz = c.x;
      }
      if (P.ext_kind != 0 && A.moved[g]) {
        const int t = A.gtype[g];
        const double en = pg_wall_energy(P, A.trial[3 * g + 2], t);
        if (en >= PG_VLE) w_out = 1.0;
        w_sum += en - pg_wall_energy(P, oz, t);
      }
      if (P.bond_kind != 0 && g + 1 < A.glen) {
        const int jg2 = jg + 1;
        double ox2, oy2, oz2;
        if (jg2 >= pg0 && jg2 < pg1) {
          ox2 = A.ptrial[3 * (jg2 - pg0)]; oy2 = A.ptrial[3 * (jg2 - pg0) + 1]; oz2 = A.ptrial[3 * (jg2 - pg0) + 2];
        } else {
          double2 a = A.xy[jg2], c = A.zq[jg2];
          ox2 = a.x; oy2 = a.y; oz2 = c.x;
        }
        b_sum += pg_bond_energy(P, A.trial[3 * g], A.trial[3 * g + 1], A.trial[3 * g + 2], A.trial[3 * g + 3],
                                A.trial[3 * g + 4], A.trial[3 * g + 5]) -
                 pg_bond_energy(P, ox, oy, oz, ox2, oy2, oz2);
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
    st->dE = dE; st->d_pair = d_pair; st->d_ext = d_ext; st->d_ewald = d_ewald; st->d_bond = d_bond;
    st->d_real = d_real; st->d_recip = d_recip; st->d_self = 0.0; st->d_dipole = d_dip;
    st->mz_current = mz_cur;
    st->stage = stage; st->n_overlap = n_overlap; st->accept = accept; st->mode = PG_MODE_MOVE;
    st->done_counter = 0;
  }
  // write the previous trial's coordinates back: every other CTA has finished reading them
  if (apply_prev) {
    for (int i = tid; i < A.pglen; i += MV_THREADS) {
      A.xy[A.pg0 + i] = make_double2(A.ptrial[3 * i], A.ptrial[3 * i + 1]);
      double2 c = A.zq[A.pg0 + i];
      c.x = A.ptrial[3 * i + 2];
      A.zq[A.pg0 + i] = c;
    }
  }
  MV_STAMP(7);
}
