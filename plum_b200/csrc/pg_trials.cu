// k_trials — the energies of one configurational-bias growth step, one launch.
//
// Replaces ForceField::BeadsEnergy (reference src/force_field/cbmc.cc:5-151), which
// CBMCFGenTrialBeads (cbmc.cc:161-212) calls once per trial position: here all n_trials trial
// positions (monomer bead1 + optional counter-ion bead2) are evaluated together against every
// resident bead except the skipped molecules [skip_b0, skip_b1), against the partial chain grown
// so far, against each other, the walls, and reciprocal space.
//
// Layout (same ideas as k_move, pg_move.cu):
//  * helper CTAs [0, n_helpers): a slice of the k vectors each, LPK lanes per k.  The lanes first
//    build S'(k) = S(k) - sum_{skipped} q e^{ik.r} + sum_{chain} q e^{ik.r} (the structure factor of
//    the partners this growth step sees) with a shuffle reduction, then share the trials:
//    e_t(k) = 2 ek2 (2 Re(conj(S') d_t) + |d_t|^2).
//  * pair CTAs: the work is n_tiles x n_trials units in tile-major order (tile = 256 partner slots;
//    a slot is one partner, or one (partner, periodic-image column) when the real-space sum runs
//    over several images).  A CTA loads its partner slot once per tile and loops over its trials.
//  * every (tile, trial) and (helper, trial) partial sum has one writer; the last CTA to finish adds
//    them in index order (deterministic), applies the per-trial scalar terms and posts three
//    self-validating 16-byte records per trial into mapped host memory.  No H2D/D2H copies for
//    growth steps of up to TR_INL_T trials and TR_INL_C chain beads (they ride in the parameters).
#pragma once
#include "pg_kernels.cuh"

#define TR_THREADS 256
#define TR_MAXT 64        // trials per launch
#define TR_INL_T 32       // most trials / chain beads whose coordinates travel in the kernel parameters
#define TR_INL_C 16
#define TR_MIN_LPK 8      // lanes per k vector: 8..32

struct PgTrialArgs {
  PgMoveDev D;
  const double2* xy; const double2* zq; const int* type; int n;
  int skip_b0, skip_b1;
  int nt, nc, current_len, use_b2, t1, t2;   // nc = chain beads: current_len monomers, then (use_b2) as many ions
  double q1, q2;
  // staged inputs (device) ...
  const double* b1; const double* b2; const double* cxyz; const double* cq; const int* ctype;
  // ... or inline ones
  int inl;
  double inl_b1[3 * TR_INL_T], inl_b2[3 * TR_INL_T], inl_cxyz[3 * TR_INL_C], inl_cq[TR_INL_C];
  int inl_ct[TR_INL_C];
  const double4* kvec; const double2* S; int nk; int lpk;
  int n_tiles, n_helpers, n_pair_ctas;
  double2* partial;      // [n_helpers + n_tiles][nt]: helpers (rec, -), tiles (lj, real)
  double* mz_partial;    // [n_tiles]
  unsigned int* counter;
  PgMailRec* mail;       // [3 * TR_MAXT]: record 3t energy, 3t+1 pair, 3t+2 ewald
  unsigned int seq;
  const PgDev* Pg;
};

__global__ void __launch_bounds__(TR_THREADS, 2) k_trials(const PgTrialArgs A) {
  const PgMoveDev& P = A.D;
  __shared__ double s_b1[3 * TR_MAXT], s_b2[3 * TR_MAXT];
  __shared__ double s_cx[3 * TR_INL_C], s_cq[TR_INL_C];
  __shared__ int s_ct[TR_INL_C];
  __shared__ double s_acc[(TR_THREADS / TR_MIN_LPK) * TR_MAXT];   // helpers: [k group][trial]; pair: [warp][trial][2]
  __shared__ double s_mz[TR_THREADS / 32];
  __shared__ int s_last;

  const int b = blockIdx.x, tid = threadIdx.x;
  const int nt = A.nt;
  const double* __restrict__ cxyz = A.cxyz;
  const double* __restrict__ cq = A.cq;
  const int* __restrict__ ctype = A.ctype;
  if (A.inl) {
    if (tid < 3 * nt) { s_b1[tid] = A.inl_b1[tid]; s_b2[tid] = A.inl_b2[tid]; }
    if (tid < 3 * A.nc) s_cx[tid] = A.inl_cxyz[tid];
    if (tid < A.nc) { s_cq[tid] = A.inl_cq[tid]; s_ct[tid] = A.inl_ct[tid]; }
    cxyz = s_cx; cq = s_cq; ctype = s_ct;
  } else {
    if (tid < 3 * nt) { s_b1[tid] = A.b1[tid]; s_b2[tid] = A.use_b2 ? A.b2[tid] : 0.0; }
  }
  __syncthreads();

  if (b < A.n_helpers) {
    // ---------------- reciprocal space
    const int LPK = A.lpk, KPP = TR_THREADS / LPK;
    const int kl = tid / LPK, e0 = tid % LPK;
    for (int i = tid; i < KPP * nt; i += TR_THREADS) s_acc[i] = 0.0;
    __syncthreads();
    unsigned k0u, k1u;
    mv_share((unsigned)A.nk, (unsigned)A.n_helpers, (unsigned)b, k0u, k1u);
    const int k0 = (int)k0u, k1 = (int)k1u;
    const int n_skip = A.skip_b1 - A.skip_b0;
    for (int kb = k0; kb < k1; kb += KPP) {
      const int k = kb + kl;
      const bool active = k < k1;
      double4 kv = make_double4(0.0, 0.0, 0.0, 0.0);
      double sre = 0.0, sim = 0.0;
      if (active) {
        kv = A.kvec[k];
        for (int e = e0; e < n_skip + A.nc; e += LPK) {
          double x, y, z, q;
          if (e < n_skip) {
            const double2 a = A.xy[A.skip_b0 + e], c = A.zq[A.skip_b0 + e];
            x = a.x; y = a.y; z = c.x; q = -c.y;
          } else {
            const int c = e - n_skip;
            x = cxyz[3 * c]; y = cxyz[3 * c + 1]; z = cxyz[3 * c + 2]; q = cq[c];
          }
          if (q == 0.0) continue;
          x = pg_wrap_pos(x, P.ebox[0], P.inv_ebox[0], P.pbc[0]);
          y = pg_wrap_pos(y, P.ebox[1], P.inv_ebox[1], P.pbc[1]);
          z = pg_wrap_pos(z, P.ebox[2], P.inv_ebox[2], P.pbc[2]);
          double sn, cs;
          sincos(kv.x * x + kv.y * y + kv.z * z, &sn, &cs);
          sre += q * cs; sim += q * sn;
        }
      }
      for (int o = LPK >> 1; o > 0; o >>= 1) {
        sre += __shfl_xor_sync(0xffffffffu, sre, o);
        sim += __shfl_xor_sync(0xffffffffu, sim, o);
      }
      if (active) {
        const double2 S = A.S[k];
        sre += S.x; sim += S.y;   // S'(k)
        for (int t = e0; t < nt; t += LPK) {
          double sn, cs;
          sincos(kv.x * pg_wrap_pos(s_b1[3 * t], P.ebox[0], P.inv_ebox[0], P.pbc[0]) +
                 kv.y * pg_wrap_pos(s_b1[3 * t + 1], P.ebox[1], P.inv_ebox[1], P.pbc[1]) +
                 kv.z * pg_wrap_pos(s_b1[3 * t + 2], P.ebox[2], P.inv_ebox[2], P.pbc[2]), &sn, &cs);
          double dre = A.q1 * cs, dim = A.q1 * sn;
          if (A.use_b2) {
            sincos(kv.x * pg_wrap_pos(s_b2[3 * t], P.ebox[0], P.inv_ebox[0], P.pbc[0]) +
                   kv.y * pg_wrap_pos(s_b2[3 * t + 1], P.ebox[1], P.inv_ebox[1], P.pbc[1]) +
                   kv.z * pg_wrap_pos(s_b2[3 * t + 2], P.ebox[2], P.inv_ebox[2], P.pbc[2]), &sn, &cs);
            dre += A.q2 * cs; dim += A.q2 * sn;
          }
          s_acc[kl * nt + t] += 2.0 * kv.w * (2.0 * (sre * dre + sim * dim) + (dre * dre + dim * dim));
        }
      }
    }
    __syncthreads();
    if (tid < nt) {
      double r = 0.0;
      for (int g = 0; g < KPP; g++) r += s_acc[g * nt + tid];
      A.partial[b * nt + tid] = make_double2(r, 0.0);
    }
  } else {
    // ---------------- trials x partner slots
    unsigned u, u1;
    mv_share((unsigned)(A.n_tiles * nt), (unsigned)A.n_pair_ctas, (unsigned)(b - A.n_helpers), u, u1);
    const int split = P.img_split;
    const int n_all = A.n + A.nc;
    const bool do_lj = (P.pair_kind != 0);
    const int warp = tid >> 5, lane = tid & 31;
    while (u < u1) {
      const int tile = (int)(u / (unsigned)nt);
      const int ta = (int)u - tile * nt;
      const int tb = min(nt, ta + (int)(u1 - u));
      const int slot = tile * TR_THREADS + tid;
      const int j = slot / split, sub = slot - j * split;
      bool valid = j < n_all && !(j >= A.skip_b0 && j < A.skip_b1);
      double px = 0, py = 0, pz = 0, pq = 0;
      int pt = 0;
      bool bonded = false;   // the last grown monomer: no LJ with bead1 (cbmc.cc:57-58)
      if (valid) {
        if (j < A.n) {
          const double2 a = A.xy[j], c = A.zq[j];
          px = a.x; py = a.y; pz = c.x; pq = c.y; pt = A.type[j];
        } else {
          const int c = j - A.n;
          px = cxyz[3 * c]; py = cxyz[3 * c + 1]; pz = cxyz[3 * c + 2]; pq = cq[c]; pt = ctype[c];
          bonded = (c == A.current_len - 1);
        }
      }
      const int tp1 = A.t1 * PG_MAX_TYPES + pt, tp2 = A.t2 * PG_MAX_TYPES + pt;
      const double qq1 = A.q1 * pq, qq2 = A.q2 * pq;
      for (int t = ta; t < tb; t++) {
        double e_lj = 0.0, e_re = 0.0;
        if (valid) {
          const double2 e1 = mv_generic_column(P, A.Pg, s_b1[3 * t], s_b1[3 * t + 1], s_b1[3 * t + 2], px, py, pz, qq1, tp1,
                                               sub, do_lj && !bonded);
          e_lj = e1.x; e_re = e1.y;
          if (A.use_b2) {
            const double2 e2 = mv_generic_column(P, A.Pg, s_b2[3 * t], s_b2[3 * t + 1], s_b2[3 * t + 2], px, py, pz, qq2,
                                                 tp2, sub, do_lj);
            e_lj += e2.x; e_re += e2.y;
          }
        }
        for (int o = 16; o > 0; o >>= 1) {
          e_lj += __shfl_xor_sync(0xffffffffu, e_lj, o);
          e_re += __shfl_xor_sync(0xffffffffu, e_re, o);
        }
        if (lane == 0) { s_acc[(warp * TR_MAXT + (t - ta)) * 2] = e_lj; s_acc[(warp * TR_MAXT + (t - ta)) * 2 + 1] = e_re; }
      }
      // the dipole moment of the partners is the same for every trial: trial 0's CTA adds it up
      if (ta == 0) {
        double mz = (valid && sub == 0) ? pq * pz : 0.0;
        for (int o = 16; o > 0; o >>= 1) mz += __shfl_xor_sync(0xffffffffu, mz, o);
        if (lane == 0) s_mz[warp] = mz;
      }
      __syncthreads();
      if (tid < tb - ta) {
        double lj = 0.0, re = 0.0;
        for (int w = 0; w < TR_THREADS / 32; w++) { lj += s_acc[(w * TR_MAXT + tid) * 2]; re += s_acc[(w * TR_MAXT + tid) * 2 + 1]; }
        A.partial[(A.n_helpers + tile) * nt + ta + tid] = make_double2(lj, re);
      }
      if (ta == 0 && tid == 0) {
        double mz = 0.0;
        for (int w = 0; w < TR_THREADS / 32; w++) mz += s_mz[w];
        A.mz_partial[tile] = mz;
      }
      __syncthreads();
      u += (unsigned)(tb - ta);
    }
  }

  // ---------------- last CTA: per-trial totals and the scalar terms
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(A.counter) : "memory");
    s_last = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) *A.counter = 0;
  // L lanes per trial (power of two, 4..32): the lanes share the periodic-image columns of the
  // bead1-bead2 pair, lane 0 then owns the trial
  int L = 32;
  while (L * nt > TR_THREADS) L >>= 1;
  const int t = tid / L, c = tid - t * L;
  const bool live = t < nt;
  const double x1 = live ? s_b1[3 * t] : 0.0, y1 = live ? s_b1[3 * t + 1] : 0.0, z1 = live ? s_b1[3 * t + 2] : 0.0;
  const double x2 = live ? s_b2[3 * t] : 0.0, y2 = live ? s_b2[3 * t + 1] : 0.0, z2 = live ? s_b2[3 * t + 2] : 0.0;
  double lj12 = 0.0, re12 = 0.0;
  if (live && A.use_b2) {
    for (int sub = c; sub < P.img_split; sub += L) {
      const double2 e = mv_generic_column(P, A.Pg, x1, y1, z1, x2, y2, z2, A.q1 * A.q2, A.t1 * PG_MAX_TYPES + A.t2, sub,
                                          P.pair_kind != 0);
      lj12 += e.x; re12 += e.y;
    }
  }
  for (int o = L >> 1; o > 0; o >>= 1) {
    lj12 += __shfl_xor_sync(0xffffffffu, lj12, o);
    re12 += __shfl_xor_sync(0xffffffffu, re12, o);
  }
  if (!live || c != 0) return;
  const PgDev& G = *A.Pg;
  double rec = 0.0, lj = 0.0, re = 0.0, mz_o = 0.0;
  for (int h = 0; h < A.n_helpers; h++) rec += __ldcg(&A.partial[h * nt + t]).x;
  for (int i = 0; i < A.n_tiles; i++) {
    const double2 v = __ldcg(&A.partial[(A.n_helpers + i) * nt + t]);
    lj += v.x; re += v.y;
  }
  if (P.dipole)
    for (int i = 0; i < A.n_tiles; i++) mz_o += __ldcg(&A.mz_partial[i]);
  double pair_e = lj + lj12, ewald_e = 0.0;
  if (P.use_ewald && pair_e < PG_VLE) {
    // self terms: -sqrt(alpha/pi) lB q^2 and half the bead's own periodic images (position independent)
    ewald_e = re + P.recip_pref * rec;
    ewald_e += (G.self_pref + G.real_self_unit) * A.q1 * A.q1;
    if (A.use_b2) ewald_e += (G.self_pref + G.real_self_unit) * A.q2 * A.q2 + re12;
    if (P.dipole) {
      double mz_n = mz_o + A.q1 * z1;
      if (A.use_b2) mz_n += A.q2 * z2;
      ewald_e += P.dipole_pref * (mz_n * mz_n - mz_o * mz_o);
    }
  }
  if (P.ext_kind != 0) {
    pair_e += pg_wall_energy(G, z1, A.t1);
    if (A.use_b2) pair_e += pg_wall_energy(G, z2, A.t2);
  }
  mv_mail(A.mail, 3 * t, (pair_e >= PG_VLE) ? PG_VLE : (pair_e + ewald_e), A.seq, t);
  mv_mail(A.mail, 3 * t + 1, pair_e, A.seq, t);
  mv_mail(A.mail, 3 * t + 2, ewald_e, A.seq, t);
}
