// plum_b200 — sm_100a kernels of the per-trial-move energy path.  See
// pg_kernels.cuh for the kernel list and DESIGN.md for the data layout and the
// roofline of each kernel.  No tensor cores: this is FP64 pair arithmetic with
// branches, not a contraction.
#include "pg_kernels.cuh"

// ----------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum NV values across the CTA (fixed order: deterministic). Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* [NV * 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) smem[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double x = (lane < nwarp) ? smem[i * 32 + lane] : 0.0;
      v[i] = warp_sum(x);
    }
  }
}

// ------------------------------------------------------------------ k_delta
// Grid: n_pair_ctas CTAs over (partner tile, group chunk) + n_k_ctas CTAs over k
// tiles (+1 idle CTA if both are zero).  Block: PG_TILE threads.
__global__ void __launch_bounds__(PG_TILE) k_delta(const PgDev P, const PgDeltaArgs A) {
  __shared__ double s_o[3][PG_GCHUNK], s_n[3][PG_GCHUNK], s_q[PG_GCHUNK];
  __shared__ int s_t[PG_GCHUNK], s_mv[PG_GCHUNK];
  __shared__ double s_red[8 * 32];
  __shared__ int s_redi[32];
  __shared__ int s_last;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  double acc_pair = 0.0, acc_real = 0.0, acc_mz = 0.0, acc_rec = 0.0;
  int acc_ov = 0;

  if (b < A.n_pair_ctas) {
    // ---------------- short-range part: one partner per thread x a chunk of the group
    const int tile = b % A.n_tiles, chunk = b / A.n_tiles;
    const int gbeg = chunk * A.chunk_size;
    const int gcnt = min(A.glen, gbeg + A.chunk_size) - gbeg;
    if (tid < gcnt) {
      const int g = gbeg + tid;
      if (A.has_old) {
        double2 a = A.xy[A.g0 + g], c = A.zq[A.g0 + g];
        s_o[0][tid] = a.x; s_o[1][tid] = a.y; s_o[2][tid] = c.x;
      } else {
        s_o[0][tid] = s_o[1][tid] = s_o[2][tid] = 0.0;
      }
      if (A.has_new) {
        s_n[0][tid] = A.trial[3 * g]; s_n[1][tid] = A.trial[3 * g + 1]; s_n[2][tid] = A.trial[3 * g + 2];
      } else {
        s_n[0][tid] = s_n[1][tid] = s_n[2][tid] = 0.0;
      }
      s_q[tid] = A.gq[g];
      s_t[tid] = A.gtype[g];
      s_mv[tid] = A.moved[g];
    }
    __syncthreads();
    const int j = tile * PG_TILE + tid;
    if (j < A.n_partners) {
      const double2 pxy = A.xy[j], pzq = A.zq[j];
      const int pt = A.type[j];
      const double px = pxy.x, py = pxy.y, pz = pzq.x, pq = pzq.y;
      if (chunk == 0) acc_mz = pq * pz;
      const bool in_group = (A.mode != PG_MODE_INSERT) && (j >= A.g0) && (j < A.g0 + A.glen);
      const int do_lj = (P.pair_kind != 0) && !(P.pair_kind == 2 && A.mode != PG_MODE_MOVE);
      if (!in_group) {
        for (int i = 0; i < gcnt; i++) {
          if (!s_mv[i]) continue;
          double lj_n = 0.0, re_n = 0.0, lj_o = 0.0, re_o = 0.0;
          if (A.has_new)
            pg_pair_both(P, s_n[0][i], s_n[1][i], s_n[2][i], s_q[i], s_t[i], px, py, pz, pq, pt, do_lj, lj_n, re_n);
          if (A.has_old)
            pg_pair_both(P, s_o[0][i], s_o[1][i], s_o[2][i], s_q[i], s_t[i], px, py, pz, pq, pt, do_lj, lj_o, re_o);
          if (lj_n >= PG_VLE) acc_ov++;
          acc_pair += (lj_n - lj_o);
          acc_real += (re_n - re_o);
        }
      } else {
        // partner is itself a group bead: intra-group pairs (g, jj) with g <= jj, once
        const int jj = j - A.g0;
        const int jmv = A.moved[jj];
        double nx = 0.0, ny = 0.0, nz = 0.0;
        if (A.has_new) { nx = A.trial[3 * jj]; ny = A.trial[3 * jj + 1]; nz = A.trial[3 * jj + 2]; }
        for (int i = 0; i < gcnt; i++) {
          const int g = gbeg + i;
          if (g > jj) break;
          if (!(s_mv[i] || jmv)) continue;
          if (g < jj) {
            int lj_here = do_lj;
            // bonded neighbours of a hard-sphere chain are forced to 0 (potential_pair.cc:165-169)
            if (P.pair_kind == 2 && jj == g + 1 && jj < A.chain_len_first) lj_here = 0;
            double lj_n = 0.0, re_n = 0.0, lj_o = 0.0, re_o = 0.0;
            if (A.has_new)
              pg_pair_both(P, s_n[0][i], s_n[1][i], s_n[2][i], s_q[i], s_t[i], nx, ny, nz, pq, pt, lj_here, lj_n, re_n);
            if (A.has_old)
              pg_pair_both(P, s_o[0][i], s_o[1][i], s_o[2][i], s_q[i], s_t[i], px, py, pz, pq, pt, lj_here, lj_o, re_o);
            if (lj_n >= PG_VLE) acc_ov++;
            acc_pair += (lj_n - lj_o);
            acc_real += (re_n - re_o);
          } else if (P.use_ewald) {
            // g == jj: the bead with its own periodic images, x0.5 (potential_ewald.cc:445-448)
            double qq = pq * pq;
            if (qq != 0) {
              double e = P.real_self_unit * qq;   // position independent: summed once on the host
              acc_real += (A.has_new ? e : 0.0) - (A.has_old ? e : 0.0);
            }
          }
        }
      }
    }
    if (A.mode == PG_MODE_INSERT && chunk == 0 && tile == 0) {
      // intra-group pairs of beads that are not resident yet: done by one CTA, strided
      const int do_lj = (P.pair_kind == 1);
      for (int p = tid; p < A.glen * A.glen; p += PG_TILE) {
        const int g = p / A.glen, jj = p % A.glen;
        if (g > jj) continue;
        const double ax = A.trial[3 * g], ay = A.trial[3 * g + 1], az = A.trial[3 * g + 2];
        if (g < jj) {
          double lj_n, re_n;
          pg_pair_both(P, ax, ay, az, A.gq[g], A.gtype[g], A.trial[3 * jj], A.trial[3 * jj + 1], A.trial[3 * jj + 2],
                       A.gq[jj], A.gtype[jj], do_lj, lj_n, re_n);
          if (lj_n >= PG_VLE) acc_ov++;
          acc_pair += lj_n;
          acc_real += re_n;
        } else if (P.use_ewald) {
          double qq = A.gq[g] * A.gq[g];
          if (qq != 0) acc_real += P.real_self_unit * qq;
        }
      }
    }
  } else if (b < A.n_pair_ctas + A.n_k_ctas) {
    // ---------------- reciprocal part: dS(k) for one k per thread
    const int k = (b - A.n_pair_ctas) * PG_KTILE + tid;
    double kx = 0.0, ky = 0.0, kz = 0.0;
    if (k < A.nk) {
      const int4 l = reinterpret_cast<const int4*>(A.kl)[k];
      // same expression as the reference's table: l*2*kPi/L (potential_ewald_coul.cc:89-97)
      const double kPi = 3.14159265359;
      kx = l.x * 2 * kPi / P.ebox[0];
      ky = l.y * 2 * kPi / P.ebox[1];
      kz = l.z * 2 * kPi / P.ebox[2];
    }
    double dre = 0.0, dim = 0.0;
    for (int gbeg = 0; gbeg < A.glen; gbeg += PG_GCHUNK) {
      const int gcnt = min(A.glen - gbeg, PG_GCHUNK);
      __syncthreads();
      if (tid < gcnt) {
        const int g = gbeg + tid;
        if (A.has_old) {
          double2 a = A.xy[A.g0 + g], c = A.zq[A.g0 + g];
          s_o[0][tid] = pg_wrap_pos(a.x, P.ebox[0], P.inv_ebox[0], P.pbc[0]);
          s_o[1][tid] = pg_wrap_pos(a.y, P.ebox[1], P.inv_ebox[1], P.pbc[1]);
          s_o[2][tid] = pg_wrap_pos(c.x, P.ebox[2], P.inv_ebox[2], P.pbc[2]);
        }
        if (A.has_new) {
          s_n[0][tid] = pg_wrap_pos(A.trial[3 * g], P.ebox[0], P.inv_ebox[0], P.pbc[0]);
          s_n[1][tid] = pg_wrap_pos(A.trial[3 * g + 1], P.ebox[1], P.inv_ebox[1], P.pbc[1]);
          s_n[2][tid] = pg_wrap_pos(A.trial[3 * g + 2], P.ebox[2], P.inv_ebox[2], P.pbc[2]);
        }
        s_q[tid] = A.gq[g];
        s_mv[tid] = A.moved[g];
      }
      __syncthreads();
      if (k < A.nk) {
        for (int i = 0; i < gcnt; i++) {
          const double q = s_q[i];
          if (!s_mv[i] || q == 0) continue;
          if (A.has_new) {
            double s, c;
            sincos(kx * s_n[0][i] + ky * s_n[1][i] + kz * s_n[2][i], &s, &c);
            dre += q * c; dim += q * s;
          }
          if (A.has_old) {
            double s, c;
            sincos(kx * s_o[0][i] + ky * s_o[1][i] + kz * s_o[2][i], &s, &c);
            dre -= q * c; dim -= q * s;
          }
        }
      }
    }
    if (k < A.nk) {
      const double2 S = A.S[k];
      // never |S_new|^2 - |S_old|^2: 2 Re(conj(S) dS) + |dS|^2; x2 for the -k half
      acc_rec = 2.0 * A.ek2[k] * (2.0 * (S.x * dre + S.y * dim) + (dre * dre + dim * dim));
      A.dS[k] = make_double2(dre, dim);
    }
  }

  // ---------------- per-CTA partials, then the last CTA to finish finalises
  {
    double v[4] = {acc_pair, acc_real, acc_mz, acc_rec};
    block_sum<4>(v, s_red);
    int ov = warp_sum_i(acc_ov);
    if ((tid & 31) == 0) s_redi[tid >> 5] = ov;
    __syncthreads();
    if (tid == 0) {
      int ovs = 0;
      for (int w = 0; w < (PG_TILE >> 5); w++) ovs += s_redi[w];
      A.partial[4 * b + 0] = v[0];
      A.partial[4 * b + 1] = v[1];
      A.partial[4 * b + 2] = v[2];
      A.partial[4 * b + 3] = v[3];
      A.partial_i[b] = ovs;
      __threadfence();
      unsigned int done = atomicAdd(&A.state->done_counter, 1u);
      s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
  }
  __threadfence();

  // ======================= finalisation (one CTA) =======================
  const int n_ctas = gridDim.x;
  double t_pair = 0.0, t_real = 0.0, t_mz = 0.0, t_rec = 0.0;
  int t_ov = 0;
  for (int c = tid; c < n_ctas; c += PG_TILE) {
    t_pair += __ldcg(&A.partial[4 * c + 0]);
    t_real += __ldcg(&A.partial[4 * c + 1]);
    t_mz += __ldcg(&A.partial[4 * c + 2]);
    t_rec += __ldcg(&A.partial[4 * c + 3]);
    t_ov += __ldcg(&A.partial_i[c]);
  }
  // O(group) terms: walls, bonds, self, group dipole moments
  double w_sum = 0.0, b_sum = 0.0, self_sum = 0.0, mzg_o = 0.0, mzg_n = 0.0;
  int w_out = 0;
  for (int g = tid; g < A.glen; g += PG_TILE) {
    const int mv = A.moved[g];
    const double q = A.gq[g];
    const int t = A.gtype[g];
    double oz = 0.0, nz = 0.0;
    if (A.has_old) oz = A.zq[A.g0 + g].x;
    if (A.has_new) nz = A.trial[3 * g + 2];
    if (A.has_old) mzg_o += q * oz;
    if (A.has_new) mzg_n += q * nz;
    if (mv) {
      if (P.ext_kind != 0) {
        double en = 0.0, eo = 0.0;
        if (A.has_new) {
          en = pg_wall_energy(P, nz, t);
          if (en >= PG_VLE) w_out = 1;
        }
        if (A.has_old) eo = pg_wall_energy(P, oz, t);
        w_sum += en - eo;
      }
      if (P.use_ewald) self_sum += P.self_pref * q * q;
    }
    if (P.bond_kind != 0 && g >= A.bond_first && g + 1 < A.bond_first + A.bond_len) {
      double en = 0.0, eo = 0.0;
      if (A.has_new)
        en = pg_bond_energy(P, A.trial[3 * g], A.trial[3 * g + 1], A.trial[3 * g + 2], A.trial[3 * g + 3],
                            A.trial[3 * g + 4], A.trial[3 * g + 5]);
      if (A.has_old) {
        double2 a = A.xy[A.g0 + g], c = A.zq[A.g0 + g], a2 = A.xy[A.g0 + g + 1], c2 = A.zq[A.g0 + g + 1];
        eo = pg_bond_energy(P, a.x, a.y, c.x, a2.x, a2.y, c2.x);
      }
      b_sum += en - eo;
    }
  }
  double v[8] = {t_pair, t_real, t_mz, t_rec, w_sum, b_sum, self_sum, mzg_o};
  block_sum<8>(v, s_red);
  double v2[1] = {mzg_n};
  block_sum<1>(v2, s_red);
  int ovw = warp_sum_i(t_ov), wow = warp_sum_i(w_out);
  __syncthreads();
  if ((tid & 31) == 0) { s_redi[tid >> 5] = ovw; s_redi[8 + (tid >> 5)] = wow; }
  __syncthreads();
  if (tid != 0) return;
  int n_overlap = 0, wall_out = 0;
  for (int w = 0; w < (PG_TILE >> 5); w++) { n_overlap += s_redi[w]; wall_out += s_redi[8 + w]; }

  PgState* st = A.state;
  const double d_pair = v[0], d_real = v[1], mz_cur = v[2];
  const double d_recip = P.use_ewald ? P.recip_pref * v[3] : 0.0;
  double d_ext = v[4];
  const double d_bond = v[5], d_self = v[6], mzgo = v[7], mzgn = v2[0];
  double dE = 0.0, d_ewald = 0.0, d_dip = 0.0;
  int stage = 0;
  if (A.mode == PG_MODE_MOVE) {
    // ForceField::EnergyDifference, force_field.cc:407-434
    if (wall_out) d_ext = PG_VLE;   // potential_external.cc:107-110: exactly 1e8, partial sum dropped
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
        d_ewald = d_real + d_recip;   // self terms cancel for a move
        if (P.dipole) {
          // DipoleE(mols) reads CURRENT coordinates: lags one accepted move
          st->trial_dipl = P.dipole_pref * mz_cur * mz_cur;
          d_dip = st->trial_dipl - st->cur_dipl;
          d_ewald += d_dip;
        }
        dE += d_ewald;
      }
      if (P.bond_kind != 0) dE += d_bond;
    }
  } else if (A.mode == PG_MODE_INSERT) {
    d_ewald = d_real + d_recip + d_self;
    if (P.use_ewald && P.dipole) {
      // EnergyInitForLastMol, potential_ewald.cc:420-425
      const double m = mz_cur + mzgn;
      const double cur = P.dipole_pref * m * m;
      d_dip = cur - st->trial_dipl;
      d_ewald += d_dip;
      st->trial_dipl = cur;   // committed as cur_dipl by k_commit
    }
    dE = d_pair + d_ext + d_ewald + d_bond;
  } else {
    d_ewald = d_real + d_recip - d_self;   // energies REMOVED carry the sign of a change here
    if (P.use_ewald && P.dipole) {
      // AdjustEnergyUponMolDeletion, potential_ewald.cc:699-705
      const double m = mz_cur - mzgo;
      const double dip = P.dipole_pref * m * m;
      d_dip = dip - st->cur_dipl;
      d_ewald += d_dip;
      st->trial_dipl = dip;
    }
    dE = d_pair + d_ext + d_ewald + d_bond;
  }
  int accept = 0;
  if (A.decide_on_device) {
    // simulation.cc:327-332: no variate is consumed when dE >= kVeryLargeEnergy
    accept = (dE < PG_VLE) && (A.u < exp(-P.beta * dE));
    if (A.replay_dE) A.replay_dE[A.replay_index] = dE;
    if (A.replay_acc) A.replay_acc[A.replay_index] = (uint8_t)accept;
  }
  st->dE = dE; st->d_pair = d_pair; st->d_ext = d_ext; st->d_ewald = d_ewald; st->d_bond = d_bond;
  st->d_real = d_real; st->d_recip = d_recip; st->d_self = d_self; st->d_dipole = d_dip;
  st->mz_current = mz_cur; st->mz_group_old = mzgo; st->mz_group_new = mzgn;
  st->stage = stage; st->n_overlap = n_overlap; st->accept = accept; st->mode = A.mode;
  st->done_counter = 0;
  if (A.result) {
    PgResult* r = A.result;
    r->dE = dE; r->pair = d_pair; r->ext = d_ext; r->ewald = d_ewald; r->bond = d_bond;
    r->real = d_real; r->recip = d_recip; r->mz_current = mz_cur; r->self_e = d_self; r->dipole = d_dip;
    r->stage = stage; r->n_overlap = n_overlap; r->accept = accept;
    __threadfence_system();
    r->seq = A.seq;
  }
}

// ------------------------------------------------------------------ k_commit
// accept_flag: 0/1 from the host, or -1 to use the decision k_delta took on the device.
__global__ void __launch_bounds__(256) k_commit(const PgDev P, int accept_flag, int mode, int g0, int glen,
                                                const double* __restrict__ trial, const double* __restrict__ gq,
                                                const int* __restrict__ gtype, double2* xy, double2* zq, int* type,
                                                double2* S, const double2* __restrict__ dS, int nk, PgState* st) {
  const int accept = (accept_flag < 0) ? st->accept : accept_flag;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (accept) {
    if (mode != PG_MODE_DELETE && i < glen) {
      xy[g0 + i] = make_double2(trial[3 * i], trial[3 * i + 1]);
      zq[g0 + i] = make_double2(trial[3 * i + 2], gq[i]);
      type[g0 + i] = gtype[i];
    }
    if (i < nk) {
      double2 s = S[i], d = dS[i];
      S[i] = make_double2(s.x + d.x, s.y + d.y);
    }
  }
  if (i == 0) {
    if (accept) {
      // FinalizeEnergyBothMaps: E_tot += dE for each potential
      st->E_pair += st->d_pair;
      st->E_ewald += st->d_ewald;
      st->E_bond += st->d_bond;
      st->E_ext += st->d_ext;
      st->E_real += st->d_real;
      st->E_recip += st->d_recip;
      st->E_self += (mode == PG_MODE_DELETE) ? -st->d_self : (mode == PG_MODE_INSERT ? st->d_self : 0.0);
      st->cur_dipl = st->trial_dipl;
      if (mode == PG_MODE_INSERT) st->n_beads += glen;
    } else {
      st->trial_dipl = st->cur_dipl;
    }
  }
}

// Remove the bead range [b0, b1) by shifting the tail down through a scratch copy.
__global__ void k_gather_tail(const double2* xy, const double2* zq, const int* type, const int* mol, int b1, int n,
                              double2* t_xy, double2* t_zq, int* t_type, int* t_mol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (b1 + i < n) { t_xy[i] = xy[b1 + i]; t_zq[i] = zq[b1 + i]; t_type[i] = type[b1 + i]; t_mol[i] = mol[b1 + i]; }
}
__global__ void k_scatter_tail(double2* xy, double2* zq, int* type, int* mol, int b0, int cnt, int mol_shift,
                               const double2* t_xy, const double2* t_zq, const int* t_type, const int* t_mol,
                               PgState* st, int n_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) { xy[b0 + i] = t_xy[i]; zq[b0 + i] = t_zq[i]; type[b0 + i] = t_type[i]; mol[b0 + i] = t_mol[i] - mol_shift; }
  if (i == 0) st->n_beads = n_new;
}

// --------------------------------------------------------------- k_tot_pairs
// All pairs i <= j: LJ (i < j) and Ewald real-space (i < j full, i == j self-image x0.5).
// 2-D grid over the triangle: CTA (r, c) owns rows [r*PG_TILE, +PG_TILE) x partners [c*pc, +pc), one row bead per
// thread, the partner slice staged in shared memory; CTAs entirely below the diagonal only write their zero.  The host
// picks pc (<= PG_TILE) so that small systems still spread over the machine (potential_pair.cc:55-100 and
// potential_ewald.cc EnergyInitialization are O(N^2) loops in the reference).
__global__ void __launch_bounds__(PG_TILE) k_tot_pairs(const PgDev P, const double2* __restrict__ xy,
                                                       const double2* __restrict__ zq, const int* __restrict__ type,
                                                       const int* __restrict__ mol, int n, int pc, double* partial) {
  __shared__ double s_x[PG_TILE], s_y[PG_TILE], s_z[PG_TILE], s_q[PG_TILE];
  __shared__ int s_t[PG_TILE], s_m[PG_TILE];
  __shared__ double s_red[2 * 32];
  const int tid = threadIdx.x;
  const int i = blockIdx.x * PG_TILE + tid;
  const int t0 = blockIdx.y * pc;
  const size_t slot = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  if (t0 + pc <= blockIdx.x * PG_TILE || t0 >= n) {   // no partner j >= i in this slice
    if (tid == 0) { partial[2 * slot] = 0.0; partial[2 * slot + 1] = 0.0; }
    return;
  }
  double ax = 0, ay = 0, az = 0, aq = 0;
  int at = 0, am = -1;
  if (i < n) {
    double2 a = xy[i], c = zq[i];
    ax = a.x; ay = a.y; az = c.x; aq = c.y; at = type[i]; am = mol[i];
  }
  double e_lj = 0.0, e_re = 0.0;
  const int cnt = min(pc, n - t0);
  if (tid < cnt) {
    const int jl = t0 + tid;
    double2 a = xy[jl], c = zq[jl];
    s_x[tid] = a.x; s_y[tid] = a.y; s_z[tid] = c.x; s_q[tid] = c.y; s_t[tid] = type[jl]; s_m[tid] = mol[jl];
  }
  __syncthreads();
  if (i < n) {
    for (int jj = 0; jj < cnt; jj++) {
      const int j = t0 + jj;
      if (j < i) continue;
      if (j > i) {
        int do_lj = (P.pair_kind != 0);
        if (P.pair_kind == 2 && j == i + 1 && s_m[jj] == am) do_lj = 0;
        double lj, re;
        pg_pair_both(P, ax, ay, az, aq, at, s_x[jj], s_y[jj], s_z[jj], s_q[jj], s_t[jj], do_lj, lj, re);
        e_lj += lj; e_re += re;
      } else if (P.use_ewald) {
        double qq = aq * aq;
        if (qq != 0) e_re += P.real_self_unit * qq;
      }
    }
  }
  double v[2] = {e_lj, e_re};
  block_sum<2>(v, s_red);
  if (tid == 0) { partial[2 * slot] = v[0]; partial[2 * slot + 1] = v[1]; }
}

// ---------------------------------------------------------------- k_sk_slice
// S(k) = sum_i q_i exp(i k.r_i) for k in [k_first, k_first + k_count) of the half list.
// Grid (k tiles, bead chunks): one k per thread, the CTA's bead chunk staged tile by tile in shared
// memory (wrapped once).  Each (chunk, k) partial sum has one writer; k_sk_reduce adds the chunks in
// index order, so the result does not depend on how the k list is sliced over ranks.
__global__ void __launch_bounds__(PG_TILE) k_sk_slice(const PgDev P, const double2* __restrict__ xy,
                                                      const double2* __restrict__ zq, int n,
                                                      const double4* __restrict__ kvec, int k_first, int k_count,
                                                      int chunk_beads, double2* partial) {
  __shared__ double s_x[PG_TILE], s_y[PG_TILE], s_z[PG_TILE], s_q[PG_TILE];
  const int tid = threadIdx.x;
  const int kk = blockIdx.x * PG_TILE + tid;
  const bool active = kk < k_count;
  double4 kv = make_double4(0.0, 0.0, 0.0, 0.0);
  if (active) kv = kvec[k_first + kk];
  const int j0 = blockIdx.y * chunk_beads, j1 = min(n, j0 + chunk_beads);
  double re = 0.0, im = 0.0;
  for (int t0 = j0; t0 < j1; t0 += PG_TILE) {
    __syncthreads();
    const int j = t0 + tid;
    if (j < j1) {
      double2 a = xy[j], c = zq[j];
      s_x[tid] = pg_wrap_pos(a.x, P.ebox[0], P.inv_ebox[0], P.pbc[0]);
      s_y[tid] = pg_wrap_pos(a.y, P.ebox[1], P.inv_ebox[1], P.pbc[1]);
      s_z[tid] = pg_wrap_pos(c.x, P.ebox[2], P.inv_ebox[2], P.pbc[2]);
      s_q[tid] = c.y;
    } else {
      s_q[tid] = 0.0;
    }
    __syncthreads();
    if (active) {
      const int cnt = min(PG_TILE, j1 - t0);
      for (int jj = 0; jj < cnt; jj++) {
        const double q = s_q[jj];
        if (q == 0) continue;
        double s, c;
        sincos(kv.x * s_x[jj] + kv.y * s_y[jj] + kv.z * s_z[jj], &s, &c);
        re += q * c; im += q * s;
      }
    }
  }
  if (active) partial[(size_t)blockIdx.y * k_count + kk] = make_double2(re, im);
}

__global__ void __launch_bounds__(PG_TILE) k_sk_reduce(const double2* __restrict__ partial, int n_chunks, int k_count,
                                                       double2* out) {
  const int kk = blockIdx.x * PG_TILE + threadIdx.x;
  if (kk >= k_count) return;
  double re = 0.0, im = 0.0;
  for (int c = 0; c < n_chunks; c++) {
    const double2 v = partial[(size_t)c * k_count + kk];
    re += v.x; im += v.y;
  }
  out[kk] = make_double2(re, im);
}

// Reciprocal energy of a k slice: recip_pref * sum 2 ek2 |S|^2 (single CTA).
__global__ void __launch_bounds__(256) k_sk_energy(const PgDev P, const double2* __restrict__ S,
                                                   const double* __restrict__ ek2, int k_first, int k_count,
                                                   double* out) {
  __shared__ double s_red[32];
  double acc = 0.0;
  for (int k = threadIdx.x; k < k_count; k += blockDim.x) {
    double2 s = S[k_first + k];
    acc += 2.0 * ek2[k_first + k] * (s.x * s.x + s.y * s.y);
  }
  double v[1] = {acc};
  block_sum<1>(v, s_red);
  if (threadIdx.x == 0) out[0] = P.recip_pref * v[0];
}

// --------------------------------------------------------------- k_tot_final
// O(N)+O(K) totals and the reduction of k_tot_pairs' partials (single CTA).
// out: pair, ewald, bond, ext, real, recip, self, dipole
__global__ void __launch_bounds__(256) k_tot_final(const PgDev P, const double2* __restrict__ xy,
                                                   const double2* __restrict__ zq, const int* __restrict__ type,
                                                   const int* __restrict__ mol, int n, const double* pair_partial,
                                                   int n_partial, const double2* __restrict__ S,
                                                   const double* __restrict__ ek2, int nk, double* out,
                                                   PgState* st, int set_state) {
  __shared__ double s_red[8 * 32];
  double lj = 0, re = 0, wall = 0, bond = 0, self = 0, mz = 0, rec = 0;
  for (int c = threadIdx.x; c < n_partial; c += blockDim.x) { lj += pair_partial[2 * c]; re += pair_partial[2 * c + 1]; }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double2 a = xy[i], c = zq[i];
    const int t = type[i];
    if (P.ext_kind != 0) wall += pg_wall_energy(P, c.x, t);
    if (P.use_ewald) { self += P.self_pref * c.y * c.y; mz += c.y * c.x; }
    if (P.bond_kind != 0 && i + 1 < n && mol[i + 1] == mol[i]) {
      double2 a2 = xy[i + 1], c2 = zq[i + 1];
      bond += pg_bond_energy(P, a.x, a.y, c.x, a2.x, a2.y, c2.x);
    }
  }
  if (P.use_ewald)
    for (int k = threadIdx.x; k < nk; k += blockDim.x) {
      double2 s = S[k];
      rec += 2.0 * ek2[k] * (s.x * s.x + s.y * s.y);
    }
  double v[7] = {lj, re, wall, bond, self, mz, rec};
  block_sum<7>(v, s_red);
  if (threadIdx.x == 0) {
    const double recip = P.use_ewald ? P.recip_pref * v[6] : 0.0;
    const double dip = (P.use_ewald && P.dipole) ? P.dipole_pref * v[5] * v[5] : 0.0;
    const double ewald = P.use_ewald ? (v[1] + recip + v[4] + dip) : 0.0;
    out[0] = v[0]; out[1] = ewald; out[2] = v[3]; out[3] = v[2];
    out[4] = v[1]; out[5] = recip; out[6] = v[4]; out[7] = dip;
    if (set_state) {
      st->E_pair = v[0]; st->E_ewald = ewald; st->E_bond = v[3]; st->E_ext = v[2];
      st->E_real = v[1]; st->E_recip = recip; st->E_self = v[4];
      st->cur_dipl = dip; st->trial_dipl = dip;
      st->n_beads = n; st->done_counter = 0;
    }
  }
}

// --------------------------------------------------------------- FP64 peak
// Dependent-chain-free DFMA loop: 8 independent accumulators per thread.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;   // never true; keeps the loop alive
}


// L2 streaming-read microbenchmark (roofline denominator next to k_fp64_peak): every thread block sweeps an L2-resident
// buffer with 16-byte loads, `reps` times; the sum keeps the loads alive.
__global__ void __launch_bounds__(256) k_l2_read(const int4* __restrict__ buf, size_t n16, int reps, int* sink) {
  int acc = 0;
  for (int r = 0; r < reps; r++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
      const int4 v = __ldcg(buf + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x7fffffff) *sink = acc;
}
