// Volume-perturbation pressure sampler: one sample of the energy change under a virtual stretch of the box
// along z by dz, split by species pair.
//
// Replaces the loops of ForceField::CalcPressureVolScalingHSELSlit (reference
// src/force_field/pressure.cc:187-387; called every 10 * sampling_frequency steps when there is no wall
// potential, src/simulation/simulation.cc:756-761):
//   * per molecule the displacement d_com_z = dz * com_z / Lz (grafted "R": dz, "L": 0), per bead the trial
//     height z + d_com_z, or z * (1 + dz / Lz) when a bond potential is on (pressure.cc:208-247);
//   * for every pair of beads (i <= k in molecule order, wall sites restricted as pressure.cc:264-265)
//       Ewald:  PairEnergyRealForP + PairEnergyReplForP (potential_ewald_coul.cc:167-250; cell edge and k_z
//               of the stretched box), half of it for a bead with itself, minus the stored pair energy;
//       LJ/HS:  PairEnergy in the stretched box minus the stored pair energy (mobile beads only);
//     both summed into the 4x4 species table (0 cation, 1 anion, 2 polymer, 3 surface);
//   * wall, bond and dipole terms (pressure.cc:300-338).
//
// The reference evaluates the reciprocal part pair by pair (O(N^2 K)); here it is the same sum written with
// one structure factor per species group and geometry:  for groups a != b
//   sum_{i in a, j in b} E_repl(i,j) = (4 pi lB / V) * 2 * sum_{k in half space} ek2(k) Re(S_a(k) conj S_b(k)),
// and for a == b (pairs once, self pair halved) (4 pi lB / V) * sum_half ek2 |S_a|^2.  The difference
// stretched - current is formed per k vector before the sum so the cancellation stays inside one term.
// The k sets of the two geometries may differ at the cutoff sphere; the host lists the union with a zero
// weight where a vector is outside one of them.
//
// Four launches: k_vol_prep (one thread per molecule), k_vol_pairs (row bead per thread, partner slices in
// shared memory, the triangle tiled over the machine as k_tot_pairs), k_vol_recip (one thread per k vector), k_vol_final (one CTA, fixed
// summation order).  A sampler, not the per-move path.
#pragma once
#include "pg_kernels.cuh"

#define VS_FINAL_THREADS 256

struct PgVolArgs {
  const double2* xy; const double2* zq; const int* type; const int* mol; const int* mol_first;
  int n, n_mol, phantom;
  double dz;
  const PgDev* Pz;           // parameter block of the stretched box (device copy)
  double* z_new;             // [n] trial heights
  unsigned char* grp;        // [n] 0 cation, 1 anion, 2 polymer, 3 surface (first half), 4 surface (second half)
  double* mol_part;          // [n_mol][4] bond dU, wall dU, Mz old, Mz new
  int pc, n_pair_part;       // partners per CTA slice; number of 32-value partials
  double* pair_part;         // [n_pair_part][32] el[16] hs[16] of one CTA of k_vol_pairs
  // reciprocal space, union list of both geometries
  const int* kl;             // [nku][4]
  const double* kw;          // [nku][2] ek2 of the current / stretched box (0: vector not in that set)
  int nku;
  double* k_part;            // [nku][16]
  double* out;               // [36] el[16] hs[16] bond dipole dU n_free
};

// ----------------------------------------------------------------- k_vol_prep
__global__ void __launch_bounds__(128) k_vol_prep(const PgDev P, const PgVolArgs A) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= A.n_mol) return;
  const PgDev& Pz = *A.Pz;
  const int first = A.mol_first[m], last = A.mol_first[m + 1], len = last - first;
  const double Lz = P.box[2];
  double com_z = 0.0;
  for (int b = first; b < last; b++) com_z += A.zq[b].x;
  com_z /= len;
  double d_com = A.dz * (com_z / Lz);
  const int g0 = P.graft[A.type[first]];
  if (g0 == 2) d_com = A.dz;        // "R": moves with the far plate
  else if (g0 == 1) d_com = 0.0;    // "L": stays
  unsigned char g;
  if (m < A.phantom) g = (m < A.phantom / 2) ? 3 : 4;
  else if (len > 1) g = 2;
  else g = (A.zq[first].y >= 0) ? 0 : 1;
  double mz_old = 0.0, mz_new = 0.0, d_wall = 0.0;
  for (int b = first; b < last; b++) {
    const double2 c = A.zq[b];
    const int t = A.type[b];
    double zn;
    if (P.bond_kind != 0 && P.graft[t] == 0) zn = c.x * (1.0 + A.dz / Lz);
    else zn = c.x + d_com;
    A.z_new[b] = zn;
    A.grp[b] = g;
    mz_old += c.y * c.x;
    mz_new += c.y * zn;
    if (P.ext_kind != 0 && m >= A.phantom) d_wall += pg_wall_energy(Pz, zn, t) - pg_wall_energy(P, c.x, t);
  }
  double d_bond = 0.0;
  if (P.bond_kind != 0) {
    double e_old = 0.0, e_new = 0.0;
    for (int b = first; b + 1 < last; b++) {
      const double2 a = A.xy[b], a2 = A.xy[b + 1];
      e_old += pg_bond_energy(P, a.x, a.y, A.zq[b].x, a2.x, a2.y, A.zq[b + 1].x);
      e_new += pg_bond_energy(P, a.x, a.y, A.z_new[b], a2.x, a2.y, A.z_new[b + 1]);
    }
    d_bond = e_new - e_old;
  }
  double* o = A.mol_part + 4 * (size_t)m;
  o[0] = d_bond; o[1] = d_wall; o[2] = mz_old; o[3] = mz_new;
}

// ---------------------------------------------------------------- k_vol_pairs
// 2-D grid over the triangle like k_tot_pairs: CTA (r, c) owns rows [r*PG_TILE, +PG_TILE) x partners [c*pc, +pc).
// Every thread keeps its row bead's sums by partner species; the CTA folds them into the 16 + 16 table entries in a
// fixed order and writes one 32-value partial.
__global__ void __launch_bounds__(PG_TILE) k_vol_pairs(const PgDev P, const PgVolArgs A) {
  __shared__ double s_x[PG_TILE], s_y[PG_TILE], s_z[PG_TILE], s_zn[PG_TILE], s_q[PG_TILE];
  __shared__ int s_t[PG_TILE], s_m[PG_TILE];
  __shared__ unsigned char s_g[PG_TILE];
  __shared__ double s_val[PG_TILE][9];
  __shared__ int s_ca[PG_TILE];
  const PgDev& Pz = *A.Pz;
  const int tid = threadIdx.x, n = A.n, pc = A.pc;
  const int i = blockIdx.x * PG_TILE + tid;
  const int t0 = blockIdx.y * pc;
  double* slot = A.pair_part + 32 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
  if (t0 + pc <= blockIdx.x * PG_TILE || t0 >= n) {   // no partner j >= i in this slice
    if (tid < 32) slot[tid] = 0.0;
    return;
  }
  double ax = 0, ay = 0, az = 0, azn = 0, aq = 0;
  int at = 0, am = -1, ag = 4;
  if (i < n) {
    const double2 a = A.xy[i], c = A.zq[i];
    ax = a.x; ay = a.y; az = c.x; aq = c.y; azn = A.z_new[i]; at = A.type[i]; am = A.mol[i]; ag = A.grp[i];
  }
  double el[4] = {0, 0, 0, 0}, hs[4] = {0, 0, 0, 0};
  const int cnt = min(pc, n - t0);
  if (tid < cnt) {
    const int jl = t0 + tid;
    const double2 a = A.xy[jl], c = A.zq[jl];
    s_x[tid] = a.x; s_y[tid] = a.y; s_z[tid] = c.x; s_q[tid] = c.y; s_zn[tid] = A.z_new[jl];
    s_t[tid] = A.type[jl]; s_m[tid] = A.mol[jl]; s_g[tid] = A.grp[jl];
  }
  __syncthreads();
  if (i < n && ag != 4) {
    for (int jj = 0; jj < cnt; jj++) {
      const int j = t0 + jj;
      if (j < i) continue;
      const int gj = s_g[jj];
      // i >= phantom, or i on the first plate and k >= phantom / 2 (pressure.cc:264-265)
      if (!(ag < 3 || gj != 3)) continue;
      const int cb = gj < 3 ? gj : 3;
      double d_el = 0.0, d_hs = 0.0;
      if (j > i) {
        const int mobile = (ag < 3 && gj < 3);
        const int lj_new = (P.pair_kind != 0) && mobile;
        int lj_old = lj_new;
        // the stored energy of bonded hard-sphere neighbours is 0 (potential_pair.cc:64-67); PairEnergy is not
        if (P.pair_kind == 2 && j == i + 1 && s_m[jj] == am) lj_old = 0;
        double lj_o, re_o, lj_n, re_n;
        pg_pair_both(P, ax, ay, az, aq, at, s_x[jj], s_y[jj], s_z[jj], s_q[jj], s_t[jj], lj_old, lj_o, re_o);
        pg_pair_both(Pz, ax, ay, azn, aq, at, s_x[jj], s_y[jj], s_zn[jj], s_q[jj], s_t[jj], lj_new, lj_n, re_n);
        d_el = re_n - re_o;
        d_hs = lj_n - lj_o;
      } else if (P.use_ewald && ag < 3) {
        const double qq = aq * aq;
        if (qq != 0) d_el = Pz.real_self_unit * qq - P.real_self_unit * qq;
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        el[c] += (cb == c) ? d_el : 0.0;
        hs[c] += (cb == c) ? d_hs : 0.0;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++) { s_val[tid][c] = el[c]; s_val[tid][4 + c] = hs[c]; }
  s_ca[tid] = ag < 3 ? ag : 3;
  __syncthreads();
  if (tid < 32) {
    const int idx = tid & 15, pass = tid >> 4;   // table entry, el / hs
    double sum = 0.0;
    for (int rrow = 0; rrow < PG_TILE; rrow++) {
      const int ca = s_ca[rrow];
#pragma unroll
      for (int cb = 0; cb < 4; cb++) {
        const int e = (ca < cb ? ca : cb) * 4 + (ca > cb ? ca : cb);
        if (e == idx) sum += s_val[rrow][4 * pass + cb];
      }
    }
    slot[tid] = sum;
  }
}

// ---------------------------------------------------------------- k_vol_recip
// One thread per k vector of the union list: the five group structure factors in both geometries, then the
// ten species-pair terms.
__global__ void __launch_bounds__(64) k_vol_recip(const PgDev P, const PgVolArgs A) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= A.nku) return;
  const PgDev& Pz = *A.Pz;
  const double two_pi = 2 * 3.14159265359;   // kPi of src/utilities/constants.h
  const int lx = A.kl[4 * k], ly = A.kl[4 * k + 1], lz = A.kl[4 * k + 2];
  const double kx = lx * two_pi / P.ebox[0], ky = ly * two_pi / P.ebox[1];
  const double kz_o = lz * two_pi / P.ebox[2], kz_n = lz * two_pi / Pz.ebox[2];
  double so_re[5] = {0, 0, 0, 0, 0}, so_im[5] = {0, 0, 0, 0, 0}, sn_re[5] = {0, 0, 0, 0, 0}, sn_im[5] = {0, 0, 0, 0, 0};
  for (int b = 0; b < A.n; b++) {
    const double2 c = A.zq[b];
    if (c.y == 0.0) continue;
    const double2 a = A.xy[b];
    const int g = A.grp[b];
    const double x = pg_wrap_pos(a.x, P.ebox[0], P.inv_ebox[0], P.pbc[0]);
    const double y = pg_wrap_pos(a.y, P.ebox[1], P.inv_ebox[1], P.pbc[1]);
    const double zo = pg_wrap_pos(c.x, P.ebox[2], P.inv_ebox[2], P.pbc[2]);
    const double zn = pg_wrap_pos(A.z_new[b], Pz.ebox[2], Pz.inv_ebox[2], Pz.pbc[2]);
    const double pxy = kx * x + ky * y;
    double s0, c0, s1, c1;
    sincos(pxy + kz_o * zo, &s0, &c0);
    sincos(pxy + kz_n * zn, &s1, &c1);
#pragma unroll
    for (int q = 0; q < 5; q++) {
      const double w = (g == q) ? c.y : 0.0;
      so_re[q] += w * c0; so_im[q] += w * s0;
      sn_re[q] += w * c1; sn_im[q] += w * s1;
    }
  }
  const double wo = A.kw[2 * k] * P.recip_pref, wn = A.kw[2 * k + 1] * Pz.recip_pref;
  double* o = A.k_part + 16 * (size_t)k;
#pragma unroll
  for (int c = 0; c < 16; c++) o[c] = 0.0;
  // mobile species among themselves
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++) {
      const double f = (a == b) ? 2.0 : 4.0;
      const double xo = so_re[a] * so_re[b] + so_im[a] * so_im[b];
      const double xn = sn_re[a] * sn_re[b] + sn_im[a] * sn_im[b];
      o[a * 4 + b] = f * wn * xn - f * wo * xo;
    }
  // first plate with everything but itself: mobile species (index a*4+3) and the second plate (15)
  for (int b = 0; b < 3; b++) {
    const double xo = so_re[3] * so_re[b] + so_im[3] * so_im[b];
    const double xn = sn_re[3] * sn_re[b] + sn_im[3] * sn_im[b];
    o[b * 4 + 3] = 4.0 * wn * xn - 4.0 * wo * xo;
  }
  {
    const double xo = so_re[3] * so_re[4] + so_im[3] * so_im[4];
    const double xn = sn_re[3] * sn_re[4] + sn_im[3] * sn_im[4];
    o[15] = 4.0 * wn * xn - 4.0 * wo * xo;
  }
}

// ---------------------------------------------------------------- k_vol_final
__global__ void __launch_bounds__(VS_FINAL_THREADS) k_vol_final(const PgDev P, const PgVolArgs A) {
  __shared__ double s_acc[16][VS_FINAL_THREADS + 1];
  __shared__ double s_red[4 * 32];
  __shared__ double s_out[36];
  const int tid = threadIdx.x;
  const PgDev& Pz = *A.Pz;
  for (int pass = 0; pass < 2; pass++) {   // 0: el, 1: hs
    for (int c = 0; c < 16; c++) s_acc[c][tid] = 0.0;
    for (int b = tid; b < A.n_pair_part; b += VS_FINAL_THREADS) {
      const double* r = A.pair_part + 32 * (size_t)b + 16 * pass;
      for (int c = 0; c < 16; c++) s_acc[c][tid] += r[c];
    }
    if (pass == 0 && P.use_ewald) {
      for (int k = tid; k < A.nku; k += VS_FINAL_THREADS)
        for (int c = 0; c < 16; c++) s_acc[c][tid] += A.k_part[16 * (size_t)k + c];
    }
    if (pass == 1 && P.ext_kind != 0) {
      for (int m = A.phantom + tid; m < A.n_mol; m += VS_FINAL_THREADS) {
        const int ca = A.grp[A.mol_first[m]];   // 0..2
        s_acc[ca * 4 + 3][tid] += A.mol_part[4 * (size_t)m + 1];
      }
    }
    __syncthreads();
    for (int s = VS_FINAL_THREADS / 2; s > 0; s >>= 1) {
      if (tid < s)
        for (int c = 0; c < 16; c++) s_acc[c][tid] += s_acc[c][tid + s];
      __syncthreads();
    }
    if (tid < 16) s_out[16 * pass + tid] = s_acc[tid][0];
    __syncthreads();
  }
  double v[3] = {0, 0, 0};   // bond, Mz old, Mz new
  for (int m = tid; m < A.n_mol; m += VS_FINAL_THREADS) {
    const double* p = A.mol_part + 4 * (size_t)m;
    v[0] += p[0]; v[1] += p[2]; v[2] += p[3];
  }
  block_sum<3>(v, s_red);
  if (tid == 0) {
    double dipole = 0.0;
    if (P.use_ewald && P.dipole) {
      // pressure.cc:333-335: the slab box (not the padded Ewald cell) on both sides
      const double vol = P.box[0] * P.box[1] * P.box[2];
      dipole = (P.lB * 2 * 3.14159265359) * (v[2] * v[2] / (P.box[0] * P.box[1] * (P.box[2] + A.dz)) - v[1] * v[1] / vol);
    }
    const double bond = (P.bond_kind != 0) ? v[0] : 0.0;
    double dU = 0.0;
    for (int c = 0; c < 32; c++) { A.out[c] = s_out[c]; dU += s_out[c]; }
    dU += bond + dipole;
    A.out[32] = bond; A.out[33] = dipole; A.out[34] = dU; A.out[35] = (double)(A.n_mol - A.phantom);
    (void)Pz;
  }
}
