// k_wall_force — one sample of the slab wall-force pressure sampler.
//
// Replaces the force sums of ForceField::CalcPressureForceLJELSlit (reference
// src/force_field/pressure.cc:404-469): for every wall site i (the `phantom` first molecules, one bead
// each, half on each plate) and every bead b of the molecules phantom/2 .. n_mol-1
//   site-site LJ      C * r_z/|r| * PairForce             (potential_truncated_lj.cc:87-122)
//   site-site Ewald   C * (PairForceZReal + PairForceZRepl) (potential_ewald_coul.cc:261-324)
// and for every non-wall bead 0.5 * BeadForceOnWall (potential_truncated_lj_wall.cc:136-217), summed
// per class of b (0 ion, 1 polymer bead, 2 wall site).  The host keeps the running averages.
// PairForce: the reference branches on an uninitialised r6 (undefined behaviour); this is the evident
// intent — r <= 0 gives 1e8, otherwise the LJ force.
//
// One thread per (site, bead) pair, then one thread per bead for the plate term; every block writes six
// partial sums, the host adds them in block order.  A sampler, not the per-move path: it runs every
// 10 * sampling_frequency steps.
#pragma once
#include "pg_kernels.cuh"

#define WF_THREADS 128

struct PgWallForceArgs {
  const double2* xy; const double2* zq; const int* type; const unsigned char* klass;
  int n, phantom, b_first;   // b_first = first bead of molecule phantom/2
  const double4* kvec; int nk;
  double* partial;           // [blocks][6]
};

__device__ double wf_pair_force(const PgDev& P, double r, int tp) {
  if (P.pair_kind == 2) return 0.0;
  if (r <= 0) return PG_VLE;
  const double sigma = P.lj_sigma[tp], eps4 = P.lj_eps4[tp];
  double force = 0.0;
  if (P.lj_cutoff < 0) {
    if (r < P.lj_rcut[tp]) {   // k216 * sigma
      const double r6 = pg_pow6(sigma / r);
      force = eps4 * (r6 * r6 * 12 / r - r6 * 6 / r);
    }
  } else if (r < P.lj_cutoff) {
    const double r6 = pg_pow6(sigma / r);
    force = eps4 * (r6 * r6 * 12 / r - r6 * 6 / r);
  } else {
    const double r6 = pg_pow6(sigma / P.lj_cutoff);
    force = eps4 * (r6 * r6 * 12 / r - r6 * 6 / r);
  }
  return force;
}

__device__ double wf_bead_force_on_wall(const PgDev& P, double z, int t) {
  if (P.ext_kind != 1) return 0.0;   // hard and well walls exert no force
  const double Lz = P.box[2];
  const double k213 = 1.25992104989, c = 2.59807621135;
  const double sigma = P.wall_sigma[t], epsilon = P.wall_eps[t];
  if (epsilon == 0) return 0.0;
  if (z <= 0 || z >= Lz) return PG_VLE;
  const double R0 = 3 * k213 * sigma;
  const int graft = P.graft[t];
  double force = 0.0;
  if (P.wall_cut < 0 && graft == 0) {
    if (z < k213 * sigma) { const double r3 = pg_pow3(sigma / z); force += c * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z); }
    if (Lz - z < k213 * sigma) {
      const double r3 = pg_pow3(sigma / (Lz - z));
      force += c * epsilon * (r3 * r3 * 6 / (Lz - z) - r3 * 3 / (Lz - z));
    }
  } else if (graft == 1 || graft == 2) {
    const double zn = (graft == 1) ? z : (Lz - z), zf = (graft == 1) ? (Lz - z) : z;
    const double r3 = pg_pow3(sigma / zn);
    force += c * epsilon * (r3 * r3 * 6 / zn - r3 * 3 / zn);
    const double t2 = zn / R0;
    force += -1.0 * R0 * zn / (1 - t2 * t2);   // FENE part
    if (zf < k213 * sigma) { const double r3b = pg_pow3(sigma / zf); force += c * epsilon * (r3b * r3b * 6 / zf - r3b * 3 / zf); }
  } else {
    const double m_cut = P.wall_cut;
    double r3 = pg_pow3(sigma / (z < m_cut ? z : m_cut));
    force += c * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);
    if (Lz - z < m_cut) {
      r3 = pg_pow3(sigma / (Lz - z));
      force += c * epsilon * (r3 * r3 * 6 / (Lz - z) - r3 * 3 / (Lz - z));
    } else {
      r3 = pg_pow3(sigma / m_cut);
      force += c * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);   // the reference uses z here (not Lz - z)
    }
  }
  return force;
}

__global__ void __launch_bounds__(WF_THREADS) k_wall_force(const PgDev P, const PgWallForceArgs A) {
  __shared__ double s_red[6 * 32];
  const int tid = threadIdx.x;
  const long long idx = (long long)blockIdx.x * WF_THREADS + tid;
  const int nb = A.n - A.b_first;                       // beads every site interacts with
  const long long n_pairs = (long long)A.phantom * nb;
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (idx < n_pairs) {
    const int i = (int)(idx / nb), b = A.b_first + (int)(idx - (long long)i * nb);
    const int kid = A.klass[b];                          // 0 ion, 1 polymer, 2 wall site
    const bool sys = kid != 2;
    const double C = !sys ? -1.0 : (i < A.phantom / 2 ? -0.5 : 0.5);
    if (i < A.phantom / 2 || sys) {
      const double2 ai = A.xy[i], ci = A.zq[i], ab = A.xy[b], cb = A.zq[b];
      if (P.pair_kind != 0) {
        // GetDistVector(bead b -> site i) in the slab box for the direction, BBDist for the magnitude
        double rx = ai.x - ab.x, ry = ai.y - ab.y, rz = ci.x - cb.x;
        if (P.pbc[0]) rx = pg_wrap(rx, P.box[0], P.inv_box[0]);
        if (P.pbc[1]) ry = pg_wrap(ry, P.box[1], P.inv_box[1]);
        if (P.pbc[2]) rz = pg_wrap(rz, P.box[2], P.inv_box[2]);
        const double zc = rz / sqrt(rx * rx + ry * ry + rz * rz);
        const double dx = pg_bbdist_axis(ab.x - ai.x, P.box[0], P.inv_box[0], P.half_box[0], P.pbc[0]);
        const double dy = pg_bbdist_axis(ab.y - ai.y, P.box[1], P.inv_box[1], P.half_box[1], P.pbc[1]);
        const double dz = pg_bbdist_axis(cb.x - ci.x, P.box[2], P.inv_box[2], P.half_box[2], P.pbc[2]);
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        acc[kid] += C * zc * wf_pair_force(P, r, A.type[b] * PG_MAX_TYPES + A.type[i]);
      }
      const double qq = ci.y * cb.y;
      if (P.use_ewald && qq != 0.0) {
        double dx, dy, dz;   // site - bead, wrapped in the padded Ewald box
        pg_dist_ewald(P, ab.x, ab.y, cb.x, ai.x, ai.y, ci.x, dx, dy, dz);
        // real space: images that can reach the cutoff, then the reference's own test
        const double rc = P.rc_relaxed;
        const int cx = P.single_image ? 0 : P.real_cell[0], cy = P.single_image ? 0 : P.real_cell[1],
                  cz = P.single_image ? 0 : P.real_cell[2];
        int i0 = max((int)ceil((-rc - dx) * P.inv_ebox[0]), -cx), i1 = min((int)floor((rc - dx) * P.inv_ebox[0]), cx);
        int j0 = max((int)ceil((-rc - dy) * P.inv_ebox[1]), -cy), j1 = min((int)floor((rc - dy) * P.inv_ebox[1]), cy);
        int k0 = max((int)ceil((-rc - dz) * P.inv_ebox[2]), -cz), k1 = min((int)floor((rc - dz) * P.inv_ebox[2]), cz);
        const double alpha = P.sqrt_alpha * P.sqrt_alpha;
        const double pre2 = 2 * P.sqrt_alpha / sqrt(3.14159265359);   // 2 sqrt(alpha / kPi)
        double fr = 0.0;
        for (int ii = i0; ii <= i1; ii++)
          for (int jj = j0; jj <= j1; jj++)
            for (int kk = k0; kk <= k1; kk++) {
              const double vx = dx + ii * P.ebox[0], vy = dy + jj * P.ebox[1], vz = dz + kk * P.ebox[2];
              const double d2 = vx * vx + vy * vy + vz * vz;
              if (d2 > P.rc2_relaxed) continue;
              const double d = sqrt(d2);
              if (d > 0 && d <= P.real_cutoff)
                fr += P.lB * qq * ((pre2 * exp(-alpha * d * d)) + (erfc(P.sqrt_alpha * d) / d)) * vz / (d * d);
            }
        // reciprocal space: kz ek2 sin(k.r) is even in k, so the half list counts twice
        double fk = 0.0;
        for (int k = 0; k < A.nk; k++) {
          const double4 kv = A.kvec[k];
          fk += kv.z * kv.w * sin(kv.x * dx + kv.y * dy + kv.z * dz);
        }
        const double ebox_vol = P.ebox[0] * P.ebox[1] * P.ebox[2];
        fk *= 2.0 * (P.lB * 4 * 3.14159265359 * qq / ebox_vol);
        acc[3 + kid] += C * fr + C * fk;
      }
    }
  } else if (idx < n_pairs + (A.n - A.phantom) && P.ext_kind != 0) {
    const int b = A.phantom + (int)(idx - n_pairs);
    acc[A.klass[b]] += 0.5 * wf_bead_force_on_wall(P, A.zq[b].x, A.type[b]);
  }
  block_sum<6>(acc, s_red);
  if (tid == 0)
    for (int k = 0; k < 6; k++) A.partial[6 * (size_t)blockIdx.x + k] = acc[k];
}
