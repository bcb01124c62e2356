// plum_b200 — FP64 pair arithmetic of the per-trial-move energy path.
//
// Every function here is __host__ __device__ so the arithmetic core can also be
// exercised on a CPU by tests/ (tests/emul/) — the product path only ever calls
// it from the sm_100a kernels in pg_kernels.cu.
//
// Reference arithmetic restated (file:line relative to /root/reference):
//   BBDist                src/molecules/bead.cc:162-176
//   GetDistVector         src/utilities/misc.cc:43-56
//   TruncatedLJ energy    src/force_field/potential_truncated_lj.cc:49-85
//   HardSphere energy     src/force_field/potential_hard_sphere.cc:38-49
//   Ewald real-space pair src/force_field/potential_ewald_coul.cc:134-164
//   wall energies         src/force_field/potential_truncated_lj_wall.cc:46-134,
//                         potential_hard_wall.cc:36-47, potential_well_wall.cc:41-55
//   spring bond           src/force_field/potential_spring.cc:22-48
#ifndef PLUM_B200_PG_MATH_CUH_
#define PLUM_B200_PG_MATH_CUH_

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD inline
#endif

#define PG_MAX_TYPES 8
#define PG_VLE 1.0e8

// Parameter block passed BY VALUE to every kernel (lives in the constant bank).
// All derived quantities are computed on the host with the reference's own
// expressions (pg_host.cc) so that table entries are bit-identical to what the
// reference computes per call.
struct PgDev {
  double box[3], inv_box[3], half_box[3];     // ForceField::box_l (LJ, walls)
  double ebox[3], inv_ebox[3];                // PotentialEwald::box_l (padded)
  int pbc[3];                                 // axis < npbc
  int same_box[3];                            // ebox[i] == box[i]
  int n_types;
  int pair_kind;
  double lj_cutoff;
  double lj_sigma[PG_MAX_TYPES * PG_MAX_TYPES];   // (s1+s2)/2
  double lj_eps4[PG_MAX_TYPES * PG_MAX_TYPES];    // 4*sqrt(e1*e2)
  double lj_rcut[PG_MAX_TYPES * PG_MAX_TYPES];    // k216*sigma or lj_cutoff
  double lj_eref[PG_MAX_TYPES * PG_MAX_TYPES];    // energy_ref
  double lj_rcut2_relaxed[PG_MAX_TYPES * PG_MAX_TYPES];  // (rcut*(1+1e-9))^2, candidate filter
  double lj_rcut2_relaxed_max;                            // max over the type pairs in use
  double hs_allowed[PG_MAX_TYPES * PG_MAX_TYPES]; // R1+R2
  int use_ewald;
  int dipole;
  double lB, sqrt_alpha, real_cutoff;
  double rc_relaxed, rc2_relaxed;             // real_cutoff*(1+1e-9): exact-pruning margin
  int real_cell[3];
  int single_image;                           // 1: only the central image can ever pass r<=real_cutoff
  double recip_pref;                          // 2*kPi*lB/V
  double self_pref;                           // -lB*sqrt(alpha/kPi)
  double real_self_unit;                      // half the real-space energy of a unit charge with its own periodic images
  double dipole_pref;                         // lB*2*kPi/V
  int bond_kind;
  int ext_kind;
  double bond_k, bond_r0;
  double wall_cut;
  double wall_sigma[PG_MAX_TYPES], wall_eps[PG_MAX_TYPES];
  int graft[PG_MAX_TYPES];
  double wall_r3ref_e[PG_MAX_TYPES];          // 2.598..*eps*(r3ref^2-r3ref) of the branch in use
  double well_width, well_depth;
  double beta;
};

// d -= L*round(d/L).  rint (ties-to-even) instead of round (ties away): the two
// differ only for d/L exactly half-integer, where both choices give the same
// |d| and the same set of images inside the real-space cutoff.
PG_HD double pg_wrap(double d, double L, double invL) { return d - L * rint(d * invL); }

PG_HD double pg_pow6(double x) {
  double x2 = x * x;
  double x3 = x2 * x;
  return x3 * x3;
}
PG_HD double pg_pow3(double x) { return x * x * x; }

// Ewald distance vector b - a wrapped in the padded box (GetDistVector).
PG_HD void pg_dist_ewald(const PgDev& P, double ax, double ay, double az, double bx, double by, double bz,
                         double& dx, double& dy, double& dz) {
  dx = bx - ax;
  dy = by - ay;
  dz = bz - az;
  if (P.pbc[0]) dx = pg_wrap(dx, P.ebox[0], P.inv_ebox[0]);
  if (P.pbc[1]) dy = pg_wrap(dy, P.ebox[1], P.inv_ebox[1]);
  if (P.pbc[2]) dz = pg_wrap(dz, P.ebox[2], P.inv_ebox[2]);
}

// BBDist: one axis of the LJ minimum-image separation, given the raw difference.
PG_HD double pg_bbdist_axis(double d, double L, double invL, double halfL, int pbc) {
  if (pbc) {
    d = pg_wrap(d, L, invL);
    if (fabs(d) > halfL) d = L - fabs(d);
  }
  return d;
}

// LJ/WCA or hard-sphere energy for separation r (already minimum-imaged), type pair index tp.
PG_HD double pg_pair_energy_r(const PgDev& P, double r, int tp) {
  if (P.pair_kind == 2) return (r <= P.hs_allowed[tp]) ? PG_VLE : 0.0;
  if (r <= 0) return PG_VLE;
  if (r < P.lj_rcut[tp]) {
    double r6 = pg_pow6(P.lj_sigma[tp] / r);
    return P.lj_eps4[tp] * (r6 * r6 - r6) - P.lj_eref[tp];
  }
  return 0.0;
}

// Full LJ pair energy from coordinates (used for intra pairs, CBMC, totals).
PG_HD double pg_pair_energy(const PgDev& P, double ax, double ay, double az, int ta, double bx, double by, double bz,
                            int tb) {
  double dx = pg_bbdist_axis(ax - bx, P.box[0], P.inv_box[0], P.half_box[0], P.pbc[0]);
  double dy = pg_bbdist_axis(ay - by, P.box[1], P.inv_box[1], P.half_box[1], P.pbc[1]);
  double dz = pg_bbdist_axis(az - bz, P.box[2], P.inv_box[2], P.half_box[2], P.pbc[2]);
  double r = sqrt(dx * dx + dy * dy + dz * dz);
  return pg_pair_energy_r(P, r, ta * PG_MAX_TYPES + tb);
}

// Real-space Ewald energy of one pair from its wrapped distance vector.
// Loops the images in the reference's i,j,k order but only over the per-axis
// index ranges that can satisfy r <= real_cutoff (|component| <= rc*(1+1e-9) is
// necessary for r <= rc), then applies the reference's exact predicate.
PG_HD double pg_pair_real_d(const PgDev& P, double dx, double dy, double dz, double qq) {
  double prefactor = P.lB * qq;
  double energy = 0.0;
  if (P.single_image) {
    double r2 = dx * dx + dy * dy + dz * dz;
    if (r2 <= P.rc2_relaxed) {
      double r = sqrt(r2);
      if (r > 0 && r <= P.real_cutoff) energy += prefactor * erfc(P.sqrt_alpha * r) / r;
    }
    return energy;
  }
  const double rc = P.rc_relaxed;
  int i0 = (int)ceil((-rc - dx) * P.inv_ebox[0]), i1 = (int)floor((rc - dx) * P.inv_ebox[0]);
  int j0 = (int)ceil((-rc - dy) * P.inv_ebox[1]), j1 = (int)floor((rc - dy) * P.inv_ebox[1]);
  int k0 = (int)ceil((-rc - dz) * P.inv_ebox[2]), k1 = (int)floor((rc - dz) * P.inv_ebox[2]);
  if (i0 < -P.real_cell[0]) i0 = -P.real_cell[0];
  if (i1 > P.real_cell[0]) i1 = P.real_cell[0];
  if (j0 < -P.real_cell[1]) j0 = -P.real_cell[1];
  if (j1 > P.real_cell[1]) j1 = P.real_cell[1];
  if (k0 < -P.real_cell[2]) k0 = -P.real_cell[2];
  if (k1 > P.real_cell[2]) k1 = P.real_cell[2];
  for (int i = i0; i <= i1; i++) {
    double rx = dx + i * P.ebox[0];
    double rx2 = rx * rx;
    for (int j = j0; j <= j1; j++) {
      double ry = dy + j * P.ebox[1];
      double rxy2 = rx2 + ry * ry;
      if (rxy2 > P.rc2_relaxed) continue;
      for (int k = k0; k <= k1; k++) {
        double rz = dz + k * P.ebox[2];
        double r2 = rxy2 + rz * rz;
        if (r2 > P.rc2_relaxed) continue;
        double r = sqrt(r2);
        if (r > 0 && r <= P.real_cutoff) energy += prefactor * erfc(P.sqrt_alpha * r) / r;
      }
    }
  }
  return energy;
}

PG_HD double pg_pair_real(const PgDev& P, double ax, double ay, double az, double qa, double bx, double by, double bz,
                          double qb) {
  double qq = qa * qb;
  if (qq == 0) return 0.0;
  double dx, dy, dz;
  pg_dist_ewald(P, ax, ay, az, bx, by, bz, dx, dy, dz);
  return pg_pair_real_d(P, dx, dy, dz, qq);
}

// Both short-range terms of one (a, b) pair in one pass.  The LJ separation
// reuses the Ewald wrap on axes where the two boxes coincide (the wrapped
// difference a-b is exactly -(b-a)); the |d| > L/2 fold of BBDist is applied
// only to LJ candidates (it can only trigger through rounding).
PG_HD void pg_pair_both(const PgDev& P, double ax, double ay, double az, double qa, int ta, double bx, double by,
                        double bz, double qb, int tb, int do_lj, double& e_lj, double& e_real) {
  double dx, dy, dz;
  pg_dist_ewald(P, ax, ay, az, bx, by, bz, dx, dy, dz);
  e_lj = 0.0;
  e_real = 0.0;
  if (do_lj) {
    double lx = dx, ly = dy, lz = dz;
    if (!P.same_box[0]) lx = P.pbc[0] ? pg_wrap(ax - bx, P.box[0], P.inv_box[0]) : ax - bx;
    if (!P.same_box[1]) ly = P.pbc[1] ? pg_wrap(ay - by, P.box[1], P.inv_box[1]) : ay - by;
    if (!P.same_box[2]) lz = P.pbc[2] ? pg_wrap(az - bz, P.box[2], P.inv_box[2]) : az - bz;
    double l2 = lx * lx + ly * ly + lz * lz;
    int tp = ta * PG_MAX_TYPES + tb;
    if (P.pair_kind == 2 || l2 <= P.lj_rcut2_relaxed[tp]) {
      if (P.pbc[0] && fabs(lx) > P.half_box[0]) lx = P.box[0] - fabs(lx);
      if (P.pbc[1] && fabs(ly) > P.half_box[1]) ly = P.box[1] - fabs(ly);
      if (P.pbc[2] && fabs(lz) > P.half_box[2]) lz = P.box[2] - fabs(lz);
      double r = sqrt(lx * lx + ly * ly + lz * lz);
      e_lj = pg_pair_energy_r(P, r, tp);
    }
  }
  if (P.use_ewald) {
    double qq = qa * qb;
    if (qq != 0) e_real = pg_pair_real_d(P, dx, dy, dz, qq);
  }
}

// Wall energy of one bead at height z (trial or current), type t.
PG_HD double pg_wall_energy(const PgDev& P, double z, int t) {
  const double Lz = P.box[2];
  const double k213 = 1.25992104989;
  const double c = 2.59807621135;
  if (P.ext_kind == 2) {
    double rad = P.wall_sigma[t];
    return (z <= rad || z + rad >= Lz) ? PG_VLE : 0.0;
  }
  if (P.ext_kind == 3) {
    double rad = P.wall_sigma[t];
    if (z <= rad || z >= Lz - rad) return PG_VLE;
    if (z < P.well_width || z > Lz - P.well_width) return P.well_depth;
    return 0.0;
  }
  double sigma = P.wall_sigma[t], epsilon = P.wall_eps[t];
  if (epsilon == 0) return 0.0;
  if (z <= 0 || z >= Lz) return PG_VLE;
  double energy = 0.0;
  int graft = P.graft[t];
  if (P.wall_cut < 0 && graft == 0) {
    double energy_ref = P.wall_r3ref_e[t];
    if (z < k213 * sigma) {
      double r3 = pg_pow3(sigma / z);
      energy += c * epsilon * (r3 * r3 - r3) - energy_ref;
    }
    if (Lz - z < k213 * sigma) {
      double r3 = pg_pow3(sigma / (Lz - z));
      energy += c * epsilon * (r3 * r3 - r3) - energy_ref;
    }
  } else if (graft == 1 || graft == 2) {
    double R0 = 3 * k213 * sigma;
    double zn = (graft == 1) ? z : (Lz - z);   // distance to the grafting wall
    double zf = (graft == 1) ? (Lz - z) : z;   // distance to the far wall
    double r3 = pg_pow3(sigma / zn);
    energy += c * epsilon * (r3 * r3 - r3);
    double t2 = zn / R0;
    energy += -0.5 * 1.0 * R0 * R0 * log(1 - t2 * t2);
    if (zf < k213 * sigma) {
      double r3b = pg_pow3(sigma / zf);
      energy += c * epsilon * (r3b * r3b - r3b);
    } else {
      energy += P.wall_r3ref_e[t];  // 2.598*eps*(r3^2-r3) at r3 = (1/k213)^3
    }
  } else {
    double energy_ref = P.wall_r3ref_e[t];
    if (z < P.wall_cut) {
      double r3 = pg_pow3(sigma / z);
      energy += c * epsilon * (r3 * r3 - r3) - energy_ref;
    }
    if (Lz - z < P.wall_cut) {
      double r3 = pg_pow3(sigma / (Lz - z));
      energy += c * epsilon * (r3 * r3 - r3) - energy_ref;
    }
  }
  return energy;
}

// One spring bond, no wrapping (BBDist with npbc = 0).
PG_HD double pg_bond_energy(const PgDev& P, double ax, double ay, double az, double bx, double by, double bz) {
  double dx = ax - bx, dy = ay - by, dz = az - bz;
  double r = sqrt(dx * dx + dy * dy + dz * dz);
  return 0.5 * P.bond_k * (r - P.bond_r0) * (r - P.bond_r0);
}

// Wrap one coordinate into the Ewald cell before taking reciprocal-space phases
// (keeps k.r small; see DESIGN.md "truncated kPi").
PG_HD double pg_wrap_pos(double x, double L, double invL, int pbc) { return pbc ? pg_wrap(x, L, invL) : x; }

#endif  // PLUM_B200_PG_MATH_CUH_
