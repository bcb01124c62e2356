// plum_b200 — trial-coordinate arithmetic of Plum's move generators, shared by the device
// kernel (k_propose, pg_propose.cu) and the host (plum_b200/host/mc_propose.h; CPU tests pin
// it to the reference's own trial coordinates bit for bit).
//
// Every operation is a separately rounded IEEE add / mul / div / sqrt in the reference's order
// (it is built for baseline x86-64: no FMA contraction): on the device the round-to-nearest
// intrinsics, which nvcc never fuses; on the host plain operators — compile host translation
// units that include this header with -ffp-contract=off.
#ifndef PLUM_B200_PG_PROPOSE_MATH_H_
#define PLUM_B200_PG_PROPOSE_MATH_H_

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define PP_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define PP_ADD(a, b) __dadd_rn((a), (b))
#define PP_SUB(a, b) __dadd_rn((a), -(b))
#define PP_MUL(a, b) __dmul_rn((a), (b))
#define PP_DIV(a, b) __ddiv_rn((a), (b))
#define PP_SQRT(a) __dsqrt_rn((a))
#else
#define PP_ADD(a, b) ((a) + (b))
#define PP_SUB(a, b) ((a) - (b))
#define PP_MUL(a, b) ((a) * (b))
#define PP_DIV(a, b) ((a) / (b))
#define PP_SQRT(a) sqrt((a))
#endif

// Molecule::BeadTranslate, molecule.cc:116-119: old + vec_len * vec[i].
PP_HD double pp_bead_translate(double old, double s, double v) { return PP_ADD(old, PP_MUL(s, v)); }

// Molecule::COMTranslate, molecule.cc:145-149: old + vec[j].
PP_HD double pp_com_translate(double old, double v) { return PP_ADD(old, v); }

// Molecule::RandomReptation, molecule.cc:301-304: old_end + bond_len * vec[i] / vec_len.
PP_HD double pp_reptation_end(double old_end, double bond_len, double v, double vlen) {
  return PP_ADD(old_end, PP_DIV(PP_MUL(bond_len, v), vlen));
}

// One step of Molecule::Pivot (molecule.cc:170-198 forward, :203-231 backward): bead i (position b)
// is displaced by msr * v and pulled back onto the sphere of radius bond_len around its already
// placed neighbour (position a); m is the translation applied to bead i and the rest of that arm.
PP_HD void pp_pivot_step(const double a[3], const double b[3], double msr, const double v[3], double bond_len,
                         double m[3]) {
  const double x = PP_ADD(b[0], PP_MUL(msr, v[0]));
  const double y = PP_ADD(b[1], PP_MUL(msr, v[1]));
  const double z = PP_ADD(b[2], PP_MUL(msr, v[2]));
  const double dx = PP_SUB(x, a[0]), dy = PP_SUB(y, a[1]), dz = PP_SUB(z, a[2]);
  const double r2 = PP_ADD(PP_ADD(PP_MUL(dx, dx), PP_MUL(dy, dy)), PP_MUL(dz, dz));
  const double norm = PP_DIV(bond_len, PP_SQRT(r2));
  const double cx = PP_ADD(PP_MUL(norm, dx), a[0]);
  const double cy = PP_ADD(PP_MUL(norm, dy), a[1]);
  const double cz = PP_ADD(PP_MUL(norm, dz), a[2]);
  m[0] = PP_SUB(cx, b[0]);
  m[1] = PP_SUB(cy, b[1]);
  m[2] = PP_SUB(cz, b[2]);
}

// Molecule::Crankshaft, molecule.cc:242-250: the rotation about the axis bead `first` -> bead `last` by an angle
// whose sine and cosine are s and c.  Vector3d::normalize (x / sqrt(x.x), skipped for a zero vector) and
// AngleAxisd::toRotationMatrix in Eigen's operation order (plum_b200/host/eigen_standin):
// sin_axis = s * axis, cos1_axis = (1 - c) * axis, off-diagonals tmp -/+ sin_axis, diagonal cos1_axis * axis + c.
PP_HD void pp_crank_matrix(const double pf[3], const double pl[3], double s, double c, double rot[9]) {
  double ax[3] = {PP_SUB(pl[0], pf[0]), PP_SUB(pl[1], pf[1]), PP_SUB(pl[2], pf[2])};
  const double n2 = PP_ADD(PP_ADD(PP_MUL(ax[0], ax[0]), PP_MUL(ax[1], ax[1])), PP_MUL(ax[2], ax[2]));
  if (n2 > 0) {
    const double n = PP_SQRT(n2);
    ax[0] = PP_DIV(ax[0], n); ax[1] = PP_DIV(ax[1], n); ax[2] = PP_DIV(ax[2], n);
  }
  const double omc = PP_SUB(1.0, c);
  const double sa[3] = {PP_MUL(s, ax[0]), PP_MUL(s, ax[1]), PP_MUL(s, ax[2])};
  const double ca[3] = {PP_MUL(omc, ax[0]), PP_MUL(omc, ax[1]), PP_MUL(omc, ax[2])};
  double tmp = PP_MUL(ca[0], ax[1]);
  rot[1] = PP_SUB(tmp, sa[2]);
  rot[3] = PP_ADD(tmp, sa[2]);
  tmp = PP_MUL(ca[0], ax[2]);
  rot[2] = PP_ADD(tmp, sa[1]);
  rot[6] = PP_SUB(tmp, sa[1]);
  tmp = PP_MUL(ca[1], ax[2]);
  rot[5] = PP_SUB(tmp, sa[0]);
  rot[7] = PP_ADD(tmp, sa[0]);
  rot[0] = PP_ADD(PP_MUL(ca[0], ax[0]), c);
  rot[4] = PP_ADD(PP_MUL(ca[1], ax[1]), c);
  rot[8] = PP_ADD(PP_MUL(ca[2], ax[2]), c);
}

// molecule.cc:254-262: v = pos - first; v = rot * v (row times vector, summed left to right); trial = v + first.
PP_HD void pp_crank_apply(const double rot[9], const double pf[3], const double pos[3], double out[3]) {
  const double v[3] = {PP_SUB(pos[0], pf[0]), PP_SUB(pos[1], pf[1]), PP_SUB(pos[2], pf[2])};
  for (int i = 0; i < 3; i++) {
    const double r = PP_ADD(PP_ADD(PP_MUL(rot[3 * i], v[0]), PP_MUL(rot[3 * i + 1], v[1])), PP_MUL(rot[3 * i + 2], v[2]));
    out[i] = PP_ADD(r, pf[i]);
  }
}


#endif
