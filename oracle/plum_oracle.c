/* TEST INFRASTRUCTURE ONLY — CPU restatement of Plum's per-trial-move energy path.
 *
 * This file is the oracle the CUDA path is checked against.  It is never linked
 * into, imported by or executed from the product path (plum_b200/, include/);
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it.
 *
 * It restates, in plain C and without the reference's N x N std::map caches, the
 * arithmetic of nuwapi/Plum (file:line relative to /root/reference):
 *   minimum-image distance   src/molecules/bead.cc:162-176 (BBDist)
 *   distance vector          src/utilities/misc.cc:43-56 (GetDistVector)
 *   constants                src/utilities/constants.h:5-62
 *   LJ / WCA pair            src/force_field/potential_truncated_lj.cc:49-85
 *   hard sphere              src/force_field/potential_hard_sphere.cc:38-49
 *   pair dE loop             src/force_field/potential_pair.cc:152-201
 *   Ewald set-up + tables    src/force_field/potential_ewald_coul.cc:29-132
 *   Ewald real / recip pair  src/force_field/potential_ewald_coul.cc:134-164,198-225
 *   self, dipole             src/force_field/potential_ewald_coul.cc:253-257,415-530
 *   Ewald init / dE / final  src/force_field/potential_ewald.cc:176-230,429-535,537-624
 *   GC energy bookkeeping    src/force_field/potential_ewald.cc:233-427,626-707,
 *                            src/force_field/potential_pair.cc:102-150,249-294
 *   spring bond              src/force_field/potential_spring.cc:22-48,63-80
 *   walls                    src/force_field/potential_external.cc:100-117,
 *                            potential_truncated_lj_wall.cc:46-134,
 *                            potential_hard_wall.cc:36-47, potential_well_wall.cc:41-55
 *   orchestration            src/force_field/force_field.cc:407-451
 *   CBMC trial energy        src/force_field/cbmc.cc:5-151
 *
 * "Old" pair energies, which the reference reads from its caches, are recomputed
 * from the current coordinates (they are a pure function of them).
 * Loop orders follow the reference so that sums agree to rounding.
 *
 * Parity pin: tests/test_oracle_vs_reference.py replays traces written by the
 * real reference (oracle/_ref/plum_ref, built by oracle/build_ref.py) and the
 * committed fixtures under tests/golden/ through this code.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/plum_b200.h" /* parameter block layout only */

/* src/utilities/constants.h */
static const double kPi = 3.14159265359;
static const double kVeryLargeEnergy = 1E+8;
static const double kEwaldCutoff = 1E-7;
static const int kDiCorrection = 3;
static const double k213 = 1.25992104989;
static const double k216 = 1.12246204831;

typedef struct po_system {
  /* parameters */
  double box[3];
  int npbc;
  double beta;
  int n_types;
  int pair_kind;
  double lj_cutoff;
  double *lj_sigma, *lj_epsilon, *hs_radius;
  int use_ewald, dipole_correction;
  double lB, alpha;
  int bond_kind;
  double bond_k, bond_r0;
  int ext_kind;
  double wall_cut;
  double *wall_sigma, *wall_epsilon;
  int *graft_kind;
  double well_width, well_depth;
  /* Ewald derived */
  double ebox[3], box_vol, real_cutoff, repl_cutoff;
  int real_cell[3], repl_cell[3], repl_ceto[3];
  double *kx, *ky, *kz, *k2, *ek2;
  int n_k;
  /* beads: current and trial coordinates */
  int n, cap, n_mol, mol_cap;
  double *cur, *tri; /* [n][3] */
  double *q;
  int *type;
  int *mol_first; /* [n_mol+1] */
  /* running totals (E_tot of each potential) + dipole lag state */
  double E_pair, E_ewald, E_bond, E_ext;
  double current_dipl_E, trial_dipl_E;
  /* pending trial */
  int pending_mol;
  unsigned char *moved; /* [n] flags for pending trial */
  double dE_pair, dE_ewald, dE_bond, dE_ext;
  int repl_mode; /* 0 = pairwise cube scan (reference form), 1 = S(k) form */
  int timing_new_only; /* bench only: skip the recomputation of "old" pair energies, which the reference
                          reads from its std::map caches — the work left is exactly the reference's */
} po_system;

static double *dup_d(const double *p, int n) {
  double *r = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  if (p) memcpy(r, p, sizeof(double) * (size_t)n);
  return r;
}

/* ---------------------------------------------------------------- set-up -- */

/* src/force_field/potential_ewald_coul.cc:29-132 */
static void ewald_setup(po_system *s) {
  for (int i = 0; i < 3; i++) s->ebox[i] = s->box[i];
  if (s->dipole_correction) {
    double min_padding = 150;
    if (s->ebox[2] * kDiCorrection > min_padding)
      s->ebox[2] = s->ebox[2] * kDiCorrection;
    else
      s->ebox[2] += min_padding;
  }
  s->box_vol = s->ebox[0] * s->ebox[1] * s->ebox[2];
  s->real_cutoff = 1;
  while (0.5 * s->lB * 1 * 1 * erfc(sqrt(s->alpha) * s->real_cutoff) / s->real_cutoff > kEwaldCutoff)
    s->real_cutoff += 1;
  for (int i = 0; i < 3; i++) s->real_cell[i] = (int)ceil(s->real_cutoff / s->ebox[i]);
  s->repl_cutoff = s->alpha;
  double min_box_vol = pow(s->ebox[0], 3);
  while (s->lB * 1 * 1 / (2 * kPi * min_box_vol) * (4 * kPi * kPi) * exp(-s->repl_cutoff / (4 * s->alpha)) /
             s->repl_cutoff > kEwaldCutoff)
    s->repl_cutoff += s->alpha;
  for (int i = 0; i < 3; i++) s->repl_cell[i] = (int)ceil(sqrt(s->repl_cutoff) * s->ebox[i] / (2 * kPi));
  for (int i = 0; i < 3; i++) s->repl_ceto[i] = 2 * s->repl_cell[i] + 1;
  s->kx = (double *)malloc(sizeof(double) * s->repl_ceto[0]);
  s->ky = (double *)malloc(sizeof(double) * s->repl_ceto[1]);
  s->kz = (double *)malloc(sizeof(double) * s->repl_ceto[2]);
  size_t cube = (size_t)s->repl_ceto[0] * s->repl_ceto[1] * s->repl_ceto[2];
  s->k2 = (double *)malloc(sizeof(double) * cube);
  s->ek2 = (double *)malloc(sizeof(double) * cube);
  for (int lx = -s->repl_cell[0]; lx <= s->repl_cell[0]; lx++) s->kx[lx + s->repl_cell[0]] = lx * 2 * kPi / s->ebox[0];
  for (int ly = -s->repl_cell[1]; ly <= s->repl_cell[1]; ly++) s->ky[ly + s->repl_cell[1]] = ly * 2 * kPi / s->ebox[1];
  for (int lz = -s->repl_cell[2]; lz <= s->repl_cell[2]; lz++) s->kz[lz + s->repl_cell[2]] = lz * 2 * kPi / s->ebox[2];
  s->n_k = 0;
  for (int ix = 0; ix < s->repl_ceto[0]; ix++)
    for (int iy = 0; iy < s->repl_ceto[1]; iy++)
      for (int iz = 0; iz < s->repl_ceto[2]; iz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * ix + (size_t)s->repl_ceto[2] * iy + iz;
        s->k2[idx] = s->kx[ix] * s->kx[ix] + s->ky[iy] * s->ky[iy] + s->kz[iz] * s->kz[iz];
        s->ek2[idx] = exp(-s->k2[idx] / (4 * s->alpha)) / s->k2[idx];
        if (s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff) s->n_k++;
      }
}

po_system *po_create(const pg_params *p) {
  po_system *s = (po_system *)calloc(1, sizeof(po_system));
  memcpy(s->box, p->box, sizeof(s->box));
  s->npbc = p->npbc;
  s->beta = p->beta;
  s->n_types = p->n_types;
  s->pair_kind = p->pair_kind;
  s->lj_cutoff = p->lj_cutoff;
  s->lj_sigma = dup_d(p->lj_sigma, p->n_types);
  s->lj_epsilon = dup_d(p->lj_epsilon, p->n_types);
  s->hs_radius = dup_d(p->hs_radius, p->n_types);
  s->use_ewald = p->use_ewald;
  s->dipole_correction = p->dipole_correction;
  s->lB = p->lB;
  s->alpha = p->alpha;
  s->bond_kind = p->bond_kind;
  s->bond_k = p->bond_k;
  s->bond_r0 = p->bond_r0;
  s->ext_kind = p->ext_kind;
  s->wall_cut = p->wall_cut;
  s->wall_sigma = dup_d(p->wall_sigma, p->n_types);
  s->wall_epsilon = dup_d(p->wall_epsilon, p->n_types);
  s->graft_kind = (int *)calloc((size_t)(p->n_types > 0 ? p->n_types : 1), sizeof(int));
  if (p->graft_kind) memcpy(s->graft_kind, p->graft_kind, sizeof(int) * (size_t)p->n_types);
  s->well_width = p->well_width;
  s->well_depth = p->well_depth;
  s->pending_mol = -1;
  if (s->use_ewald) ewald_setup(s);
  return s;
}

void po_destroy(po_system *s) {
  if (!s) return;
  free(s->lj_sigma); free(s->lj_epsilon); free(s->hs_radius);
  free(s->wall_sigma); free(s->wall_epsilon); free(s->graft_kind);
  free(s->kx); free(s->ky); free(s->kz); free(s->k2); free(s->ek2);
  free(s->cur); free(s->tri); free(s->q); free(s->type); free(s->mol_first); free(s->moved);
  free(s);
}

void po_set_repl_mode(po_system *s, int mode) { s->repl_mode = mode; }
/* TIMING ONLY (bench.py cpu_baseline / --impl reference): with flag = 1 po_delta_e evaluates the
 * new-configuration energies only, like the reference (potential_pair.cc:186-192,
 * potential_ewald.cc:488-500 read the old ones from maps); the returned dE is then meaningless. */
void po_set_timing_new_only(po_system *s, int flag) { s->timing_new_only = flag; }

int po_get_ewald_info(const po_system *s, pg_ewald_info *o) {
  memset(o, 0, sizeof(*o));
  if (!s->use_ewald) return 0;
  for (int i = 0; i < 3; i++) {
    o->ewald_box[i] = s->ebox[i];
    o->real_cell[i] = s->real_cell[i];
    o->repl_cell[i] = s->repl_cell[i];
  }
  o->box_vol = s->box_vol;
  o->real_cutoff = s->real_cutoff;
  o->repl_cutoff = s->repl_cutoff;
  o->n_k = s->n_k;
  o->n_k_half = s->n_k / 2;
  return 0;
}

static void reserve(po_system *s, int n, int n_mol) {
  if (n > s->cap || s->cur == NULL) {
    int c = n * 2 + 64;
    s->cur = (double *)realloc(s->cur, sizeof(double) * 3 * (size_t)c);
    s->tri = (double *)realloc(s->tri, sizeof(double) * 3 * (size_t)c);
    s->q = (double *)realloc(s->q, sizeof(double) * (size_t)c);
    s->type = (int *)realloc(s->type, sizeof(int) * (size_t)c);
    s->moved = (unsigned char *)realloc(s->moved, (size_t)c);
    s->cap = c;
  }
  if (n_mol + 1 > s->mol_cap) {
    int c = n_mol * 2 + 64;
    s->mol_first = (int *)realloc(s->mol_first, sizeof(int) * (size_t)c);
    s->mol_cap = c;
  }
}

int po_upload_system(po_system *s, int n, const double *xyz, const double *q, const int32_t *type, int n_mol,
                     const int32_t *mol_first) {
  reserve(s, n, n_mol);
  s->n = n;
  s->n_mol = n_mol;
  memcpy(s->cur, xyz, sizeof(double) * 3 * (size_t)n);
  memcpy(s->tri, xyz, sizeof(double) * 3 * (size_t)n);
  memcpy(s->q, q, sizeof(double) * (size_t)n);
  for (int i = 0; i < n; i++) s->type[i] = type[i];
  for (int i = 0; i <= n_mol; i++) s->mol_first[i] = mol_first[i];
  if (n > 0) memset(s->moved, 0, (size_t)n);
  s->pending_mol = -1;
  return 0;
}

int po_num_beads(const po_system *s) { return s->n; }
int po_download_positions(const po_system *s, double *xyz) {
  memcpy(xyz, s->cur, sizeof(double) * 3 * (size_t)s->n);
  return 0;
}

/* ------------------------------------------------------ pair primitives -- */

/* src/molecules/bead.cc:162-176 */
static double bbdist(const po_system *s, const double *a, const double *b, int npbc) {
  double dist = 0;
  for (int i = 0; i < 3; i++) {
    double d = a[i] - b[i];
    if (i < npbc) {
      d -= s->box[i] * round(d / s->box[i]);
      if (fabs(d) > 0.5 * s->box[i]) d = s->box[i] - fabs(d);
    }
    dist += d * d;
  }
  return sqrt(dist);
}

/* src/force_field/potential_truncated_lj.cc:49-85, potential_hard_sphere.cc:38-49 */
static double pair_energy(const po_system *s, const double *a, int ta, const double *b, int tb) {
  double r = bbdist(s, a, b, s->npbc);
  if (s->pair_kind == PG_PAIR_HARD_SPHERE) {
    double allowed = s->hs_radius[ta] + s->hs_radius[tb];
    return (r <= allowed) ? kVeryLargeEnergy : 0.0;
  }
  double energy = 0;
  double sigma = (s->lj_sigma[ta] + s->lj_sigma[tb]) / 2;
  double epsilon = sqrt(s->lj_epsilon[ta] * s->lj_epsilon[tb]);
  double r6 = 0;
  if (r <= 0) {
    energy = kVeryLargeEnergy;
  } else if (s->lj_cutoff < 0) {
    double r6_ref = pow((1.0 / k216), 6);
    double energy_ref = 4 * epsilon * (r6_ref * r6_ref - r6_ref);
    if (r < k216 * sigma) {
      r6 = pow((sigma / r), 6);
      energy = 4 * epsilon * (r6 * r6 - r6) - energy_ref;
    }
  } else if (r < s->lj_cutoff) {
    double r6_ref = pow((sigma / s->lj_cutoff), 6);
    double energy_ref = 4 * epsilon * (r6_ref * r6_ref - r6_ref);
    r6 = pow((sigma / r), 6);
    energy = 4 * epsilon * (r6 * r6 - r6) - energy_ref;
  }
  return energy;
}

/* src/utilities/misc.cc:43-56 with the Ewald (padded) box */
static void dist_vector(const po_system *s, const double *b1, const double *b2, double d[3]) {
  for (int i = 0; i < 3; i++) {
    double di = b2[i] - b1[i];
    if (i < s->npbc) di -= s->ebox[i] * round(di / s->ebox[i]);
    d[i] = di;
  }
}

/* src/force_field/potential_ewald_coul.cc:134-164 */
static double pair_real(const po_system *s, const double *b1, double q1, const double *b2, double q2) {
  double energy = 0;
  if (q1 * q2 == 0) return energy;
  double dist[3];
  dist_vector(s, b1, b2, dist);
  double prefactor = s->lB * q1 * q2;
  for (int i = -s->real_cell[0]; i <= s->real_cell[0]; i++)
    for (int j = -s->real_cell[1]; j <= s->real_cell[1]; j++)
      for (int k = -s->real_cell[2]; k <= s->real_cell[2]; k++) {
        double rx = dist[0] + i * s->ebox[0];
        double ry = dist[1] + j * s->ebox[1];
        double rz = dist[2] + k * s->ebox[2];
        double r = sqrt(rx * rx + ry * ry + rz * rz);
        if (r > 0 && r <= s->real_cutoff) energy += prefactor * erfc(sqrt(s->alpha) * r) / r;
      }
  return energy;
}

/* src/force_field/potential_ewald_coul.cc:198-225 */
static double pair_repl(const po_system *s, const double *b1, double q1, const double *b2, double q2) {
  double energy = 0;
  if (q1 * q2 == 0) return energy;
  double r[3];
  dist_vector(s, b1, b2, r);
  double prefactor = s->lB * q1 * q2 / (kPi * s->box_vol) * (4 * kPi * kPi);
  for (int lx = 0; lx < s->repl_ceto[0]; lx++)
    for (int ly = 0; ly < s->repl_ceto[1]; ly++)
      for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
        if (s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)
          energy += prefactor * s->ek2[idx] * cos(s->kx[lx] * r[0] + s->ky[ly] * r[1] + s->kz[lz] * r[2]);
      }
  return energy;
}

/* src/force_field/potential_ewald_coul.cc:253-257 */
static double self_energy(const po_system *s, double q) { return -s->lB * sqrt(s->alpha / kPi) * q * q; }

/* src/force_field/potential_truncated_lj_wall.cc:46-134 and the hard / well walls */
static double wall_energy(const po_system *s, const double *pos, int t) {
  double z = pos[2];
  double Lz = s->box[2];
  if (s->ext_kind == PG_EXT_HARD_WALL) {
    double rad = s->wall_sigma[t];
    return (z <= rad || z + rad >= Lz) ? kVeryLargeEnergy : 0.0;
  }
  if (s->ext_kind == PG_EXT_WELL_WALL) {
    double rad = s->wall_sigma[t];
    if (z <= rad || z >= Lz - rad) return kVeryLargeEnergy;
    if (z < s->well_width || z > Lz - s->well_width) return s->well_depth;
    return 0.0;
  }
  double energy = 0;
  double sigma = s->wall_sigma[t];
  double epsilon = s->wall_epsilon[t];
  double R0 = 3 * k213 * sigma;
  double K = 1;
  int graft = s->graft_kind[t];
  if (epsilon == 0) {
    energy = 0;
  } else if (z <= 0 || z >= Lz) {
    return kVeryLargeEnergy;
  } else {
    if (s->wall_cut < 0 && graft == PG_GRAFT_NONE) {
      double r3_ref = pow((1.0 / k213), 3);
      double energy_ref = 2.59807621135 * epsilon * (r3_ref * r3_ref - r3_ref);
      if (z < k213 * sigma) {
        double r3 = pow((sigma / z), 3);
        energy += 2.59807621135 * epsilon * (r3 * r3 - r3) - energy_ref;
      }
      if (Lz - z < k213 * sigma) {
        double r3 = pow((sigma / (Lz - z)), 3);
        energy += 2.59807621135 * epsilon * (r3 * r3 - r3) - energy_ref;
      }
    } else if (graft == PG_GRAFT_LEFT) {
      double r3 = pow((sigma / z), 3);
      energy += 2.59807621135 * epsilon * (r3 * r3 - r3);
      energy += -0.5 * K * R0 * R0 * log(1 - pow(z / R0, 2));
      if (Lz - z < k213 * sigma) {
        double r3b = pow((sigma / (Lz - z)), 3);
        energy += 2.59807621135 * epsilon * (r3b * r3b - r3b);
      } else {
        double r3b = pow((1.0 / k213), 3);
        energy += 2.59807621135 * epsilon * (r3b * r3b - r3b);
      }
    } else if (graft == PG_GRAFT_RIGHT) {
      double r3 = pow((sigma / (Lz - z)), 3);
      energy += 2.59807621135 * epsilon * (r3 * r3 - r3);
      energy += -0.5 * K * R0 * R0 * log(1 - pow((Lz - z) / R0, 2));
      if (z < k213 * sigma) {
        double r3b = pow((sigma / z), 3);
        energy += 2.59807621135 * epsilon * (r3b * r3b - r3b);
      } else {
        double r3b = pow((1.0 / k213), 3);
        energy += 2.59807621135 * epsilon * (r3b * r3b - r3b);
      }
    } else {
      double r3_ref = pow((1.0 / s->wall_cut), 3);
      double energy_ref = 2.59807621135 * epsilon * (r3_ref * r3_ref - r3_ref);
      if (z < s->wall_cut) {
        double r3 = pow((sigma / z), 3);
        energy += 2.59807621135 * epsilon * (r3 * r3 - r3) - energy_ref;
      }
      if (Lz - z < s->wall_cut) {
        double r3 = pow((sigma / (Lz - z)), 3);
        energy += 2.59807621135 * epsilon * (r3 * r3 - r3) - energy_ref;
      }
    }
  }
  return energy;
}

/* src/force_field/potential_spring.cc:22-48 — consecutive beads, no wrapping */
static double molecule_bond_energy(const po_system *s, const double *pos, int first, int last) {
  double e = 0;
  for (int i = first; i < last - 1; i++) {
    double r = bbdist(s, pos + 3 * i, pos + 3 * (i + 1), 0);
    e += 0.5 * s->bond_k * (r - s->bond_r0) * (r - s->bond_r0);
  }
  return e;
}

/* src/force_field/potential_ewald_coul.cc:415-428, current coordinates, skip range of beads */
static double mz_current(const po_system *s, int skip_first, int skip_last) {
  double Mz = 0;
  for (int i = 0; i < s->n; i++) {
    if (i >= skip_first && i < skip_last) continue;
    Mz += s->q[i] * s->cur[3 * i + 2];
  }
  return Mz;
}
static double dipole_from_mz(const po_system *s, double Mz) { return s->lB * 2 * kPi / (s->box_vol / 1.0) * Mz * Mz; }

double po_mz_current(const po_system *s) { return mz_current(s, -1, -1); }

/* ------------------------------------------------ structure-factor form -- */
/* E_rec = (2 kPi lB / V) sum_{k pass} ek2(k) |S(k)|^2, S(k) = sum_i q_i e^{i k.r_i}
 * — algebraically the pairwise sum of potential_ewald_coul.cc:198-225 plus the
 * 0.5 self-image terms of potential_ewald.cc:445-448.
 * Coordinates are wrapped into the Ewald cell first, as the device does. */
static void wrap_pos(const po_system *s, const double *p, double w[3]) {
  for (int i = 0; i < 3; i++) {
    w[i] = p[i];
    if (i < s->npbc) w[i] -= s->ebox[i] * round(w[i] / s->ebox[i]);
  }
}

/* Full-cube S(k): out[2*idx], out[2*idx+1] for every cube entry (zeros when filtered out). */
int po_sk_full(const po_system *s, const double *pos, double *out) {
  size_t cube = (size_t)s->repl_ceto[0] * s->repl_ceto[1] * s->repl_ceto[2];
  memset(out, 0, sizeof(double) * 2 * cube);
  for (int lx = 0; lx < s->repl_ceto[0]; lx++)
    for (int ly = 0; ly < s->repl_ceto[1]; ly++)
      for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
        if (!(s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)) continue;
        double re = 0, im = 0;
        for (int i = 0; i < s->n; i++) {
          if (s->q[i] == 0) continue;
          double w[3];
          wrap_pos(s, pos + 3 * i, w);
          double ph = s->kx[lx] * w[0] + s->ky[ly] * w[1] + s->kz[lz] * w[2];
          re += s->q[i] * cos(ph);
          im += s->q[i] * sin(ph);
        }
        out[2 * idx] = re;
        out[2 * idx + 1] = im;
      }
  return 0;
}

/* Half-space list in cube order: entries whose first non-zero index of (lx,ly,lz) is
 * positive.  kvec_out [n][3] (integer triplets as doubles), returns count. */
int po_k_half_list(const po_system *s, int32_t *l_out, double *ek2_out, int max_n) {
  int n = 0;
  for (int lx = 0; lx < s->repl_ceto[0]; lx++)
    for (int ly = 0; ly < s->repl_ceto[1]; ly++)
      for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
        if (!(s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)) continue;
        int ax = lx - s->repl_cell[0], ay = ly - s->repl_cell[1], az = lz - s->repl_cell[2];
        int pos = (ax > 0) || (ax == 0 && ay > 0) || (ax == 0 && ay == 0 && az > 0);
        if (!pos) continue;
        if (n < max_n) {
          if (l_out) { l_out[3 * n] = ax; l_out[3 * n + 1] = ay; l_out[3 * n + 2] = az; }
          if (ek2_out) ek2_out[n] = s->ek2[idx];
        }
        n++;
      }
  return n;
}

/* S(k) on the half-space list, (re,im) interleaved, for positions pos. */
int po_sk_half(const po_system *s, const double *pos, double *out, int max_n) {
  int n = 0;
  for (int lx = 0; lx < s->repl_ceto[0]; lx++)
    for (int ly = 0; ly < s->repl_ceto[1]; ly++)
      for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
        if (!(s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)) continue;
        int ax = lx - s->repl_cell[0], ay = ly - s->repl_cell[1], az = lz - s->repl_cell[2];
        int posi = (ax > 0) || (ax == 0 && ay > 0) || (ax == 0 && ay == 0 && az > 0);
        if (!posi) continue;
        if (n < max_n) {
          double re = 0, im = 0;
          for (int i = 0; i < s->n; i++) {
            if (s->q[i] == 0) continue;
            double w[3];
            wrap_pos(s, pos + 3 * i, w);
            double ph = s->kx[lx] * w[0] + s->ky[ly] * w[1] + s->kz[lz] * w[2];
            re += s->q[i] * cos(ph);
            im += s->q[i] * sin(ph);
          }
          out[2 * n] = re;
          out[2 * n + 1] = im;
        }
        n++;
      }
  return n;
}

static double recip_energy_sk(const po_system *s, const double *pos) {
  size_t cube = (size_t)s->repl_ceto[0] * s->repl_ceto[1] * s->repl_ceto[2];
  double *sk = (double *)malloc(sizeof(double) * 2 * cube);
  po_sk_full(s, pos, sk);
  double e = 0;
  for (size_t idx = 0; idx < cube; idx++)
    if (s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)
      e += s->ek2[idx] * (sk[2 * idx] * sk[2 * idx] + sk[2 * idx + 1] * sk[2 * idx + 1]);
  free(sk);
  return 2 * kPi * s->lB / s->box_vol * e;
}

/* ------------------------------------------------------ total energies -- */

/* out: pair, ewald, bond, ext, real, recip, self, dipole
 * PotentialPair::EnergyInitialization potential_pair.cc:55-100,
 * PotentialEwald::EnergyInitialization potential_ewald.cc:176-230,
 * PotentialBond potential_bond.cc:16-27, PotentialExternal potential_external.cc:48-61. */
int po_compute_totals(const po_system *s, pg_totals *o) {
  memset(o, 0, sizeof(*o));
  const double *x = s->cur;
  if (s->pair_kind != PG_PAIR_NONE) {
    double E = 0;
    for (int m = 0; m < s->n_mol; m++)
      for (int j = s->mol_first[m]; j < s->mol_first[m + 1] - 1; j++)
        for (int k = j + 1; k < s->mol_first[m + 1]; k++) {
          if (s->pair_kind == PG_PAIR_HARD_SPHERE && k == j + 1) continue;
          E += pair_energy(s, x + 3 * j, s->type[j], x + 3 * k, s->type[k]);
        }
    for (int m = 0; m < s->n_mol - 1; m++)
      for (int m2 = m + 1; m2 < s->n_mol; m2++)
        for (int k = s->mol_first[m]; k < s->mol_first[m + 1]; k++)
          for (int l = s->mol_first[m2]; l < s->mol_first[m2 + 1]; l++)
            E += pair_energy(s, x + 3 * k, s->type[k], x + 3 * l, s->type[l]);
    o->pair = E;
  }
  if (s->use_ewald) {
    double Er = 0, Ek = 0, Et = 0, Es = 0;
    /* bead order == molecule order, so the reference's 4-deep loop is i<=j over beads */
    for (int i = 0; i < s->n; i++)
      for (int j = i; j < s->n; j++) {
        double er = pair_real(s, x + 3 * i, s->q[i], x + 3 * j, s->q[j]);
        double ek = (s->repl_mode == 0) ? pair_repl(s, x + 3 * i, s->q[i], x + 3 * j, s->q[j]) : 0.0;
        if (i == j) { er *= 0.5; ek *= 0.5; }
        Er += er; Ek += ek; Et += (er + ek);
      }
    if (s->repl_mode != 0) { Ek = recip_energy_sk(s, x); Et += Ek; }
    for (int i = 0; i < s->n; i++) { double e = self_energy(s, s->q[i]); Es += e; Et += e; }
    o->real = Er; o->recip = Ek; o->self = Es;
    if (s->dipole_correction) { o->dipole = dipole_from_mz(s, mz_current(s, -1, -1)); Et += o->dipole; }
    o->ewald = Et;
  }
  if (s->bond_kind != PG_BOND_NONE) {
    double E = 0;
    for (int m = 0; m < s->n_mol; m++) E += molecule_bond_energy(s, x, s->mol_first[m], s->mol_first[m + 1]);
    o->bond = E;
  }
  if (s->ext_kind != PG_EXT_NONE) {
    double E = 0;
    for (int i = 0; i < s->n; i++) E += wall_energy(s, x + 3 * i, s->type[i]);
    o->ext = E;
  }
  return 0;
}

int po_init_energy(po_system *s, pg_totals *o) {
  pg_totals t;
  po_compute_totals(s, &t);
  s->E_pair = t.pair; s->E_ewald = t.ewald; s->E_bond = t.bond; s->E_ext = t.ext;
  s->current_dipl_E = t.dipole; s->trial_dipl_E = t.dipole;
  if (o) *o = t;
  return 0;
}

int po_get_totals(const po_system *s, pg_totals *o) {
  memset(o, 0, sizeof(*o));
  o->pair = s->E_pair; o->ewald = s->E_ewald; o->bond = s->E_bond; o->ext = s->E_ext;
  o->dipole = s->current_dipl_E;
  return 0;
}

/* ------------------------------------------------------------ per move -- */

/* ForceField::EnergyDifference force_field.cc:407-434 over the four potentials. */
int po_delta_e(po_system *s, int mol, const double *trial_xyz, const uint8_t *moved, pg_delta *o) {
  memset(o, 0, sizeof(*o));
  int f = s->mol_first[mol], l = s->mol_first[mol + 1], len = l - f;
  for (int i = 0; i < len; i++) {
    s->moved[f + i] = moved[i] ? 1 : 0;
    for (int a = 0; a < 3; a++) s->tri[3 * (f + i) + a] = trial_xyz[3 * i + a];
  }
  s->pending_mol = mol;
  const double *C = s->cur, *T = s->tri;
  s->dE_pair = s->dE_ewald = s->dE_bond = s->dE_ext = 0;
  o->mz_current = mz_current(s, -1, -1);
  double dE = 0;
  /* --- pair: potential_pair.cc:152-201 */
  if (s->pair_kind != PG_PAIR_NONE) {
    double d = 0;
    for (int i = f; i < l - 1; i++)
      for (int j = i + 1; j < l; j++)
        if (s->moved[i] || s->moved[j]) {
          double new_e, old_e;
          if (s->pair_kind == PG_PAIR_HARD_SPHERE && j == i + 1) {
            new_e = 0; old_e = 0;
          } else {
            new_e = pair_energy(s, T + 3 * i, s->type[i], T + 3 * j, s->type[j]);
            old_e = s->timing_new_only ? 0.0 : pair_energy(s, C + 3 * i, s->type[i], C + 3 * j, s->type[j]);
          }
          if (new_e >= kVeryLargeEnergy) o->n_overlap++;
          d += (new_e - old_e);
        }
    for (int j = 0; j < s->n; j++) {
      if (j >= f && j < l) continue;
      for (int k = f; k < l; k++)
        if (s->moved[k]) {
          double new_e = pair_energy(s, T + 3 * k, s->type[k], C + 3 * j, s->type[j]);
          double old_e = s->timing_new_only ? 0.0 : pair_energy(s, C + 3 * k, s->type[k], C + 3 * j, s->type[j]);
          if (new_e >= kVeryLargeEnergy) o->n_overlap++;
          d += (new_e - old_e);
        }
    }
    s->dE_pair = d;
    o->pair = d;
    dE += d;
    if (dE >= kVeryLargeEnergy) { o->dE = dE; o->stage = 1; return 0; }
  }
  /* --- external: potential_external.cc:100-117 */
  if (s->ext_kind != PG_EXT_NONE) {
    double d = 0;
    for (int i = f; i < l; i++)
      if (s->moved[i]) {
        double eNew = wall_energy(s, T + 3 * i, s->type[i]);
        if (eNew >= kVeryLargeEnergy) { d = kVeryLargeEnergy; break; }
        d += eNew - wall_energy(s, C + 3 * i, s->type[i]);
      }
    s->dE_ext = d;
    o->ext = d;
    dE += d;
    if (dE >= kVeryLargeEnergy) { o->dE = dE; o->stage = 2; return 0; }
  }
  /* --- Ewald: potential_ewald.cc:429-535 */
  if (s->use_ewald) {
    double d = 0, dr = 0, dk = 0;
    for (int i = f; i < l; i++)
      for (int j = i; j < l; j++)
        if (s->moved[i] || s->moved[j]) {
          double nr = pair_real(s, T + 3 * i, s->q[i], T + 3 * j, s->q[j]);
          double orr = s->timing_new_only ? 0.0 : pair_real(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
          double nk = 0, ok = 0;
          if (s->repl_mode == 0) {
            nk = pair_repl(s, T + 3 * i, s->q[i], T + 3 * j, s->q[j]);
            ok = s->timing_new_only ? 0.0 : pair_repl(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
          }
          if (j == i) { nr *= 0.5; nk *= 0.5; orr *= 0.5; ok *= 0.5; }
          d += (nr + nk - (orr + ok));
          dr += nr - orr; dk += nk - ok;
        }
    for (int j = 0; j < s->n; j++) {
      if (j >= f && j < l) continue;
      for (int k = f; k < l; k++)
        if (s->moved[k]) {
          double nr = pair_real(s, T + 3 * k, s->q[k], C + 3 * j, s->q[j]);
          double orr = s->timing_new_only ? 0.0 : pair_real(s, C + 3 * k, s->q[k], C + 3 * j, s->q[j]);
          double nk = 0, ok = 0;
          if (s->repl_mode == 0) {
            nk = pair_repl(s, T + 3 * k, s->q[k], C + 3 * j, s->q[j]);
            ok = s->timing_new_only ? 0.0 : pair_repl(s, C + 3 * k, s->q[k], C + 3 * j, s->q[j]);
          }
          d += (nr + nk - (orr + ok));
          dr += nr - orr; dk += nk - ok;
        }
    }
    if (s->repl_mode != 0) {
      /* S(k) form: E_rec(trial configuration) - E_rec(current configuration) via
       * 2 Re(conj(S) dS) + |dS|^2, never |S_new|^2 - |S_old|^2 */
      size_t cube = (size_t)s->repl_ceto[0] * s->repl_ceto[1] * s->repl_ceto[2];
      double *sk = (double *)malloc(sizeof(double) * 2 * cube);
      po_sk_full(s, C, sk);
      double acc = 0;
      for (int lx = 0; lx < s->repl_ceto[0]; lx++)
        for (int ly = 0; ly < s->repl_ceto[1]; ly++)
          for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
            size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
            if (!(s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)) continue;
            double dre = 0, dim = 0;
            for (int k = f; k < l; k++) {
              if (!s->moved[k] || s->q[k] == 0) continue;
              double wn[3], wo[3];
              wrap_pos(s, T + 3 * k, wn);
              wrap_pos(s, C + 3 * k, wo);
              double pn = s->kx[lx] * wn[0] + s->ky[ly] * wn[1] + s->kz[lz] * wn[2];
              double po = s->kx[lx] * wo[0] + s->ky[ly] * wo[1] + s->kz[lz] * wo[2];
              dre += s->q[k] * (cos(pn) - cos(po));
              dim += s->q[k] * (sin(pn) - sin(po));
            }
            acc += s->ek2[idx] * (2 * (sk[2 * idx] * dre + sk[2 * idx + 1] * dim) + dre * dre + dim * dim);
          }
      free(sk);
      dk = 2 * kPi * s->lB / s->box_vol * acc;
      d += dk;
    }
    /* self: potential_ewald.cc:516-524 — always zero for a move */
    if (s->dipole_correction) {
      /* DipoleE(mols) reads CURRENT coordinates: the lag of SURVEY.md §0.5 */
      s->trial_dipl_E = dipole_from_mz(s, o->mz_current);
      d += s->trial_dipl_E - s->current_dipl_E;
    }
    s->dE_ewald = d;
    o->ewald = d; o->real = dr; o->recip = dk;
    dE += d;
  }
  /* --- bond: potential_spring.cc:63-80 */
  if (s->bond_kind != PG_BOND_NONE) {
    double d = 0;
    if (len > 1) d = molecule_bond_energy(s, T, f, l) - molecule_bond_energy(s, C, f, l);
    s->dE_bond = d;
    o->bond = d;
    dE += d;
  }
  o->dE = dE;
  return 0;
}

/* ForceField::FinalizeEnergies force_field.cc:436-451 (+ the driver's
 * UpdateCurrentPos / UpdateTrialPos, simulation.cc:337-351). */
int po_commit(po_system *s, int accept) {
  if (s->pending_mol < 0) return -4;
  int f = s->mol_first[s->pending_mol], l = s->mol_first[s->pending_mol + 1];
  if (accept) {
    s->E_pair += s->dE_pair;
    s->E_ewald += s->dE_ewald;
    s->E_bond += s->dE_bond;
    s->E_ext += s->dE_ext;
    memcpy(s->cur + 3 * f, s->tri + 3 * f, sizeof(double) * 3 * (size_t)(l - f));
    if (s->dipole_correction) s->current_dipl_E = s->trial_dipl_E;
  } else {
    memcpy(s->tri + 3 * f, s->cur + 3 * f, sizeof(double) * 3 * (size_t)(l - f));
    if (s->dipole_correction) s->trial_dipl_E = s->current_dipl_E;
  }
  memset(s->moved + f, 0, (size_t)(l - f));
  s->pending_mol = -1;
  return 0;
}

/* ---------------------------------------------------------------- CBMC -- */

/* ForceField::BeadsEnergy cbmc.cc:5-151.  chain_xyz/chain_q/chain_type hold
 * current_len monomers followed by current_len counter-ions (only when use_bead2).
 * Partners skipped: molecules [skip_mol_first, skip_mol_last].
 * Returns the function's return value; pair_e/ewald_e are its two partial sums. */
double po_beads_energy(const po_system *s, const double *b1, int t1, double q1, const double *b2, int t2, double q2,
                       int use_bead2, int current_len, const double *chain_xyz, const double *chain_q,
                       const int32_t *chain_type, int skip_mol_first, int skip_mol_last, double *pair_out,
                       double *ewald_out) {
  double pair_e = 0, ewald_e = 0;
  double pair_e1 = 0, pair_e2 = 0, ewald_r1 = 0, ewald_r2 = 0, ewald_k1 = 0, ewald_k2 = 0;
  const double *C = s->cur;
  int broke = 0;
  for (int m = 0; m < s->n_mol && !broke; m++) {
    if (!(skip_mol_first < 0 || (m < skip_mol_first || m > skip_mol_last))) continue;
    for (int j = s->mol_first[m]; j < s->mol_first[m + 1]; j++) {
      if (s->pair_kind != PG_PAIR_NONE) {
        pair_e1 = pair_energy(s, b1, t1, C + 3 * j, s->type[j]);
        if (use_bead2) pair_e2 = pair_energy(s, b2, t2, C + 3 * j, s->type[j]);
      }
      pair_e += pair_e1 + pair_e2;
      if (pair_e >= kVeryLargeEnergy) { broke = 1; break; }
      if (s->use_ewald) {
        ewald_r1 = pair_real(s, b1, q1, C + 3 * j, s->q[j]);
        ewald_k1 = pair_repl(s, b1, q1, C + 3 * j, s->q[j]);
        if (use_bead2) {
          ewald_r2 = pair_real(s, b2, q2, C + 3 * j, s->q[j]);
          ewald_k2 = pair_repl(s, b2, q2, C + 3 * j, s->q[j]);
        }
      }
      ewald_e += (ewald_r1 + ewald_k1 + ewald_r2 + ewald_k2);
    }
  }
  for (int i = 0; i < current_len; i++) {
    const double *ci = chain_xyz + 3 * i;
    const double *ii = chain_xyz + 3 * (i + current_len);
    if (s->pair_kind != PG_PAIR_NONE) {
      pair_e1 = 0;
      if (i < current_len - 1) pair_e1 += pair_energy(s, b1, t1, ci, chain_type[i]);
      if (use_bead2) {
        pair_e1 += pair_energy(s, b1, t1, ii, chain_type[i + current_len]);
        pair_e2 = pair_energy(s, b2, t2, ci, chain_type[i]);
        pair_e2 += pair_energy(s, b2, t2, ii, chain_type[i + current_len]);
      }
    }
    pair_e += pair_e1 + pair_e2;
    if (pair_e >= kVeryLargeEnergy) break;
    if (s->use_ewald) {
      ewald_r1 = pair_real(s, b1, q1, ci, chain_q[i]);
      ewald_k1 = pair_repl(s, b1, q1, ci, chain_q[i]);
      if (use_bead2) {
        ewald_r1 += pair_real(s, b1, q1, ii, chain_q[i + current_len]);
        ewald_k1 += pair_repl(s, b1, q1, ii, chain_q[i + current_len]);
        ewald_r2 = pair_real(s, b2, q2, ci, chain_q[i]);
        ewald_k2 = pair_repl(s, b2, q2, ci, chain_q[i]);
        ewald_r2 += pair_real(s, b2, q2, ii, chain_q[i + current_len]);
        ewald_k2 += pair_repl(s, b2, q2, ii, chain_q[i + current_len]);
      }
    }
    ewald_e += (ewald_r1 + ewald_k1 + ewald_r2 + ewald_k2);
  }
  if (s->pair_kind != PG_PAIR_NONE && use_bead2) pair_e += pair_energy(s, b1, t1, b2, t2);
  if (s->use_ewald && pair_e < kVeryLargeEnergy) {
    ewald_e += self_energy(s, q1);
    ewald_e += 0.5 * pair_real(s, b1, q1, b1, q1);
    ewald_e += 0.5 * pair_repl(s, b1, q1, b1, q1);
    if (use_bead2) {
      ewald_e += self_energy(s, q2);
      ewald_e += pair_real(s, b1, q1, b2, q2);
      ewald_e += pair_repl(s, b1, q1, b2, q2);
      ewald_e += 0.5 * pair_real(s, b2, q2, b2, q2);
      ewald_e += 0.5 * pair_repl(s, b2, q2, b2, q2);
    }
    if (s->dipole_correction) {
      /* DipoleEDiff potential_ewald_coul.cc:473-530 */
      double Mz_o = 0;
      for (int m = 0; m < s->n_mol; m++) {
        if (!(skip_mol_first < 0 || (m < skip_mol_first || m > skip_mol_last))) continue;
        for (int j = s->mol_first[m]; j < s->mol_first[m + 1]; j++) Mz_o += s->q[j] * C[3 * j + 2];
      }
      for (int i = 0; i < current_len; i++) Mz_o += chain_q[i] * chain_xyz[3 * i + 2];
      if (use_bead2)
        for (int i = current_len; i < 2 * current_len; i++) Mz_o += chain_q[i] * chain_xyz[3 * i + 2];
      double Mz_n = Mz_o;
      Mz_n += q1 * b1[2];
      if (use_bead2) Mz_n += q2 * b2[2];
      ewald_e += s->lB * 2 * kPi / (s->box_vol / 1.0) * (Mz_n * Mz_n - Mz_o * Mz_o);
    }
  }
  if (s->ext_kind != PG_EXT_NONE) {
    pair_e += wall_energy(s, b1, t1);
    if (use_bead2) pair_e += wall_energy(s, b2, t2);
  }
  if (pair_out) *pair_out = pair_e;
  if (ewald_out) *ewald_out = ewald_e;
  if (pair_e >= kVeryLargeEnergy) return kVeryLargeEnergy;
  return pair_e + ewald_e;
}

/* ------------------------------------------------ GC insertion/deletion -- */

/* Accepted insertion: molecules appended at the end (cbmc.cc:306-320), then
 * EnergyInitForAddedMolecule (force_field.cc:1088-1105):
 *   pair   potential_pair.cc:102-150
 *   ewald  TrialChainEnergy potential_ewald.cc:233-291 + EnergyInitForLastMol :337-427
 *   ext    potential_external.cc:63-80
 *   bond   potential_bond.cc:29-34 (energy of the LAST molecule only) */
int po_insert_molecules(po_system *s, int n_new_mol, const int32_t *mol_len, const double *xyz, const double *q,
                        const int32_t *type, pg_totals *added) {
  int n_add = 0;
  for (int m = 0; m < n_new_mol; m++) n_add += mol_len[m];
  int n0 = s->n;
  reserve(s, n0 + n_add, s->n_mol + n_new_mol);
  memcpy(s->cur + 3 * n0, xyz, sizeof(double) * 3 * (size_t)n_add);
  memcpy(s->tri + 3 * n0, xyz, sizeof(double) * 3 * (size_t)n_add);
  memcpy(s->q + n0, q, sizeof(double) * (size_t)n_add);
  for (int i = 0; i < n_add; i++) { s->type[n0 + i] = type[i]; s->moved[n0 + i] = 0; }
  if (s->n_mol == 0) s->mol_first[0] = 0;
  for (int m = 0; m < n_new_mol; m++) s->mol_first[s->n_mol + m + 1] = s->mol_first[s->n_mol + m] + mol_len[m];
  const double *C = s->cur;
  pg_totals a;
  memset(&a, 0, sizeof(a));
  if (s->pair_kind == PG_PAIR_TRUNCATED_LJ) {
    double E = 0;
    for (int i = n0; i < n0 + n_add; i++) {
      for (int k = 0; k < n0; k++) E += pair_energy(s, C + 3 * i, s->type[i], C + 3 * k, s->type[k]);
      for (int k = n0; k < i; k++) E += pair_energy(s, C + 3 * i, s->type[i], C + 3 * k, s->type[k]);
    }
    a.pair = E; /* HardSphere adds 0 (potential_pair.cc:118-120) */
  }
  if (s->use_ewald) {
    double d = 0, er = 0, ek = 0, es = 0;
    for (int i = n0; i < n0 + n_add; i++)
      for (int j = n0; j <= i; j++) {
        double r = pair_real(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
        double k = pair_repl(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
        if (j == i) { r *= 0.5; k *= 0.5; }
        d += r + k; er += r; ek += k;
      }
    for (int i = n0; i < n0 + n_add; i++)
      for (int j = 0; j < n0; j++) {
        double r = pair_real(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
        double k = pair_repl(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
        d += r + k; er += r; ek += k;
      }
    for (int i = n0; i < n0 + n_add; i++) { double e = self_energy(s, s->q[i]); d += e; es += e; }
    a.real = er; a.recip = ek; a.self = es;
    s->n += n_add; /* dipole over the new configuration */
    if (s->dipole_correction) {
      double cur = dipole_from_mz(s, mz_current(s, -1, -1));
      a.dipole = cur - s->trial_dipl_E;
      d += a.dipole;
      s->current_dipl_E = cur;
      s->trial_dipl_E = cur;
    }
    s->n -= n_add;
    a.ewald = d;
  }
  if (s->ext_kind != PG_EXT_NONE) {
    double E = 0;
    for (int i = n0; i < n0 + n_add; i++) E += wall_energy(s, C + 3 * i, s->type[i]);
    a.ext = E;
  }
  s->n += n_add;
  s->n_mol += n_new_mol;
  if (s->bond_kind != PG_BOND_NONE)
    a.bond = molecule_bond_energy(s, C, s->mol_first[s->n_mol - 1], s->mol_first[s->n_mol]);
  s->E_pair += a.pair; s->E_ewald += a.ewald; s->E_ext += a.ext; s->E_bond += a.bond;
  if (added) *added = a;
  return 0;
}

/* Accepted deletion of molecules [mf, ml] (a chain followed by its counter-ions):
 * the four AdjustEnergyUponMolDeletion (potential_pair.cc:249-294,
 * potential_ewald.cc:626-707, potential_external.cc:82-98, potential_bond.cc:36-40). */
int po_delete_molecules(po_system *s, int mf, int ml, pg_totals *removed) {
  int b0 = s->mol_first[mf], b1 = s->mol_first[ml + 1];
  const double *C = s->cur;
  pg_totals r;
  memset(&r, 0, sizeof(r));
  if (s->pair_kind != PG_PAIR_NONE) {
    double E = 0;
    int gap = (s->pair_kind == PG_PAIR_HARD_SPHERE) ? 1 : 0;
    /* every pair with at least one bead in the deleted range, once */
    for (int i = b0; i < b1; i++) {
      for (int j = 0; j < b0; j++) E += pair_energy(s, C + 3 * i, s->type[i], C + 3 * j, s->type[j]);
      for (int j = b1; j < s->n; j++) E += pair_energy(s, C + 3 * i, s->type[i], C + 3 * j, s->type[j]);
      for (int j = i + 1; j < b1; j++) {
        int same_chain = (i < s->mol_first[mf + 1] && j < s->mol_first[mf + 1]);
        if (same_chain && gap && j == i + 1) continue;
        E += pair_energy(s, C + 3 * i, s->type[i], C + 3 * j, s->type[j]);
      }
    }
    r.pair = E;
  }
  if (s->use_ewald) {
    double d = 0, er = 0, ek = 0, es = 0;
    for (int i = b0; i < b1; i++) {
      for (int j = 0; j < s->n; j++) {
        if (j >= b0 && j < i) continue; /* counted when i was the smaller index */
        double rr = pair_real(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
        double kk = pair_repl(s, C + 3 * i, s->q[i], C + 3 * j, s->q[j]);
        if (j == i) { rr *= 0.5; kk *= 0.5; }
        d += rr + kk; er += rr; ek += kk;
      }
    }
    for (int i = b0; i < b1; i++) { double e = self_energy(s, s->q[i]); d += e; es += e; }
    r.real = er; r.recip = ek; r.self = es;
    if (s->dipole_correction) {
      double dip = dipole_from_mz(s, mz_current(s, b0, b1));
      r.dipole = s->current_dipl_E - dip;
      d += r.dipole;
      s->current_dipl_E = dip;
      s->trial_dipl_E = dip;
    }
    r.ewald = d;
  }
  if (s->ext_kind != PG_EXT_NONE) {
    double E = 0;
    for (int i = b0; i < b1; i++) E += wall_energy(s, C + 3 * i, s->type[i]);
    r.ext = E;
  }
  if (s->bond_kind != PG_BOND_NONE) r.bond = molecule_bond_energy(s, C, s->mol_first[mf], s->mol_first[mf + 1]);
  s->E_pair -= r.pair; s->E_ewald -= r.ewald; s->E_ext -= r.ext; s->E_bond -= r.bond;
  /* compact */
  int nrem = b1 - b0, nm = ml - mf + 1;
  memmove(s->cur + 3 * b0, s->cur + 3 * b1, sizeof(double) * 3 * (size_t)(s->n - b1));
  memmove(s->tri + 3 * b0, s->tri + 3 * b1, sizeof(double) * 3 * (size_t)(s->n - b1));
  memmove(s->q + b0, s->q + b1, sizeof(double) * (size_t)(s->n - b1));
  memmove(s->type + b0, s->type + b1, sizeof(int) * (size_t)(s->n - b1));
  for (int m = ml + 1; m <= s->n_mol; m++) s->mol_first[m - nm] = s->mol_first[m] - nrem;
  s->n -= nrem;
  s->n_mol -= nm;
  if (removed) *removed = r;
  return 0;
}

/* ------------------------------------------- wall-force pressure sampler -- */

/* src/force_field/potential_truncated_lj.cc:87-122 as written for r > 0; the reference tests an
 * uninitialised r6 there (undefined behaviour - gcc -O3 folds it to "always 1e8"), so the LJ columns of
 * its output cannot be pinned: this is the evident intent (r <= 0 -> 1e8, else the LJ force).
 * potential_hard_sphere.cc:51-57 returns 0. */
static double pair_force(const po_system *s, const double *a, int ta, const double *b, int tb) {
  if (s->pair_kind == PG_PAIR_HARD_SPHERE) return 0.0;
  double r = bbdist(s, a, b, s->npbc);
  double sigma = (s->lj_sigma[ta] + s->lj_sigma[tb]) / 2;
  double epsilon = sqrt(s->lj_epsilon[ta] * s->lj_epsilon[tb]);
  double force = 0, r6;
  if (r <= 0) {
    force = kVeryLargeEnergy;
  } else if (s->lj_cutoff < 0) {
    if (r < k216 * sigma) {
      r6 = pow((sigma / r), 6);
      force = 4 * epsilon * (r6 * r6 * 12 / r - r6 * 6 / r);
    }
  } else if (r < s->lj_cutoff) {
    r6 = pow((sigma / r), 6);
    force = 4 * epsilon * (r6 * r6 * 12 / r - r6 * 6 / r);
  } else {
    r6 = pow((sigma / s->lj_cutoff), 6);
    force = 4 * epsilon * (r6 * r6 * 12 / r - r6 * 6 / r);
  }
  return force;
}

/* src/force_field/potential_ewald_coul.cc:261-293: z force on bead 1 (the wall site) */
static double pair_force_z_real(const po_system *s, const double *b1, double q1, const double *b2, double q2) {
  double force_z = 0;
  if (q1 * q2 == 0) return force_z;
  double r[3];
  dist_vector(s, b2, b1, r);
  double prefactor = s->lB * q1 * q2;
  double prefactor2 = 2 * sqrt(s->alpha / kPi);
  for (int i = -s->real_cell[0]; i <= s->real_cell[0]; i++)
    for (int j = -s->real_cell[1]; j <= s->real_cell[1]; j++)
      for (int k = -s->real_cell[2]; k <= s->real_cell[2]; k++) {
        double rx = r[0] + i * s->ebox[0], ry = r[1] + j * s->ebox[1], rz = r[2] + k * s->ebox[2];
        double d = sqrt(rx * rx + ry * ry + rz * rz);
        if (d > 0 && d <= s->real_cutoff)
          force_z += prefactor * ((prefactor2 * exp(-s->alpha * d * d)) + (erfc(sqrt(s->alpha) * d) / d)) * rz / (d * d);
      }
  return force_z;
}

/* src/force_field/potential_ewald_coul.cc:297-324 */
static double pair_force_z_repl(const po_system *s, const double *b1, double q1, const double *b2, double q2) {
  double force_z = 0;
  if (q1 * q2 == 0) return force_z;
  double r[3];
  dist_vector(s, b2, b1, r);
  double prefactor = s->lB * 4 * kPi * q1 * q2 / s->box_vol;
  for (int lx = 0; lx < s->repl_ceto[0]; lx++)
    for (int ly = 0; ly < s->repl_ceto[1]; ly++)
      for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
        if (s->k2[idx] > 0 && s->k2[idx] <= s->repl_cutoff)
          force_z += prefactor * s->kz[lz] * s->ek2[idx] * sin(s->kx[lx] * r[0] + s->ky[ly] * r[1] + s->kz[lz] * r[2]);
      }
  return force_z;
}

/* src/force_field/potential_truncated_lj_wall.cc:136-217 (hard and well walls: 0,
 * potential_hard_wall.cc:49-53, potential_well_wall.cc:57-62) */
static double bead_force_on_wall(const po_system *s, const double *p, int t) {
  if (s->ext_kind != PG_EXT_TRUNCATED_LJ_WALL) return 0.0;
  double force = 0;
  double z = p[2], Lz = s->box[2];
  double sigma = s->wall_sigma[t], epsilon = s->wall_epsilon[t];
  double R0 = 3 * k213 * sigma, K = 1;
  int graft = s->graft_kind[t];
  if (epsilon == 0) return 0.0;
  if (z <= 0 || z >= Lz) return kVeryLargeEnergy;
  if (s->wall_cut < 0 && graft == 0) {
    if (z < k213 * sigma) {
      double r3 = pow((sigma / z), 3);
      force += 2.59807621135 * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);
    }
    if (Lz - z < k213 * sigma) {
      double r3 = pow((sigma / (Lz - z)), 3);
      force += 2.59807621135 * epsilon * (r3 * r3 * 6 / (Lz - z) - r3 * 3 / (Lz - z));
    }
  } else if (graft == 1) { /* "L" */
    double r3 = pow((sigma / z), 3);
    force += 2.59807621135 * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);
    force += -K * R0 * z / (1 - pow(z / R0, 2));
    if (Lz - z < k213 * sigma) {
      double r3b = pow((sigma / (Lz - z)), 3);
      force += 2.59807621135 * epsilon * (r3b * r3b * 6 / (Lz - z) - r3b * 3 / (Lz - z));
    }
  } else if (graft == 2) { /* "R" */
    double r3 = pow((sigma / (Lz - z)), 3);
    force += 2.59807621135 * epsilon * (r3 * r3 * 6 / (Lz - z) - r3 * 3 / (Lz - z));
    force += -K * R0 * (Lz - z) / (1 - pow((Lz - z) / R0, 2));
    if (z < k213 * sigma) {
      double r3b = pow((sigma / z), 3);
      force += 2.59807621135 * epsilon * (r3b * r3b * 6 / z - r3b * 3 / z);
    }
  } else {
    double m_cut = s->wall_cut;
    if (z < m_cut) {
      double r3 = pow((sigma / z), 3);
      force += 2.59807621135 * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);
    } else {
      double r3 = pow((sigma / m_cut), 3);
      force += 2.59807621135 * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);
    }
    if (Lz - z < m_cut) {
      double r3 = pow((sigma / (Lz - z)), 3);
      force += 2.59807621135 * epsilon * (r3 * r3 * 6 / (Lz - z) - r3 * 3 / (Lz - z));
    } else {
      double r3 = pow((sigma / m_cut), 3);
      force += 2.59807621135 * epsilon * (r3 * r3 * 6 / z - r3 * 3 / z);
    }
  }
  return force;
}

/* One sample of ForceField::CalcPressureForceLJELSlit, src/force_field/pressure.cc:404-469: the six
 * force sums of this configuration, out = {LJ ion, LJ polymer, LJ wall-wall, EL ion, EL polymer, EL wall-wall}
 * (the caller accumulates them into p_tensor[6..11]; wall-wall only on its first sample).
 * `phantom` wall sites are the first molecules (one bead each), half on each plate. */
int po_wall_force(const po_system *s, int phantom, double out[6]) {
  for (int i = 0; i < 6; i++) out[i] = 0;
  if (phantom < 0 || phantom > s->n_mol) return -1;
  for (int i = 0; i < phantom; i++) {
    const double *pi = s->cur + 3 * (size_t)s->mol_first[i];
    int ti = s->type[s->mol_first[i]];
    double qi = s->q[s->mol_first[i]];
    for (int j = phantom / 2; j < s->n_mol; j++) {
      int len = s->mol_first[j + 1] - s->mol_first[j];
      for (int k = 0; k < len; k++) {
        int b = s->mol_first[j] + k;
        const double *pj = s->cur + 3 * (size_t)b;
        double C;
        if (i < phantom / 2 && j >= phantom) C = -0.5;
        else if (i < phantom && j >= phantom) C = 0.5;
        else C = -1;
        int k_id = 2;
        if (j >= phantom) k_id = (len > 1) ? 1 : 0;
        if (s->pair_kind != PG_PAIR_NONE && (i < phantom / 2 || j >= phantom)) {
          double r[3];
          for (int a = 0; a < 3; a++) { /* GetDistVector(mols[j], mols[i], ForceField::box_l) */
            double di = pi[a] - pj[a];
            if (a < s->npbc) di -= s->box[a] * round(di / s->box[a]);
            r[a] = di;
          }
          double zc = r[2] / sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
          out[k_id] += C * zc * pair_force(s, pj, s->type[b], pi, ti);
        }
        if (s->use_ewald && (i < phantom / 2 || j >= phantom)) {
          double f = C * pair_force_z_real(s, pi, qi, pj, s->q[b]);
          f += C * pair_force_z_repl(s, pi, qi, pj, s->q[b]);
          out[3 + k_id] += f;
        }
      }
    }
  }
  if (s->ext_kind != PG_EXT_NONE) {
    for (int i = phantom; i < s->n_mol; i++) {
      int len = s->mol_first[i + 1] - s->mol_first[i];
      for (int j = 0; j < len; j++) {
        int b = s->mol_first[i] + j;
        out[len > 1 ? 1 : 0] += 0.5 * bead_force_on_wall(s, s->cur + 3 * (size_t)b, s->type[b]);
      }
    }
  }
  return 0;
}

/* ------------------------------------------ volume-perturbation pressure -- */
/* src/force_field/potential_ewald_coul.cc:167-196 — PairEnergyRealForP: distance vector wrapped in the cell
 * stretched along z by dz, images stepped by the stretched edge; no zero-charge shortcut. */
static double pair_real_forp(const po_system *s, double dz, const double *b1, double q1, const double *b2, double q2) {
  double energy = 0;
  double dist[3];
  double box_l_scaled[3] = {s->ebox[0], s->ebox[1], s->ebox[2] + dz};
  for (int i = 0; i < 3; i++) {
    double di = b2[i] - b1[i];
    if (i < s->npbc) di -= box_l_scaled[i] * round(di / box_l_scaled[i]);
    dist[i] = di;
  }
  double prefactor = s->lB * q1 * q2;
  for (int i = -s->real_cell[0]; i <= s->real_cell[0]; i++)
    for (int j = -s->real_cell[1]; j <= s->real_cell[1]; j++)
      for (int k = -s->real_cell[2]; k <= s->real_cell[2]; k++) {
        double rx = dist[0] + i * s->ebox[0];
        double ry = dist[1] + j * s->ebox[1];
        double rz = dist[2] + k * (s->ebox[2] + dz);
        double r = sqrt(rx * rx + ry * ry + rz * rz);
        if (r > 0 && r <= s->real_cutoff) energy += prefactor * erfc(sqrt(s->alpha) * r) / r;
      }
  return energy;
}

/* src/force_field/potential_ewald_coul.cc:228-250 — PairEnergyReplForP with the k_z, k^2 and ek2 tables of
 * the stretched cell (potential_ewald_coul.cc:86-112) and its volume. */
static double pair_repl_forp(const po_system *s, double dz, const double *kz_forp, const double *k2_forp,
                             const double *ek2_forp, const double *b1, double q1, const double *b2, double q2) {
  double energy = 0;
  double r[3];
  double box_l_scaled[3] = {s->ebox[0], s->ebox[1], s->ebox[2] + dz};
  for (int i = 0; i < 3; i++) {
    double di = b2[i] - b1[i];
    if (i < s->npbc) di -= box_l_scaled[i] * round(di / box_l_scaled[i]);
    r[i] = di;
  }
  double box_vol_forp = s->ebox[0] * s->ebox[1] * (s->ebox[2] + dz);
  double prefactor = s->lB * q1 * q2 / (kPi * box_vol_forp) * (4 * kPi * kPi);
  for (int lx = 0; lx < s->repl_ceto[0]; lx++)
    for (int ly = 0; ly < s->repl_ceto[1]; ly++)
      for (int lz = 0; lz < s->repl_ceto[2]; lz++) {
        size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * lx + (size_t)s->repl_ceto[2] * ly + lz;
        if (k2_forp[idx] > 0 && k2_forp[idx] <= s->repl_cutoff)
          energy += prefactor * ek2_forp[idx] * cos(s->kx[lx] * r[0] + s->ky[ly] * r[1] + kz_forp[lz] * r[2]);
      }
  return energy;
}

/* One sample of ForceField::CalcPressureVolScalingHSELSlit, src/force_field/pressure.cc:187-338 (the
 * accumulation into averages, :340-384, is the caller's): out[0..15] electrostatic dU per species pair
 * (index i*4+j, i <= j; 0 cation, 1 anion, 2 polymer, 3 surface), out[16..31] LJ/HS + wall dU, out[32] bond
 * dU, out[33] dipole dU, out[34] total dU, out[35] n_mol - phantom.  "Stored" pair energies of the
 * reference's maps are recomputed from the current coordinates (they are functions of those). */
int po_vol_scaling(const po_system *s, int phantom, double dz, double out[36]) {
  for (int i = 0; i < 36; i++) out[i] = 0;
  if (phantom < 0 || phantom > s->n_mol) return -1;
  const int n_mol = s->n_mol;
  double dU = 0, oldE, newE;
  po_system sz = *s; /* shallow: the same tables, the slab box stretched (box_l_scaled, pressure.cc:195) */
  sz.box[2] = s->box[2] + dz;
  double *kz_forp = NULL, *k2_forp = NULL, *ek2_forp = NULL;
  if (s->use_ewald) {
    size_t cube = (size_t)s->repl_ceto[0] * s->repl_ceto[1] * s->repl_ceto[2];
    kz_forp = (double *)malloc(sizeof(double) * s->repl_ceto[2]);
    k2_forp = (double *)malloc(sizeof(double) * cube);
    ek2_forp = (double *)malloc(sizeof(double) * cube);
    for (int lz = -s->repl_cell[2]; lz <= s->repl_cell[2]; lz++)
      kz_forp[lz + s->repl_cell[2]] = lz * 2 * kPi / (s->ebox[2] + dz);
    for (int ix = 0; ix < s->repl_ceto[0]; ix++)
      for (int iy = 0; iy < s->repl_ceto[1]; iy++)
        for (int iz = 0; iz < s->repl_ceto[2]; iz++) {
          size_t idx = (size_t)s->repl_ceto[1] * s->repl_ceto[2] * ix + (size_t)s->repl_ceto[2] * iy + iz;
          k2_forp[idx] = s->kx[ix] * s->kx[ix] + s->ky[iy] * s->ky[iy] + kz_forp[iz] * kz_forp[iz];
          ek2_forp[idx] = exp(-k2_forp[idx] / (4 * s->alpha)) / k2_forp[idx];
        }
  }
  double *d_com_z = (double *)calloc((size_t)(n_mol > 0 ? n_mol : 1), sizeof(double));
  double *tz = (double *)malloc(sizeof(double) * 3 * (size_t)(s->n > 0 ? s->n : 1)); /* scaled (trial) coordinates */
  /* 2. displacements */
  for (int i = 0; i < n_mol; i++) {
    int first = s->mol_first[i], len = s->mol_first[i + 1] - first;
    double com_z = 0;
    for (int j = 0; j < len; j++) com_z += s->cur[3 * (first + j) + 2];
    com_z /= len;
    d_com_z[i] = dz * (com_z / s->box[2]);
    int g = s->graft_kind[s->type[first]];
    if (g == PG_GRAFT_RIGHT) d_com_z[i] = dz;
    else if (g == PG_GRAFT_LEFT) d_com_z[i] = 0;
  }
  /* 3. scaled coordinates */
  for (int i = 0; i < n_mol; i++)
    for (int b = s->mol_first[i]; b < s->mol_first[i + 1]; b++) {
      tz[3 * b] = s->cur[3 * b]; tz[3 * b + 1] = s->cur[3 * b + 1];
      if (s->bond_kind != PG_BOND_NONE && s->graft_kind[s->type[b]] == PG_GRAFT_NONE)
        tz[3 * b + 2] = s->cur[3 * b + 2] * (1.0 + dz / s->box[2]);
      else
        tz[3 * b + 2] = s->cur[3 * b + 2] + d_com_z[i];
    }
  for (int i = 0; i < n_mol; i++) {
    int len_i = s->mol_first[i + 1] - s->mol_first[i];
    int id_i;
    if (i < phantom) id_i = 3;
    else if (len_i > 1) id_i = 2;
    else if (s->q[s->mol_first[i]] >= 0) id_i = 0;
    else id_i = 1;
    for (int j = 0; j < len_i; j++) {
      int bi = s->mol_first[i] + j;
      for (int k = i; k < n_mol; k++) {
        int len_k = s->mol_first[k + 1] - s->mol_first[k];
        int id_k;
        if (k < phantom) id_k = 3;
        else if (len_k > 1) id_k = 2;
        else if (s->q[s->mol_first[k]] >= 0) id_k = 0;
        else id_k = 1;
        for (int l = 0; l < len_k; l++) {
          int bk = s->mol_first[k] + l;
          if ((k > i || (k == i && l >= j)) && (i >= phantom || (i < phantom / 2 && k >= phantom / 2))) {
            int index1 = (id_i < id_k) ? id_i : id_k;
            int index2 = (id_i > id_k) ? id_i : id_k;
            int index = index1 * 4 + index2;
            if (s->use_ewald) {
              /* GetERealRepl(0, ., .): the stored real + reciprocal pair energy, half for a bead with itself
               * (potential_ewald.cc EnergyInitialization) */
              oldE = pair_real(s, s->cur + 3 * bi, s->q[bi], s->cur + 3 * bk, s->q[bk]) +
                     pair_repl(s, s->cur + 3 * bi, s->q[bi], s->cur + 3 * bk, s->q[bk]);
              if (bi == bk) oldE *= 0.5;
              newE = pair_real_forp(s, dz, tz + 3 * bi, s->q[bi], tz + 3 * bk, s->q[bk]);
              newE += pair_repl_forp(s, dz, kz_forp, k2_forp, ek2_forp, tz + 3 * bi, s->q[bi], tz + 3 * bk, s->q[bk]);
              if (bi == bk) newE *= 0.5;
              dU += newE - oldE;
              out[index] += newE - oldE;
            }
            if (s->pair_kind != PG_PAIR_NONE && i >= phantom && k >= phantom && bi != bk) {
              /* stored energy: 0 for bonded hard-sphere neighbours (potential_pair.cc:64-67) */
              if (s->pair_kind == PG_PAIR_HARD_SPHERE && k == i && l == j + 1) oldE = 0;
              else oldE = pair_energy(s, s->cur + 3 * bi, s->type[bi], s->cur + 3 * bk, s->type[bk]);
              newE = pair_energy(&sz, tz + 3 * bi, s->type[bi], tz + 3 * bk, s->type[bk]);
              dU += newE - oldE;
              out[16 + index] += newE - oldE;
            }
          }
        }
      }
      if (s->ext_kind != PG_EXT_NONE && i >= phantom) {
        int index1 = (id_i < 3) ? id_i : 3;
        int index2 = (id_i > 3) ? id_i : 3;
        int index = index1 * 4 + index2;
        oldE = wall_energy(s, s->cur + 3 * bi, s->type[bi]);
        newE = wall_energy(&sz, tz + 3 * bi, s->type[bi]);
        dU += newE - oldE;
        out[16 + index] += newE - oldE;
      }
    }
    if (s->bond_kind != PG_BOND_NONE) {
      oldE = molecule_bond_energy(s, s->cur, s->mol_first[i], s->mol_first[i + 1]);
      newE = molecule_bond_energy(&sz, tz, s->mol_first[i], s->mol_first[i + 1]);
      dU += newE - oldE;
      out[32] += newE - oldE;
    }
  }
  if (s->use_ewald && s->dipole_correction) {
    double Mz_old = 0, Mz_new = 0;
    for (int b = 0; b < s->n; b++) {
      Mz_old += s->q[b] * s->cur[3 * b + 2];
      Mz_new += s->q[b] * tz[3 * b + 2];
    }
    double vol = s->box[0] * s->box[1] * s->box[2];
    double d_di = (s->lB * 2 * kPi) * (Mz_new * Mz_new / (s->box[0] * s->box[1] * (s->box[2] + dz)) - Mz_old * Mz_old / vol);
    dU += d_di;
    out[33] += d_di;
  }
  out[34] = dU;
  out[35] = n_mol - phantom;
  free(d_com_z); free(tz); free(kz_forp); free(k2_forp); free(ek2_forp);
  return 0;
}

/* --------------------------------------------------- standalone helpers -- */
/* Single-pair entry points for unit tests of the primitives. */
double po_pair_energy(const po_system *s, const double *a, int ta, const double *b, int tb) {
  return pair_energy(s, a, ta, b, tb);
}
double po_pair_real(const po_system *s, const double *a, double qa, const double *b, double qb) {
  return pair_real(s, a, qa, b, qb);
}
double po_pair_repl(const po_system *s, const double *a, double qa, const double *b, double qb) {
  return pair_repl(s, a, qa, b, qb);
}
double po_wall_energy(const po_system *s, const double *p, int t) { return wall_energy(s, p, t); }
