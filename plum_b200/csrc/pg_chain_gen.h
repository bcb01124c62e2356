// plum_b200 — the random stream and the draw order of Plum's translational step, as the device-resident
// Markov-chain kernel (k_chain, pg_chain.cu) consumes them.  Host + device: the CPU tests compile this header
// with g++ and pin it, draw for draw, to plum_b200/host/mc_propose.h (which is pinned to the trial
// coordinates the reference itself wrote).
//
// Reference behaviour restated (file:line relative to /root/reference):
//   std::mt19937 (the driver's `rand_gen`, src/simulation/simulation.h)  — libstdc++ layout: 624 state words +
//       position p; a draw tempers word p, the whole state is regenerated ("twist") when p reaches 624
//   Simulation::Run / TranslationalMove          src/simulation/simulation.cc:216-355   (draw order of a step)
//   Molecule::BeadTranslate / COMTranslate / Pivot / RandomReptation   src/molecules/molecule.cc:103-312
//   randSphere                                   src/utilities/misc.cc:95-109
//
// The generator state lives in TWO buffers: the current generation and the next one (already twisted), so
// that any thread can look ahead up to 624 draws without changing anything; only `advance` moves the position.
#ifndef PLUM_B200_PG_CHAIN_GEN_H_
#define PLUM_B200_PG_CHAIN_GEN_H_

#include <stdint.h>

#include "pg_propose_math.h"

#define CG_N 624
#define CG_M 397

// move kinds (same numbering as include/plum_b200.h PG_MOVE_*); negative: no energy evaluation this step
#define CG_BEAD 0
#define CG_COM 1
#define CG_PIVOT 2
#define CG_CRANK 3
#define CG_M_PI 3.14159265358979323846   // M_PI of <cmath>, molecule.cc:247 (not the truncated kPi)
#define CG_REPT 4
#define CG_NONE (-1)      // the step attempts nothing (simulation.cc:242,284,315)
#define CG_STOP_GC (-2)   // the next step is a grand-canonical step: the chain stops IN FRONT of it
#define CG_STOP_ERR (-3)

struct PgMt {
  uint32_t x[2][CG_N];   // x[cur]: state words of the current generation, x[cur ^ 1]: the next generation
  int cur;
  int p;                 // next word of the current generation (0 .. 623 between steps)
  int need;              // the next-generation buffer is stale (a cooperative twist is due)
  int pad;
};

PP_HD uint32_t cg_temper(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

PP_HD uint32_t cg_mix(uint32_t a, uint32_t b, uint32_t far) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// One word of the next generation.  Phase structure of the in-place reference loop: words 0..226 read only the
// old generation, 227..453 read new words 0..226, 454..623 read new words 227..396 (and word 623 the new word 0).
PP_HD uint32_t cg_twist_word(const uint32_t* old_, const uint32_t* new_, int k) {
  if (k < CG_N - CG_M) return cg_mix(old_[k], old_[k + 1], old_[k + CG_M]);
  if (k < CG_N - 1) return cg_mix(old_[k], old_[k + 1], new_[k - (CG_N - CG_M)]);
  return cg_mix(old_[CG_N - 1], new_[0], new_[CG_M - 1]);
}

// Serial twist (host tests, and the reference for the device's cooperative version).
PP_HD void cg_twist_serial(const uint32_t* old_, uint32_t* new_) {
  for (int k = 0; k < CG_N; k++) new_[k] = cg_twist_word(old_, new_, k);
}

// Raw 32-bit output `i` draws ahead of the current position (0 <= p + i < 2 * 624).
PP_HD uint32_t cg_raw(const PgMt& m, int i) {
  const int idx = m.p + i;
  return (idx < CG_N) ? cg_temper(m.x[m.cur][idx]) : cg_temper(m.x[m.cur ^ 1][idx - CG_N]);
}

// (double)rand_gen() / rand_gen.max()
PP_HD double cg_uniform_of(uint32_t raw) { return PP_DIV((double)raw, 4294967295.0); }

// Consume n draws.  When the position crosses into the next generation the buffers swap and the (new) next
// generation must be regenerated before anyone looks past the current one again: `need` is raised.
PP_HD void cg_advance(PgMt& m, int n) {
  m.p += n;
  if (m.p >= CG_N) {
    m.p -= CG_N;
    m.cur ^= 1;
    m.need = 1;
  }
}

// A cursor over the look-ahead window: what one thread uses to walk the stream sequentially.  The first `npre` draws may
// have been tempered and converted ahead of time (by other lanes, in parallel: a uniform costs an FP64 division).
struct CgCursor {
  const PgMt* m;
  int used;
  const uint32_t* pre_raw;
  const double* pre_u;
  int npre;
  PP_HD uint32_t raw() { const int i = used++; return (i < npre) ? pre_raw[i] : cg_raw(*m, i); }
  PP_HD double uniform() { const int i = used++; return (i < npre) ? pre_u[i] : cg_uniform_of(cg_raw(*m, i)); }
};

// randSphere, misc.cc:95-109
PP_HD void cg_rand_sphere(CgCursor& r, double vec[3]) {
  double rand_square = 2, r1 = 0, r2 = 0;
  while (rand_square > 1) {
    r1 = PP_SUB(1.0, PP_MUL(2.0, r.uniform()));
    r2 = PP_SUB(1.0, PP_MUL(2.0, r.uniform()));
    rand_square = PP_ADD(PP_MUL(r1, r1), PP_MUL(r2, r2));
    if (r.used > 600) break;   // never in practice (p = 0.215^300); the caller flags the overrun
  }
  const double ranh = PP_MUL(2.0, PP_SQRT(PP_SUB(1.0, rand_square)));
  vec[0] = PP_MUL(r1, ranh);
  vec[1] = PP_MUL(r2, ranh);
  vec[2] = PP_SUB(1.0, PP_MUL(2.0, rand_square));
}

// One candidate pair of the randSphere rejection loop from two uniforms: accepted?  (pivot rows, in parallel)
PP_HD bool cg_sphere_pair(double u1, double u2, double vec[3]) {
  const double r1 = PP_SUB(1.0, PP_MUL(2.0, u1));
  const double r2 = PP_SUB(1.0, PP_MUL(2.0, u2));
  const double rand_square = PP_ADD(PP_MUL(r1, r1), PP_MUL(r2, r2));
  if (rand_square > 1) return false;
  const double ranh = PP_MUL(2.0, PP_SQRT(PP_SUB(1.0, rand_square)));
  vec[0] = PP_MUL(r1, ranh);
  vec[1] = PP_MUL(r2, ranh);
  vec[2] = PP_SUB(1.0, PP_MUL(2.0, rand_square));
  return true;
}

PP_HD double cg_vec_len(const double v[3]) {
  return PP_SQRT(PP_ADD(PP_ADD(PP_MUL(v[0], v[0]), PP_MUL(v[1], v[1])), PP_MUL(v[2], v[2])));
}

// molecule.cc:184-187: bond_len + (bond_len / 5.0) * (uniform - 0.5)
PP_HD double cg_varied_bond(CgCursor& r, double bond_len, int vary) {
  if (!vary) return bond_len;
  return PP_ADD(bond_len, PP_MUL(PP_DIV(bond_len, 5.0), PP_SUB(r.uniform(), 0.5)));
}

// What does not depend on coordinates in one translational step (the kernel's copy of pg_move_desc).
struct CgStep {
  int kind;        // CG_*
  int mol;
  int i0;          // PIVOT: pivot bead; REPT: direction; CRANK: first bead of the axis
  int i1;          // CRANK: last bead of the axis
  int n_rows;      // PIVOT: len - 1 rows still to be drawn (by cg_pivot_rows_*)
  double s;        // BEAD: 3*move_size/|v|; PIVOT: move_size_rand; REPT: bond_len; CRANK: angle
  double v[3];
  double vlen;
};

struct CgConfig {
  int n_chain, n_ion;        // movable chains (len > 1) and ions, in molecule order behind the phantoms
  int gc_freq;               // > 0: a step whose first draw % gc_freq == 0 is a GC step
  int vary_bond;
  double move_size, bond_len;
  double prob[5];
};

// The head of Simulation::Run's loop body + TranslationalMove up to the generator's own small draws
// (everything except the pivot rows, which are bulk work).  `chains` / `ions` map the pre-selected counters to
// molecule ids, `mol_len(mol)` is needed by the pivot.  Returns the draws consumed (not yet advanced).
template <class LenFn>
PP_HD int cg_step_header(const PgMt& mt, const CgConfig& c, const int* chains, const int* ions, LenFn mol_len, CgStep& d,
                         const uint32_t* pre_raw = nullptr, const double* pre_u = nullptr, int npre = 0) {
  CgCursor r{&mt, 0, pre_raw, pre_u, npre};
  d.kind = CG_NONE; d.mol = -1; d.i0 = 0; d.i1 = 0; d.n_rows = 0; d.s = 0.0; d.v[0] = d.v[1] = d.v[2] = 0.0; d.vlen = 0.0;
  const int rand_num = (int)r.raw();                                   // simulation.cc:221
  if (c.gc_freq > 0 && (rand_num % c.gc_freq == 0)) { d.kind = CG_STOP_GC; return 0; }
  if (c.n_chain + c.n_ion <= 0) return r.used;                         // simulation.cc:242
  int which_chain = (int)floor(PP_MUL(r.uniform(), (double)c.n_chain));
  int which_ion = (int)floor(PP_MUL(r.uniform(), (double)c.n_ion));
  if (which_chain == c.n_chain) which_chain--;
  if (which_ion == c.n_ion) which_ion--;
  int move_type = 0;
  const double rn = r.uniform();
  double current = c.prob[0];
  while (current < rn && move_type < 4) {                              // simulation.cc:271-276
    move_type++;
    current = PP_ADD(current, c.prob[move_type]);
  }
  if (move_type == 0 && c.n_ion > 0) {                                 // BeadTranslate, molecule.cc:103-119
    d.mol = ions[which_ion];
    cg_rand_sphere(r, d.v);
    double vec_len = cg_vec_len(d.v);
    if (vec_len > 0) vec_len = PP_DIV(PP_MUL(3.0, c.move_size), vec_len);
    d.s = vec_len;
    d.kind = CG_BEAD;
    return r.used;
  }
  if (c.n_chain <= 0 || move_type == 0) return r.used;                 // simulation.cc:284, :315-316
  d.mol = chains[which_chain];
  const int len = mol_len(d.mol);
  if (move_type == 1) {                                                // COMTranslate, molecule.cc:136-153
    for (int k = 0; k < 3; k++) d.v[k] = PP_DIV(PP_MUL(PP_MUL(0.5, c.move_size), (double)r.raw()), 4294967295.0);
    for (int k = 0; k < 3; k++)
      if (r.raw() % 2 == 0) d.v[k] = -d.v[k];
    d.kind = CG_COM;
  } else if (move_type == 2) {                                         // Pivot, molecule.cc:155-237
    int pivot = (int)floor(PP_DIV(PP_MUL((double)len, (double)r.raw()), 4294967295.0));
    if (pivot == len) pivot--;
    d.i0 = pivot;
    d.s = PP_DIV(PP_MUL(c.move_size, (double)r.raw()), 4294967295.0);
    d.n_rows = len - 1;
    d.kind = CG_PIVOT;
  } else if (move_type == 3) {                                         // Crankshaft, molecule.cc:239-247
    const int first = (int)floor(PP_DIV(PP_MUL((double)(len - 2), (double)r.raw()), 4294967295.0));
    const int last = (int)floor(PP_DIV(PP_MUL((double)(len - first), (double)r.raw()), 4294967295.0)) + first;
    d.i0 = first;
    d.i1 = last;
    d.s = PP_MUL(c.move_size, PP_SUB(PP_MUL(PP_MUL(r.uniform(), 2.0), CG_M_PI), CG_M_PI));
    d.kind = CG_CRANK;
  } else {                                                             // RandomReptation, molecule.cc:268-312
    d.s = cg_varied_bond(r, c.bond_len, c.vary_bond);
    d.i0 = (r.raw() % 2 == 0) ? -1 : 1;
    cg_rand_sphere(r, d.v);
    d.vlen = cg_vec_len(d.v);
    d.kind = CG_REPT;
  }
  return r.used;
}

// Pivot rows, serial form (one thread; also the only form when the bond length varies, because the extra draw
// behind every accepted pair shifts the pairing of everything that follows): rows [first, n_rows) are drawn until
// `budget` draws are used up.  Returns the draws consumed; *next_row is where a later call continues.
PP_HD int cg_pivot_rows_serial(const PgMt& mt, const CgConfig& c, int first, int n_rows, int budget, double* rows4, int* next_row) {
  CgCursor r{&mt, 0, nullptr, nullptr, 0};
  int row = first;
  while (row < n_rows && r.used < budget) {
    double v[3];
    cg_rand_sphere(r, v);
    const double bl = cg_varied_bond(r, c.bond_len, c.vary_bond);
    rows4[4 * row] = v[0]; rows4[4 * row + 1] = v[1]; rows4[4 * row + 2] = v[2]; rows4[4 * row + 3] = bl;
    row++;
  }
  *next_row = row;
  return r.used;
}

#endif  // PLUM_B200_PG_CHAIN_GEN_H_
