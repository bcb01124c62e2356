// Host-side Metropolis loop over the C ABI, for bench.py's end-to-end leg.
//
// This is the *caller's* side of the boundary written natively (C++), so that the
// end-to-end number measures pg_delta_e / pg_commit with host buffers and not a
// Python interpreter.  It proposes moves with the reference's move definitions
//   bead translation   src/molecules/molecule.cc:103-133  (|d| = 3*move_size, random direction)
//   COM translation    src/molecules/molecule.cc:135-151
//   pivot              src/molecules/molecule.cc:153-237
//   reptation          src/molecules/molecule.cc:267-312
// picks molecule and move type like Simulation::TranslationalMove
// (src/simulation/simulation.cc:241-355) and applies the same Metropolis rule
// (:324-332, no variate drawn when dE >= 1e8).  It is benchmark harness code: the
// product's real driver is Plum's own (bin/plum_gpu).  Every proposal is recorded so
// bench.py can replay the identical sequence device-resident (pg_replay_run).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "plum_b200.h"

namespace {

struct Ctx {
  pg_engine* eng;
  int n_mol, phantom;
  std::vector<int> mol_first;
  std::vector<double> pos;   // host copy of accepted coordinates [n][3] (the driver's Bead::current_pos)
  std::vector<int> chains, ions;
  double box[3];
  double beta, move_size, bond_len;
  double prob[5];
  std::mt19937 rng;
  // state of the proposal in flight and of the current run
  std::vector<double> trial;
  std::vector<uint8_t> moved;
  int cur_mol = -1, cur_f = 0, cur_len = 0;
  int done = 0, used = 0, acc_count = 0;
  double evals = 0, flops = 0;
  int32_t *rec_mol = nullptr, *rec_off = nullptr;
  double *rec_u = nullptr, *rec_dE = nullptr, *rec_trial = nullptr;
  uint8_t *rec_acc = nullptr, *rec_moved = nullptr;
  int rec_cap_beads = 0;
};

double uni(std::mt19937& g) { return (double)g() / g.max(); }

// src/utilities/misc.cc:95-109
void rand_sphere(double v[3], std::mt19937& g) {
  double rs = 2, r1 = 0, r2 = 0;
  while (rs > 1) {
    r1 = 1 - 2 * uni(g);
    r2 = 1 - 2 * uni(g);
    rs = r1 * r1 + r2 * r2;
  }
  double ranh = 2 * std::sqrt(1 - rs);
  v[0] = r1 * ranh;
  v[1] = r2 * ranh;
  v[2] = 1 - 2 * rs;
}

// One proposal with the reference's move definitions; false when the system has nothing to move.
bool propose(Ctx* c) {
  std::mt19937& g = c->rng;
  {
    // molecule + move type, simulation.cc:247-276
    int wc = (int)std::floor(uni(g) * (double)c->chains.size());
    int wi = (int)std::floor(uni(g) * (double)c->ions.size());
    if (wc == (int)c->chains.size()) wc--;
    if (wi == (int)c->ions.size()) wi--;
    int move_type = 0;
    double r = uni(g), cum = c->prob[0];
    while (cum < r && move_type < 4) { move_type++; cum += c->prob[move_type]; }
    int mol;
    if (move_type == 0 && !c->ions.empty()) mol = c->ions[wi];
    else if (!c->chains.empty()) { mol = c->chains[wc]; if (move_type == 0) move_type = 1; }
    else return false;
    const int f = c->mol_first[mol], len = c->mol_first[mol + 1] - f;
    c->trial.assign(c->pos.begin() + 3 * f, c->pos.begin() + 3 * (f + len));
    c->moved.assign(len, 1);
    c->cur_mol = mol; c->cur_f = f; c->cur_len = len;
    double* T = c->trial.data();
    if (move_type == 0) {
      double v[3];
      rand_sphere(v, g);
      double vl = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      if (vl > 0) vl = 3 * c->move_size / vl;
      for (int a = 0; a < 3; a++) T[a] += vl * v[a];
    } else if (move_type == 1) {
      double v[3] = {0.5 * c->move_size * uni(g), 0.5 * c->move_size * uni(g), 0.5 * c->move_size * uni(g)};
      for (int a = 0; a < 3; a++) if (g() % 2 == 0) v[a] = -v[a];
      for (int i = 0; i < len; i++) for (int a = 0; a < 3; a++) T[3 * i + a] += v[a];
    } else if (move_type == 2 || move_type == 3) {
      // pivot (crankshaft has probability 0 in every shipped run.in; it falls back to a pivot here)
      int pivot = (int)std::floor(len * uni(g));
      if (pivot == len) pivot--;
      double ms = c->move_size * uni(g);
      for (int dir = 0; dir < 2; dir++) {
        int i = dir == 0 ? pivot + 1 : pivot - 1;
        while (dir == 0 ? i < len : i >= 0) {
          int prev = dir == 0 ? i - 1 : i + 1;
          double v[3];
          rand_sphere(v, g);
          double dx[3], n2 = 0;
          for (int a = 0; a < 3; a++) { dx[a] = T[3 * i + a] + ms * v[a] - T[3 * prev + a]; n2 += dx[a] * dx[a]; }
          double norm = c->bond_len / std::sqrt(n2);
          double m[3];
          for (int a = 0; a < 3; a++) m[a] = norm * dx[a] + T[3 * prev + a] - T[3 * i + a];
          if (dir == 0) { for (int j = i; j < len; j++) for (int a = 0; a < 3; a++) T[3 * j + a] += m[a]; }
          else { for (int j = i; j >= 0; j--) for (int a = 0; a < 3; a++) T[3 * j + a] += m[a]; }
          i += dir == 0 ? 1 : -1;
        }
      }
    } else {
      // reptation: every bead takes its neighbour's place, a new end bead is grown
      int direction = 1, begin = 0, end = len - 1;
      if (g() % 2 == 0) { direction = -1; begin = len - 1; end = 0; }
      const double* C = c->pos.data() + 3 * f;
      for (int i = begin; i != end; i += direction) for (int a = 0; a < 3; a++) T[3 * i + a] = C[3 * (i + direction) + a];
      double v[3];
      rand_sphere(v, g);
      double vl = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for (int a = 0; a < 3; a++) T[3 * end + a] = C[3 * end + a] + c->bond_len * v[a] / vl;
    }
  }
  return true;
}

// Metropolis rule (simulation.cc:324-332) + bookkeeping for a proposal whose dE is known.
int finish(Ctx* c, const pg_delta& d) {
  std::mt19937& g = c->rng;
  const int N = c->mol_first[c->n_mol];
  const int f = c->cur_f, len = c->cur_len, it = c->done;
  const double* T = c->trial.data();
  bool accept = false;
  double u = -1.0;
  if (d.dE < PG_VERY_LARGE_ENERGY) {
    u = uni(g);
    accept = u < std::exp(-c->beta * d.dE);
  }
  int rc = pg_commit(c->eng, accept ? 1 : 0);
  if (rc) return rc;
  if (accept) { std::memcpy(c->pos.data() + 3 * f, T, sizeof(double) * 3 * len); c->acc_count++; }
  // reference-unit work of this move: pair-dE evaluations and algorithmic flops
  double n_intra = (len > 1) ? 0.5 * len * (len - 1) : 0.0;
  double ev = (double)len * (double)(N - len) + n_intra;
  c->evals += ev;
  c->flops += 2.0 * ev * 36.0;   // + reciprocal-space terms added by the caller (needs charges)
  if (c->rec_mol) {
    if (c->used + len > c->rec_cap_beads) return PG_ERR_CAPACITY;
    c->rec_mol[it] = c->cur_mol; c->rec_off[it] = c->used; c->rec_u[it] = (u < 0) ? 2.0 : u; c->rec_dE[it] = d.dE;
    c->rec_acc[it] = accept ? 1 : 0;
    std::memcpy(c->rec_trial + 3 * (size_t)c->used, T, sizeof(double) * 3 * len);
    std::memcpy(c->rec_moved + c->used, c->moved.data(), len);
    c->used += len;
  }
  c->done++;
  return 0;
}

}  // namespace

extern "C" {

void* pb_create(pg_engine* eng, int n_mol, const int32_t* mol_first, const double* xyz, const double* box, double beta,
                double move_size, double bond_len, const double* prob5, int phantom, unsigned seed) {
  Ctx* c = new Ctx();
  c->eng = eng;
  c->n_mol = n_mol;
  c->phantom = phantom;
  c->mol_first.assign(mol_first, mol_first + n_mol + 1);
  c->pos.assign(xyz, xyz + 3 * (size_t)mol_first[n_mol]);
  for (int i = 0; i < 3; i++) c->box[i] = box[i];
  c->beta = beta;
  c->move_size = move_size;
  c->bond_len = bond_len;
  for (int i = 0; i < 5; i++) c->prob[i] = prob5[i];
  for (int m = phantom; m < n_mol; m++) {
    if (mol_first[m + 1] - mol_first[m] > 1) c->chains.push_back(m);
    else c->ions.push_back(m);
  }
  c->rng.seed(seed);
  return c;
}

void pb_destroy(void* p) { delete static_cast<Ctx*>(p); }

// Recording arrays of the next run (may all be NULL): packed trial coordinates / moved flags of every
// proposal, capacity rec_cap_beads beads.
void pb_set_records(void* p, int32_t* rec_mol, int32_t* rec_off, double* rec_u, double* rec_dE, uint8_t* rec_acc,
                    double* rec_trial, uint8_t* rec_moved, int rec_cap_beads) {
  Ctx* c = static_cast<Ctx*>(p);
  c->rec_mol = rec_mol; c->rec_off = rec_off; c->rec_u = rec_u; c->rec_dE = rec_dE; c->rec_acc = rec_acc;
  c->rec_trial = rec_trial; c->rec_moved = rec_moved; c->rec_cap_beads = rec_cap_beads;
}

void pb_stats(void* p, int32_t* rec_beads, double* pair_evals, double* alg_flops, int32_t* n_accept) {
  Ctx* c = static_cast<Ctx*>(p);
  if (rec_beads) *rec_beads = c->used;
  if (pair_evals) *pair_evals = c->evals;
  if (alg_flops) *alg_flops = c->flops;
  if (n_accept) *n_accept = c->acc_count;
}

// n_moves Metropolis steps on each of n_ctx independent replicas, driven by THIS thread: the round
// trips of the replicas overlap (pg_delta_e_begin / pg_delta_e_poll), each chain stays sequential.
// Returns 0 or a pg_status.
int pb_run_multi(void** ps, int n_ctx, int n_moves, double* wall_seconds) {
  std::vector<Ctx*> cs(n_ctx);
  for (int i = 0; i < n_ctx; i++) {
    cs[i] = static_cast<Ctx*>(ps[i]);
    cs[i]->done = 0; cs[i]->used = 0; cs[i]->acc_count = 0; cs[i]->evals = 0; cs[i]->flops = 0;
  }
  auto t0 = std::chrono::steady_clock::now();
  std::vector<char> flying(n_ctx, 0);
  int live = 0;
  for (int i = 0; i < n_ctx; i++) {
    if (n_moves <= 0 || !propose(cs[i])) continue;
    int rc = pg_delta_e_begin(cs[i]->eng, cs[i]->cur_mol, cs[i]->trial.data(), cs[i]->moved.data());
    if (rc) return rc;
    flying[i] = 1; live++;
  }
  pg_delta d;
  while (live > 0) {
    for (int i = 0; i < n_ctx; i++) {
      if (!flying[i]) continue;
      Ctx* c = cs[i];
      int rc = (n_ctx == 1) ? 0 : pg_delta_e_poll(c->eng, &d);
      if (rc == 1) continue;
      if (rc < 0) return rc;
      if (n_ctx == 1) {   // nothing to overlap with: block
        do { rc = pg_delta_e_poll(c->eng, &d); } while (rc == 1);
        if (rc < 0) return rc;
      }
      rc = finish(c, d);
      if (rc) return rc;
      if (c->done < n_moves && propose(c)) {
        rc = pg_delta_e_begin(c->eng, c->cur_mol, c->trial.data(), c->moved.data());
        if (rc) return rc;
      } else {
        flying[i] = 0; live--;
      }
    }
  }
  // drain the streams: the last commits must have landed before the clock stops
  for (int i = 0; i < n_ctx; i++) {
    pg_totals tot;
    int rc = pg_get_totals(cs[i]->eng, &tot);
    if (rc) return rc;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (wall_seconds) *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

// Single-replica form with the records passed in one call.
int pb_run(void* p, int n_moves, int32_t* rec_mol, int32_t* rec_off, double* rec_u, double* rec_dE, uint8_t* rec_acc,
           double* rec_trial, uint8_t* rec_moved, int rec_cap_beads, int32_t* rec_beads, double* wall_seconds,
           double* pair_evals, double* alg_flops, int32_t* n_accept, int K_full) {
  (void)K_full;
  pb_set_records(p, rec_mol, rec_off, rec_u, rec_dE, rec_acc, rec_trial, rec_moved, rec_cap_beads);
  int rc = pb_run_multi(&p, 1, n_moves, wall_seconds);
  pb_stats(p, rec_beads, pair_evals, alg_flops, n_accept);
  return rc;
}

}  // extern "C"
