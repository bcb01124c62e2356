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
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <sstream>
#include <vector>

#include "plum_b200.h"
#include "mc_propose.h"
#include "mt_state.h"

namespace {

struct Ctx {
  pg_engine* eng;
  int n_mol, phantom;
  std::vector<int> mol_first;
  std::vector<double> pos;   // host copy of accepted coordinates [n][3] (the driver's Bead::current_pos)
  plum_mc::Proposer prop;
  plum_mc::Batch one;        // the step in flight on the per-move path
  plum_mc::Batch batch[2];   // batched path: the batch on the device and the one generated ahead
  plum_mc::BatchSizer sizer;
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
  int pivot_mode = 0;        // pg_chain_config.pivot_mode of the chain path
  std::mt19937 rng_handed;   // the generator as the device handed it back last time
  bool rng_resident = false; // ... and whether that hand-back happened (the device then still holds the state)
};

// One proposal, drawn in the reference's order (mc_propose.h) and built on the host; false when the
// system has nothing to move or the step is one the harness does not drive (GC).
bool propose(Ctx* c) {
  for (;;) {
    if (c->prop.generate(c->rng, 1, c->one) != 1) return false;
    if (c->one.kind[0] >= 0) break;   // a step that attempts nothing: next step
  }
  const pg_move_desc& d = c->one.moves[0];
  const int mol = d.mol, f = c->mol_first[mol], len = c->mol_first[mol + 1] - f;
  c->trial.resize(3 * (size_t)len);
  c->moved.assign(len, 1);
  if (d.kind == PG_MOVE_BEAD) for (int i = 1; i < len; i++) c->moved[i] = 0;
  if (d.kind == PG_MOVE_CRANKSHAFT && std::min(d.rv_offset, len - 1) - d.i0 > 1)   // molecule.cc:251-253 (none moved: all, at rest)
    for (int i = 0; i < len; i++) c->moved[i] = (i > d.i0 && i < std::min(d.rv_offset, len - 1)) ? 1 : 0;
  plum_mc::Proposer::apply(d, c->one.rvec.data(), len, c->pos.data() + 3 * (size_t)f, c->trial.data());
  c->cur_mol = mol; c->cur_f = f; c->cur_len = len;
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
    u = c->one.moves[0].u;   // drawn by the generator right behind the proposal, as the reference does
    accept = u < std::exp(-c->beta * d.dE);
  } else {
    c->one.rewind_after_overlap(g, 1);   // simulation.cc:327-332: no variate is drawn
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
    c->rec_mol[it] = c->cur_mol; c->rec_off[it] = c->used; c->rec_u[it] = (u < 0) ? 2.0 : u; c->rec_dE[it] = d.dE;
    c->rec_acc[it] = accept ? 1 : 0;
    if (c->rec_cap_beads > 0) {   // trial coordinates are recorded only when the caller gave room for them
      if (c->used + len > c->rec_cap_beads) return PG_ERR_CAPACITY;
      std::memcpy(c->rec_trial + 3 * (size_t)c->used, T, sizeof(double) * 3 * len);
      std::memcpy(c->rec_moved + c->used, c->moved.data(), len);
      c->used += len;
    }
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
  plum_mc::Config cfg;
  cfg.phantom = phantom;
  for (int m = 0; m < n_mol; m++) cfg.mol_len.push_back(mol_first[m + 1] - mol_first[m]);
  cfg.move_size = move_size;
  for (int i = 0; i < 5; i++) cfg.move_prob[i] = prob5[i];
  cfg.bond_len = bond_len;
  cfg.vary_bond = false;
  cfg.gc_freq = 0;
  c->prop.configure(cfg);
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


// Chain path: 0 builds pivot arms in the reference's operation order, 1 as prefix sums (include/plum_b200.h).
void pb_set_pivot_mode(void* p, int mode) { static_cast<Ctx*>(p)->pivot_mode = mode; }

// Spring-bond variant: the generators vary the bond length by +-10 % (simulation.cc:293-296).
void pb_set_vary_bond(void* p, int vary) {
  Ctx* c = static_cast<Ctx*>(p);
  plum_mc::Config cfg = c->prop.config();
  cfg.vary_bond = vary != 0;
  c->prop.configure(cfg);
}

// Host copy of the accepted coordinates [n][3] (after pb_run_mc: downloaded from the device).
void pb_positions(void* p, double* xyz) {
  Ctx* c = static_cast<Ctx*>(p);
  std::memcpy(xyz, c->pos.data(), sizeof(double) * c->pos.size());
}

// The same Markov chains through the BATCHED path (pg_mc_upload / pg_mc_begin / pg_mc_end): this thread draws
// each replica's random stream a batch ahead and the device does the rest — trial coordinates, dE, Metropolis
// test, commit.  While a batch runs, the next one is generated; a batch that stopped at a dE >= 1e8 step
// (no acceptance draw in the reference) makes the generator rewind to that step.  The records (mol, u, dE,
// accept) are filled like pb_run_multi's; trial coordinates are not recorded (they never leave the device).
int pb_run_mc(void** ps, int n_ctx, int n_moves, int batch_moves, double* wall_seconds) {
  std::vector<Ctx*> cs(n_ctx);
  for (int i = 0; i < n_ctx; i++) {
    cs[i] = static_cast<Ctx*>(ps[i]);
    cs[i]->done = 0; cs[i]->used = 0; cs[i]->acc_count = 0; cs[i]->evals = 0; cs[i]->flops = 0;
  }
  if (batch_moves <= 0) batch_moves = 1024;
  auto t0 = std::chrono::steady_clock::now();
  std::vector<int> cur(n_ctx, 0), ahead(n_ctx, 0), flying(n_ctx, 0);
  std::vector<std::mt19937> rng_after(n_ctx);   // generator state behind the batch that is on the device
  std::vector<double> dE(batch_moves + 8);
  std::vector<uint8_t> acc(batch_moves + 8);
  auto launch = [&](int i) -> int {
    Ctx* c = cs[i];
    plum_mc::Batch& b = c->batch[cur[i]];
    if ((int)b.moves.size() > n_moves - c->done) return PG_ERR_STATE;
    if (b.moves.empty()) return 1;   // nothing (left) to do
    int rc = pg_mc_upload(c->eng, (int)b.moves.size(), b.moves.data(), (int)(b.rvec.size() / 4), b.rvec.data());
    if (rc) return rc;
    rc = pg_mc_begin(c->eng, 0, (int)b.moves.size());
    if (rc) return rc;
    flying[i] = 1;
    return 0;
  };
  auto generate = [&](int i, int slot, int budget) {
    // a batch of `budget` MOVES at most (steps that attempt nothing do not count)
    Ctx* c = cs[i];
    c->prop.generate(c->rng, budget, c->batch[slot]);
  };
  int live = 0;
  for (int i = 0; i < n_ctx; i++) {
    cs[i]->sizer.reset(batch_moves);
    generate(i, 0, std::min(cs[i]->sizer.next(), n_moves));
    int rc = launch(i);
    if (rc < 0) return rc;
    if (rc == 0) live++;
  }
  // generate ahead while the first batches run
  for (int i = 0; i < n_ctx; i++)
    if (flying[i]) {
      Ctx* c = cs[i];
      rng_after[i] = c->rng;
      const int left = n_moves - c->done - (int)c->batch[cur[i]].moves.size();
      generate(i, cur[i] ^ 1, std::max(0, std::min(c->sizer.next(), left)));
      ahead[i] = 1;
    }
  while (live > 0) {
    for (int i = 0; i < n_ctx; i++) {
      if (!flying[i]) continue;
      Ctx* c = cs[i];
      plum_mc::Batch& b = c->batch[cur[i]];
      int n_done = 0;
      int rc = pg_mc_end(c->eng, dE.data(), acc.data(), &n_done, nullptr);
      if (rc) return rc;
      flying[i] = 0;
      const int N = c->mol_first[c->n_mol];
      for (int m = 0; m < n_done; m++) {
        const pg_move_desc& d = b.moves[m];
        const int len = c->mol_first[d.mol + 1] - c->mol_first[d.mol];
        const double n_intra = (len > 1) ? 0.5 * len * (len - 1) : 0.0;
        const double ev = (double)len * (double)(N - len) + n_intra;
        c->evals += ev;
        c->flops += 2.0 * ev * 36.0;
        if (acc[m]) c->acc_count++;
        if (c->rec_mol) {
          const int it = c->done + m;
          c->rec_mol[it] = d.mol; c->rec_off[it] = 0;
          c->rec_u[it] = (dE[m] >= PG_VERY_LARGE_ENERGY) ? 2.0 : d.u;
          c->rec_dE[it] = dE[m]; c->rec_acc[it] = acc[m];
        }
      }
      c->done += n_done;
      const bool stopped = n_done > 0 && dE[n_done - 1] >= PG_VERY_LARGE_ENERGY;
      if (!stopped && n_done < (int)b.moves.size()) return PG_ERR_STATE;
      if (stopped) {
        // the last done step drew no acceptance variate (also when it was the last step of the batch):
        // everything generated behind it is void
        b.rewind_after_overlap(c->rng, n_done);
        ahead[i] = 0;
      }
      c->sizer.update(n_done, stopped);
      if (c->done >= n_moves) { live--; continue; }
      if (ahead[i]) {
        cur[i] ^= 1;
      } else {
        generate(i, cur[i], std::min(c->sizer.next(), n_moves - c->done));
      }
      rc = launch(i);
      if (rc < 0) return rc;
      if (rc == 1) { live--; continue; }
      rng_after[i] = c->rng;
      const int left = n_moves - c->done - (int)c->batch[cur[i]].moves.size();
      generate(i, cur[i] ^ 1, std::max(0, std::min(c->sizer.next(), left)));
      ahead[i] = 1;
    }
  }
  for (int i = 0; i < n_ctx; i++) {
    Ctx* c = cs[i];
    int rc = pg_download_positions(c->eng, c->pos.data());
    if (rc) return rc;
    // the generator ran one batch ahead: put it back behind the last step that was executed
    if (ahead[i]) c->rng = rng_after[i];
  }
  auto t1 = std::chrono::steady_clock::now();
  if (wall_seconds) *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

// ---- the device-resident chain (pg_chain_*): the host only carries the generator state across the boundary ----
namespace {

// std::mt19937 <-> (624 state words, position): the textual form operator<< / operator>> define.
void mt_export(const std::mt19937& g, uint32_t* state, int* pos) { plum_mt::export_state(g, state, pos); }
void mt_import(std::mt19937& g, const uint32_t* state, int pos) { plum_mt::import_state(g, state, pos); }

int chain_configure(Ctx* c, int cluster, int keep_trials) {

  const plum_mc::Config& pc = c->prop.config();
  pg_chain_config cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.phantom = pc.phantom; cfg.gc_freq = pc.gc_freq; cfg.vary_bond = pc.vary_bond ? 1 : 0;
  cfg.cluster_ctas = cluster; cfg.keep_trials = keep_trials; cfg.pivot_mode = c->pivot_mode;
  cfg.move_size = pc.move_size; cfg.bond_len = pc.bond_len;
  for (int i = 0; i < 5; i++) cfg.move_prob[i] = pc.move_prob[i];
  return pg_chain_configure(c->eng, &cfg);
}

// Books the records of n_steps finished steps of replica c; returns the moves among them.
int chain_book(Ctx* c, const pg_chain_step* st, int n_steps) {
  const int N = c->mol_first[c->n_mol];
  int moves = 0;
  for (int s = 0; s < n_steps; s++) {
    if (st[s].kind < 0) continue;
    const int mol = st[s].mol, len = c->mol_first[mol + 1] - c->mol_first[mol];
    const double n_intra = (len > 1) ? 0.5 * len * (len - 1) : 0.0;
    const double ev = (double)len * (double)(N - len) + n_intra;
    c->evals += ev;
    c->flops += 2.0 * ev * 36.0;
    if (st[s].accept) c->acc_count++;
    if (c->rec_mol) {
      const int it = c->done + moves;
      c->rec_mol[it] = mol; c->rec_off[it] = 0; c->rec_u[it] = 0.0; c->rec_dE[it] = st[s].dE; c->rec_acc[it] = st[s].accept;
    }
    moves++;
  }
  c->done += moves;
  return moves;
}

}  // namespace

// n_moves Metropolis steps of every replica through pg_chain_*: per batch the generator state goes down (2.5 KB), one
// launch runs `batch` steps of ALL replicas (one cluster each), and the step log (16 B per step), the generator state
// and — like ForceField::TranslationalBatch, whose driver reads Bead::current_pos afterwards — the accepted
// coordinates come back.  One host thread drives everything.  (Systems without grand-canonical steps.)
int pb_run_chain(void** ps, int n_ctx, int n_moves, int batch, int cluster, int keep_trials, int positions_every_batch,
                 double* wall_seconds) {
  std::vector<Ctx*> cs(n_ctx);
  std::vector<pg_engine*> engs(n_ctx);
  for (int i = 0; i < n_ctx; i++) {
    cs[i] = static_cast<Ctx*>(ps[i]);
    cs[i]->done = 0; cs[i]->used = 0; cs[i]->acc_count = 0; cs[i]->evals = 0; cs[i]->flops = 0;
    engs[i] = cs[i]->eng;
    int rc = chain_configure(cs[i], cluster, keep_trials);
    if (rc) return rc;
  }
  if (batch <= 0) batch = 1024;
  auto t0 = std::chrono::steady_clock::now();
  // (buffers of the call live across calls: a fresh 100 MB vector per step would cost more than the transfers)
  static thread_local std::vector<pg_chain_step> st;
  static thread_local std::vector<uint32_t> rng;
  static thread_local std::vector<uint8_t> up;
  static thread_local std::vector<double*> xyz;
  if (st.size() < (size_t)batch * n_ctx) st.resize((size_t)batch * n_ctx);
  rng.resize((size_t)625 * n_ctx); up.resize(n_ctx); xyz.resize(n_ctx);
  std::vector<int> n_done(n_ctx);
  int left = n_moves;
  while (left > 0) {
    const int b = std::min(batch, left);
    for (int i = 0; i < n_ctx; i++) {
      // the generator goes down only if the host has drawn from it since the device handed it back
      up[i] = !(cs[i]->rng_resident && cs[i]->rng == cs[i]->rng_handed);
      if (up[i]) {
        int pos = 0;
        mt_export(cs[i]->rng, rng.data() + (size_t)625 * i, &pos);
        rng[(size_t)625 * i + 624] = (uint32_t)pos;
      }
      xyz[i] = positions_every_batch ? cs[i]->pos.data() : nullptr;
    }
    int rc = pg_chain_run_multi_io(engs.data(), n_ctx, b, rng.data(), up.data(), st.data(), positions_every_batch ? xyz.data() : nullptr,
                                   n_done.data(), nullptr);
    if (rc) return rc;
    int moves0 = -1;
    for (int i = 0; i < n_ctx; i++) {
      if (n_done[i] != b) return PG_ERR_STATE;
      const int moves = chain_book(cs[i], st.data() + (size_t)b * i, b);
      if (i == 0) moves0 = moves;
      mt_import(cs[i]->rng, rng.data() + (size_t)625 * i, (int)rng[(size_t)625 * i + 624]);
      cs[i]->rng_handed = cs[i]->rng;
      cs[i]->rng_resident = true;
    }
    // every step of these systems attempts a move; the loop counts replica 0's
    left -= std::max(moves0, 1);
  }
  if (!positions_every_batch)
    for (int i = 0; i < n_ctx; i++) {
      int rc = pg_download_positions(engs[i], cs[i]->pos.data());
      if (rc) return rc;
    }
  auto t1 = std::chrono::steady_clock::now();
  if (wall_seconds) *wall_seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

// The same, device-resident: the generator state is already on the device (pb_chain_seed), nothing crosses the
// boundary inside the call but the launch itself; returns the CUDA-event time of the launches.
int pb_chain_seed(void** ps, int n_ctx, int cluster) {
  uint32_t state[624];
  int pos = 0;
  for (int i = 0; i < n_ctx; i++) {
    Ctx* c = static_cast<Ctx*>(ps[i]);
    int rc = chain_configure(c, cluster, 0);
    if (rc) return rc;
    mt_export(c->rng, state, &pos);
    rc = pg_chain_set_rng(c->eng, state, pos);
    if (rc) return rc;
  }
  return 0;
}
int pb_chain_resident(void** ps, int n_ctx, int n_steps, int book, float* elapsed_ms) {
  std::vector<pg_engine*> engs(n_ctx);
  for (int i = 0; i < n_ctx; i++) engs[i] = static_cast<Ctx*>(ps[i])->eng;
  std::vector<int> n_done(n_ctx);
  int rc = pg_chain_run_multi(engs.data(), n_ctx, n_steps, n_done.data(), elapsed_ms);
  if (rc) return rc;
  for (int i = 0; i < n_ctx; i++)
    if (n_done[i] != n_steps) return PG_ERR_STATE;
  if (book) {
    std::vector<pg_chain_step> st((size_t)n_steps);
    for (int i = 0; i < n_ctx; i++) {
      Ctx* c = static_cast<Ctx*>(ps[i]);
      c->done = 0; c->acc_count = 0; c->evals = 0; c->flops = 0;
      rc = pg_chain_steps(engs[i], 0, n_steps, st.data());
      if (rc) return rc;
      chain_book(c, st.data(), n_steps);
    }
  }
  return 0;
}

// ---- the generator alone (CPU tests pin it to the reference's own trial coordinates) ----
struct Gen {
  plum_mc::Proposer prop;
  plum_mc::Batch one;
  std::mt19937 rng;
};

void* pmc_create(int n_mol, const int32_t* mol_len, int phantom, double move_size, const double* prob5, double bond_len,
                 int vary_bond, int gc_freq, unsigned seed) {
  Gen* g = new Gen();
  plum_mc::Config cfg;
  cfg.phantom = phantom;
  cfg.mol_len.assign(mol_len, mol_len + n_mol);
  cfg.move_size = move_size;
  for (int i = 0; i < 5; i++) cfg.move_prob[i] = prob5[i];
  cfg.bond_len = bond_len;
  cfg.vary_bond = vary_bond != 0;
  cfg.gc_freq = gc_freq;
  g->prop.configure(cfg);
  g->rng.seed(seed);
  return g;
}
void pmc_destroy(void* p) { delete static_cast<Gen*>(p); }
// One step.  Returns the move kind (descriptor in *d, pivot rows in rvec, at most rvec_cap_rows), -1 for a step
// that attempts nothing, -2 when the next step is a GC step, -3 for a move the device does not offer (the
// generator then stands in front of that step).  The acceptance variate is already drawn (d->u): call
// pmc_no_accept_draw when the step's dE turned out >= 1e8.
int pmc_next(void* p, pg_move_desc* d, double* rvec, int rvec_cap_rows, int* n_rows) {
  Gen* g = static_cast<Gen*>(p);
  if (g->prop.generate(g->rng, 1, g->one) != 1) return g->one.stop == plum_mc::STOP_GC ? -2 : -3;
  if (n_rows) *n_rows = 0;
  if (g->one.kind[0] < 0) return -1;
  *d = g->one.moves[0];
  const int rows = (int)(g->one.rvec.size() / 4);
  if (rows > rvec_cap_rows) return -4;
  if (rows) std::memcpy(rvec, g->one.rvec.data(), sizeof(double) * 4 * (size_t)rows);
  if (n_rows) *n_rows = rows;
  return d->kind;
}
void pmc_no_accept_draw(void* p) {
  Gen* g = static_cast<Gen*>(p);
  g->one.rewind_after_overlap(g->rng, 1);
}
// Draws the generator consumed so far cannot be observed from outside; tests compare the NEXT raw output instead.
unsigned pmc_peek_raw(void* p) {
  Gen* g = static_cast<Gen*>(p);
  std::mt19937 copy = g->rng;
  return (unsigned)copy();
}
void pmc_apply(const pg_move_desc* d, const double* rvec, int len, const double* cur, double* out) {
  plum_mc::Proposer::apply(*d, rvec, len, cur, out);
}

}  // extern "C"
