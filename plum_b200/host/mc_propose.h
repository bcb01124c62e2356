// plum_b200 — host half of the batched Markov-chain steps (include/plum_b200.h, pg_mc_*).
//
// Walks the caller's std::mt19937 in EXACTLY the draw order of the reference's translational
// step — Simulation::Run / TranslationalMove (src/simulation/simulation.cc:216-355), the move
// generators (src/molecules/molecule.cc:103-312) and randSphere (src/utilities/misc.cc:95-109) —
// and records, per step, the coordinate-independent part of the proposal as a pg_move_desc.
// Nothing here looks at coordinates: which molecule, which move and which random vectors a step
// uses is a function of the random stream alone, so the host can run ahead of the device by a
// whole batch.  The one coupling is the acceptance draw, which the reference skips when
// dE >= 1e8 (simulation.cc:327-332): `pos_accept` remembers where each step's acceptance draw
// sits in the stream so that the caller can rewind to it (Batch::rewind) when the device reports
// such a step.
//
// apply() is the host version of the device's k_propose (same arithmetic header), used by the
// per-move path, the tests and the CPU pin against the reference's own trial coordinates.
//
// C++11, header only.  Compile with -ffp-contract=off (see pg_propose_math.h).
#ifndef PLUM_B200_MC_PROPOSE_H_
#define PLUM_B200_MC_PROPOSE_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <vector>

#include "plum_b200.h"
#include "pg_propose_math.h"

namespace plum_mc {

struct Config {
  int phantom = 0;                 // molecules [0, phantom) are never moved (simulation.cc:255)
  std::vector<int> mol_len;        // beads per molecule, all molecules
  double move_size = 0.0;
  double move_prob[5] = {0, 0, 0, 0, 0};   // bead, COM, pivot, crankshaft, reptation (simulation.cc:46-50)
  double bond_len = 0.0;           // RigidBondLen() or EqBondLen() (simulation.cc:288-296)
  bool vary_bond = false;          // UseBondPot()
  int gc_freq = 0;                 // > 0: UseGC(), a step is a GC step when rand_num % gc_freq == 0 (simulation.cc:221-224)
};

enum StopReason { STOP_FULL = 0, STOP_GC = 1, STOP_UNSUPPORTED = 2 };

struct Batch {
  std::vector<pg_move_desc> moves;      // the attempted moves, in step order
  std::vector<double> rvec;             // [n][4] pivot rows
  std::vector<int> kind;                // per STEP: PG_MOVE_* or -1 (a step that attempts nothing)
  std::vector<int> move_of_step;        // per STEP: index into moves or -1
  std::vector<int> step_of_move;        // per move
  std::vector<uint64_t> pos_accept;     // per move: draws consumed since the batch began, up to its acceptance draw
  std::vector<uint64_t> pos_step_end;   // per STEP: draws consumed once the step is over (acceptance draw included)
  std::mt19937 start;                   // generator state when the batch began
  StopReason stop = STOP_FULL;
  void clear() { moves.clear(); rvec.clear(); kind.clear(); move_of_step.clear(); step_of_move.clear(); pos_accept.clear(); pos_step_end.clear(); stop = STOP_FULL; }
  int n_steps() const { return (int)kind.size(); }
  // Puts g where the reference's generator stands after `moves_done` moves of which the LAST one returned
  // dE >= 1e8 (its acceptance variate was never drawn); returns the number of STEPS that are over.
  int rewind_after_overlap(std::mt19937& g, int moves_done) const {
    g = start;
    g.discard(pos_accept[moves_done - 1]);
    return step_of_move[moves_done - 1] + 1;
  }
};

// How many steps to hand to the device at once.  A batch that stops early (a dE >= 1e8 step) leaves the kernels
// queued behind the stop as no-ops that still cost w ~ 3 us of device time each (measured, tools/mc_batch_probe.py),
// and every batch costs C ~ 50 us of upload / launch / completion latency.  With a stop probability p per step the
// overhead per done step, ~ C/B + w p B/2 for B << 1/p, is smallest at B = sqrt(2C/(w p)) ~ 6 sqrt(run), `run` = 1/p
// being the smoothed number of steps between stops.  `worthwhile()` is false where even that optimum costs more
// than the ~3 us round trip per step it saves (p above ~1.5 %: dense or walled systems, long pivots).
struct BatchSizer {
  int cap = 1024;
  double run = 1024.0;
  void reset(int max_batch) { cap = max_batch < 1 ? 1 : max_batch; run = 512.0; }
  void update(int n_done, bool stopped) {
    if (stopped) run = 0.8 * run + 0.2 * n_done;
    else run = 0.8 * run + 0.2 * std::max(run, 2.0 * n_done + 8.0);   // no stop seen: at least this long, probably longer
    if (run > 4096.0) run = 4096.0;
  }
  int next() const {
    int b = (int)(6.0 * std::sqrt(run));
    return b < 8 ? 8 : (b > cap ? cap : b);
  }
  bool worthwhile() const { return run >= 64.0; }
};

class Proposer {
 public:
  void configure(const Config& c) {
    cfg = c;
    chains.clear(); ions.clear();
    for (int i = c.phantom; i < (int)c.mol_len.size(); i++) (c.mol_len[i] > 1 ? chains : ions).push_back(i);
  }
  const Config& config() const { return cfg; }

  // Generates steps until `max_steps` are there, or the next step is a GC step / a move the device does not offer
  // (then g is left exactly where it was before that step began).  Returns the number of steps generated.
  int generate(std::mt19937& g, int max_steps, Batch& b) const {
    b.clear();
    b.start = g;
    Counter r{&g, 0};
    while (b.n_steps() < max_steps) {
      const std::mt19937 before = g;
      const uint64_t n_before = r.n;
      if (cfg.gc_freq > 0) {
        const int rand_num = (int)(uint32_t)r();               // simulation.cc:221
        if (rand_num % cfg.gc_freq == 0) { g = before; r.n = n_before; b.stop = STOP_GC; break; }
      } else {
        (void)r();                                             // drawn whether or not GC is on
      }
      pg_move_desc d;
      const int kind = one_step(r, d, b.rvec);
      b.kind.push_back(kind);
      if (kind >= 0) {
        b.pos_accept.push_back(r.n);
        d.u = r.uniform();                                     // simulation.cc:331
        b.move_of_step.push_back((int)b.moves.size());
        b.step_of_move.push_back(b.n_steps() - 1);
        b.moves.push_back(d);
      } else {
        b.move_of_step.push_back(-1);
      }
      b.pos_step_end.push_back(r.n);
    }
    return b.n_steps();
  }

  // Host version of k_propose: trial coordinates of move d from the molecule's current coordinates.
  static void apply(const pg_move_desc& d, const double* rvec, int len, const double* cur /* [len][3] */, double* out) {
    for (int i = 0; i < 3 * len; i++) out[i] = cur[i];
    switch (d.kind) {
      case PG_MOVE_BEAD:
        for (int k = 0; k < 3; k++) out[k] = pp_bead_translate(cur[k], d.s, d.v[k]);
        break;
      case PG_MOVE_COM:
        for (int i = 0; i < len; i++)
          for (int k = 0; k < 3; k++) out[3 * i + k] = pp_com_translate(cur[3 * i + k], d.v[k]);
        break;
      case PG_MOVE_REPTATION: {
        const int dir = d.i0, begin = dir > 0 ? 0 : len - 1, end = dir > 0 ? len - 1 : 0;
        for (int i = begin; i != end; i += dir)
          for (int k = 0; k < 3; k++) out[3 * i + k] = cur[3 * (i + dir) + k];
        for (int k = 0; k < 3; k++) out[3 * end + k] = pp_reptation_end(cur[3 * end + k], d.s, d.v[k], d.vlen);
        break;
      }
      case PG_MOVE_PIVOT: {
        const int p = d.i0;
        const double* row = rvec + 4 * (size_t)d.rv_offset;
        for (int i = p + 1; i < len; i++, row += 4) {          // molecule.cc:170-198
          double m[3];
          pp_pivot_step(out + 3 * (i - 1), out + 3 * i, d.s, row, row[3], m);
          for (int j = i; j < len; j++)
            for (int k = 0; k < 3; k++) out[3 * j + k] = PP_ADD(out[3 * j + k], m[k]);
        }
        for (int i = p - 1; i >= 0; i--, row += 4) {           // molecule.cc:203-231
          double m[3];
          pp_pivot_step(out + 3 * (i + 1), out + 3 * i, d.s, row, row[3], m);
          for (int j = i; j >= 0; j--)
            for (int k = 0; k < 3; k++) out[3 * j + k] = PP_ADD(out[3 * j + k], m[k]);
        }
        break;
      }
      case PG_MOVE_CRANKSHAFT: {                             // molecule.cc:242-263
        const int first = d.i0, last = d.rv_offset < len ? d.rv_offset : len - 1;
        double rot[9];
        pp_crank_matrix(cur + 3 * first, cur + 3 * last, d.v[0], d.v[1], rot);
        for (int i = first + 1; i < last; i++) pp_crank_apply(rot, cur + 3 * first, cur + 3 * i, out + 3 * i);
        break;
      }
      default: break;
    }
  }

 private:
  struct Counter {
    std::mt19937* g;
    uint64_t n;
    uint32_t operator()() { n++; return (uint32_t)(*g)(); }
    // (double)rand_gen() / rand_gen.max()
    double uniform() { return (double)(*this)() / 4294967295.0; }
  };

  // randSphere, misc.cc:95-109.
  static void rand_sphere(Counter& r, double vec[3]) {
    double rand_square = 2, r1 = 0, r2 = 0;
    while (rand_square > 1) {
      r1 = 1 - 2 * r.uniform();
      r2 = 1 - 2 * r.uniform();
      rand_square = r1 * r1 + r2 * r2;
    }
    const double ranh = 2 * std::sqrt(1 - rand_square);
    vec[0] = r1 * ranh;
    vec[1] = r2 * ranh;
    vec[2] = (1 - 2 * rand_square);
  }

  double varied_bond(Counter& r) const {
    double bond_len = cfg.bond_len;
    if (cfg.vary_bond) bond_len += (cfg.bond_len / 5.0) * (r.uniform() - 0.5);   // molecule.cc:184-187
    return bond_len;
  }

  // One Simulation::TranslationalMove up to (not including) the energy evaluation.  Returns the move kind,
  // or -1 when the step attempts nothing.  Pivot rows are appended to rvec.
  int one_step(Counter& r, pg_move_desc& d, std::vector<double>& rvec) const {
    const int n_chain = (int)chains.size(), n_ion = (int)ions.size();
    if (n_chain + n_ion <= 0) return -1;                       // simulation.cc:242
    int which_chain = (int)std::floor(r.uniform() * n_chain);
    int which_ion = (int)std::floor(r.uniform() * n_ion);
    if (which_chain == n_chain) which_chain--;
    if (which_ion == n_ion) which_ion--;
    int move_type = 0;
    const double rand_num = r.uniform();
    double current = cfg.move_prob[0];
    while (current < rand_num && move_type < 4) {              // simulation.cc:271-276 (clamped to kNoMoveType - 1)
      move_type++;
      current += cfg.move_prob[move_type];
    }
    d = pg_move_desc();
    if (move_type == 0 && n_ion > 0) {
      // Molecule::BeadTranslate, molecule.cc:103-119
      d.mol = ions[which_ion];
      d.kind = PG_MOVE_BEAD;
      rand_sphere(r, d.v);
      double vec_len = std::sqrt(d.v[0] * d.v[0] + d.v[1] * d.v[1] + d.v[2] * d.v[2]);
      if (vec_len > 0) vec_len = 3 * cfg.move_size / vec_len;
      d.s = vec_len;
      return PG_MOVE_BEAD;
    }
    if (n_chain <= 0 || move_type == 0) return -1;             // simulation.cc:284, :315-316
    d.mol = chains[which_chain];
    const int len = cfg.mol_len[d.mol];
    switch (move_type) {
      case 1: {                                                // COMTranslate, molecule.cc:136-153
        for (int k = 0; k < 3; k++) d.v[k] = 0.5 * cfg.move_size * (double)r() / 4294967295.0;
        for (int k = 0; k < 3; k++)
          if (r() % 2 == 0) d.v[k] = -d.v[k];
        d.kind = PG_MOVE_COM;
        return PG_MOVE_COM;
      }
      case 2: {                                                // Pivot, molecule.cc:155-237
        int pivot = (int)std::floor(len * (double)r() / 4294967295.0);
        if (pivot == len) pivot--;
        d.i0 = pivot;
        d.s = cfg.move_size * (double)r() / 4294967295.0;
        d.rv_offset = (int)(rvec.size() / 4);
        for (int i = 0; i < len - 1; i++) {                    // forward arm, then backward arm: len - 1 draws of (sphere, bond)
          double v[3];
          rand_sphere(r, v);
          const double bl = varied_bond(r);
          rvec.push_back(v[0]); rvec.push_back(v[1]); rvec.push_back(v[2]); rvec.push_back(bl);
        }
        d.kind = PG_MOVE_PIVOT;
        return PG_MOVE_PIVOT;
      }
      case 3: {                                                // Crankshaft, molecule.cc:239-247
        const int first = (int)std::floor((len - 2) * (double)r() / 4294967295.0);
        const int last = (int)std::floor((len - first) * (double)r() / 4294967295.0) + first;
        d.i0 = first;
        d.rv_offset = last;
        d.s = cfg.move_size * ((double)r() / 4294967295.0 * 2 * M_PI - M_PI);
        d.v[0] = std::sin(d.s);                                // what AngleAxisd::toRotationMatrix takes of the angle:
        d.v[1] = std::cos(d.s);                                // evaluated here by the same libm as the reference's
        d.kind = PG_MOVE_CRANKSHAFT;
        return PG_MOVE_CRANKSHAFT;
      }
      default: {                                               // RandomReptation, molecule.cc:268-312
        d.s = varied_bond(r);
        d.i0 = (r() % 2 == 0) ? -1 : 1;
        rand_sphere(r, d.v);
        d.vlen = std::sqrt(d.v[0] * d.v[0] + d.v[1] * d.v[1] + d.v[2] * d.v[2]);
        d.kind = PG_MOVE_REPTATION;
        return PG_MOVE_REPTATION;
      }
    }
  }

  Config cfg;
  std::vector<int> chains, ions;
};

}  // namespace plum_mc

#endif
