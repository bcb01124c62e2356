// CPU pin of plum_b200/csrc/pg_chain_gen.h (the random stream + draw order the device-resident chain kernel uses)
// against plum_b200/host/mc_propose.h, which is itself pinned to the reference's own trial coordinates
// (tests/test_proposals_cpu.py).  Walks both generators side by side from one std::mt19937 seed and compares every
// descriptor field, every pivot row and the stream position bit for bit.  Also checks the two-buffer mt19937
// against std::mt19937 directly.  Exit code 0 = identical.
//   g++ -O2 -std=c++14 -ffp-contract=off -I include -I plum_b200/host -I plum_b200/csrc tests/native/chain_gen_check.cc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sstream>
#include <vector>

#include "mc_propose.h"
#include "mt_state.h"
#include "pg_chain_gen.h"

static void load_mt(const std::mt19937& g, PgMt& m) {
  std::stringstream ss;
  ss << g;
  for (int i = 0; i < CG_N; i++) { unsigned long v; ss >> v; m.x[0][i] = (uint32_t)v; }
  unsigned long p; ss >> p;
  m.cur = 0; m.p = (int)p; m.need = 0;
  cg_twist_serial(m.x[0], m.x[1]);
  if (m.p >= CG_N) { m.p -= CG_N; m.cur = 1; cg_twist_serial(m.x[1], m.x[0]); }
}
static void fix(PgMt& m) { if (m.need) { cg_twist_serial(m.x[m.cur], m.x[m.cur ^ 1]); m.need = 0; } }

int main(int argc, char** argv) {
  const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
  const int n_steps = argc > 2 ? atoi(argv[2]) : 20000;
  const int vary = argc > 3 ? atoi(argv[3]) : 0;
  const int gc_freq = argc > 4 ? atoi(argv[4]) : 0;
  // host <-> device hand-over of the generator (plum_b200/host/mt_state.h): the byte-copy path, when its self-test
  // accepts the library's layout, must agree with the text form and continue the same stream
  {
    std::mt19937 g(seed + 77u);
    for (int i = 0; i < 1000; i++) (void)g();
    uint32_t a[624], b[624]; int pa = 0, pb = 0;
    plum_mt::slow_export(g, a, &pa);
    plum_mt::export_state(g, b, &pb);
    if (pa != pb || memcmp(a, b, sizeof(a))) { printf("mt_state: export differs from the text form\n"); return 1; }
    std::mt19937 h1, h2;
    plum_mt::slow_import(h1, a, pa);
    plum_mt::import_state(h2, b, pb);
    for (int i = 0; i < 2000; i++) { const auto x = g(); if (h1() != x || h2() != x) { printf("mt_state: import does not continue the stream\n"); return 1; } }
    printf("mt_state fast path: %s\n", plum_mt::fast_ok() ? "on" : "off (text form)");
  }
  // raw stream first
  {
    std::mt19937 g(seed);
    PgMt m; load_mt(g, m);
    for (int i = 0; i < 5000; i++) {
      const uint32_t a = (uint32_t)g(), b = cg_raw(m, 0);
      cg_advance(m, 1); fix(m);
      if (a != b) { printf("raw stream differs at draw %d\n", i); return 1; }
    }
    // look-ahead across the generation boundary
    std::mt19937 g2 = g;
    for (int i = 0; i < 624; i++) if ((uint32_t)g2() != cg_raw(m, i)) { printf("look-ahead differs at %d\n", i); return 1; }
  }
  // a system: 3 phantoms, chains of mixed length and ions, interleaved
  plum_mc::Config cfg;
  cfg.phantom = 3;
  std::vector<int> len = {1, 1, 1, 100, 1, 16, 1, 1, 7, 2, 1, 100, 1};
  cfg.mol_len = len;
  cfg.move_size = 2.0;
  const double prob[5] = {0.4, 0.1, 0.25, 0.15, 0.1};   // crankshaft on (zero in the reference's own examples)
  for (int i = 0; i < 5; i++) cfg.move_prob[i] = prob[i];
  cfg.bond_len = 2.5; cfg.vary_bond = vary != 0; cfg.gc_freq = gc_freq;
  plum_mc::Proposer prop; prop.configure(cfg);
  std::vector<int> chains, ions;
  for (int i = cfg.phantom; i < (int)len.size(); i++) (len[i] > 1 ? chains : ions).push_back(i);
  CgConfig c;
  c.n_chain = (int)chains.size(); c.n_ion = (int)ions.size(); c.gc_freq = gc_freq; c.vary_bond = vary;
  c.move_size = cfg.move_size; c.bond_len = cfg.bond_len;
  for (int i = 0; i < 5; i++) c.prob[i] = prob[i];

  std::mt19937 g(seed);
  std::mt19937 coin(seed * 7919u + 13u);
  PgMt m; load_mt(g, m);
  plum_mc::Batch b;
  std::vector<double> rows(4 * 256);
  long n_moves = 0, n_none = 0, n_gc = 0, n_skip = 0;
  for (int s = 0; s < n_steps; s++) {
    const int got = prop.generate(g, 1, b);
    CgStep d;
    int used = cg_step_header(m, c, chains.data(), ions.data(), [&](int mol) { return len[mol]; }, d);
    if (got == 0) {
      if (b.stop != plum_mc::STOP_GC || d.kind != CG_STOP_GC) { printf("step %d: stop mismatch\n", s); return 1; }
      // the driver would now run its GC step: consume the draw both sides and go on
      n_gc++;
      (void)g(); cg_advance(m, 1); fix(m);
      continue;
    }
    if (d.kind == CG_STOP_GC) { printf("step %d: spurious GC stop\n", s); return 1; }
    cg_advance(m, used); fix(m);
    if (b.kind[0] != d.kind) { printf("step %d: kind %d vs %d\n", s, b.kind[0], d.kind); return 1; }
    if (d.kind < 0) { n_none++; goto check_pos; }
    {
      const pg_move_desc& r = b.moves[0];
      if (d.kind == CG_PIVOT) {
        int row = 0;
        while (row < d.n_rows) {   // in budgeted passes, like the kernel
          int next = row;
          const int u2 = cg_pivot_rows_serial(m, c, row, d.n_rows, 300, rows.data(), &next);
          cg_advance(m, u2); fix(m);
          row = next;
        }
        if ((int)(b.rvec.size() / 4) != d.n_rows || memcmp(b.rvec.data(), rows.data(), sizeof(double) * 4 * d.n_rows)) {
          printf("step %d: pivot rows differ\n", s); return 1;
        }
      }
      if (d.kind == CG_CRANK) {   // the descriptor carries last bead, sin and cos where the chain kernel has i1 and its own sincos
        if (r.rv_offset != d.i1 || r.mol != d.mol || r.i0 != d.i0 || memcmp(&r.s, &d.s, 8)) {
          printf("step %d: crankshaft differs (first %d/%d last %d/%d angle %a/%a)\n", s, r.i0, d.i0, r.rv_offset, d.i1, r.s, d.s);
          return 1;
        }
      } else if (r.mol != d.mol || r.kind != d.kind || r.i0 != d.i0 || memcmp(&r.s, &d.s, 8) || memcmp(r.v, d.v, 24) ||
          memcmp(&r.vlen, &d.vlen, 8)) {
        printf("step %d: descriptor differs (kind %d mol %d/%d i0 %d/%d s %a/%a)\n", s, d.kind, r.mol, d.mol, r.i0, d.i0, r.s, d.s);
        return 1;
      }
      n_moves++;
      // the acceptance draw: skipped when dE >= 1e8 (simulation.cc:327-332) — decided by a coin here
      if (coin() % 16 == 0) {
        b.rewind_after_overlap(g, 1);
        n_skip++;
      } else {
        const double u = cg_uniform_of(cg_raw(m, 0));
        cg_advance(m, 1); fix(m);
        if (memcmp(&u, &r.u, 8)) { printf("step %d: acceptance variate differs\n", s); return 1; }
      }
    }
  check_pos:
    {
      std::mt19937 peek = g;
      if ((uint32_t)peek() != cg_raw(m, 0)) { printf("step %d: stream position differs\n", s); return 1; }
    }
  }
  printf("ok steps=%d moves=%ld none=%ld gc=%ld no_accept_draw=%ld\n", n_steps, n_moves, n_none, n_gc, n_skip);
  return 0;
}
