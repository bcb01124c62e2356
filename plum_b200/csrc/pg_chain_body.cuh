// plum_b200 — the thread-count dependent part of pg_chain.cu (shared-memory block, pivot arms, the kernel itself);
// included twice with CH_THREADS / CH_MIN_CTAS / ChSmem / k_chain defined by the includer.  No include guard on purpose.

struct ChSmem {
  PgChainArgs A;
  PgMt mt;
  CgStep step;
  int g0, glen, nq, row, last_pair, accept, err, stop;
  double dE;
  double cur[3][CH_MAXLEN], trl[3][CH_MAXLEN], gq[CH_MAXLEN];
  double4 rv[CH_MAXLEN];
  int gtype[CH_MAXLEN], qidx[CH_MAXLEN];
  float4 fe[2 * CH_MAXLEN];
  double sq[2 * CH_MAXLEN];
  double2 tab[CH_TAB];
  int q_near[CH_WARPS][CH_NEAR];
  int2 q_hit[CH_WARPS][CH_QCAP];
  int ovf_hw;                  // overflow list high-water mark (kept in step by rank 0 through DSMEM)
  uint32_t pre_raw[32];        // the next 32 draws, tempered ...
  double pre_u[32];            // ... and converted to uniforms by warp 0's lanes in parallel (a uniform is an FP64 division)
  double uacc;                 // the draw behind the proposal as a uniform: the acceptance variate, if it gets drawn
  long long wst[CH_NPHASE][CH_WARPS];   // instrumentation: per-warp clock at the end of each phase
  double red[CH_WARPS][CH_NACC];
  double part[2][CH_GMAX][CH_NACC];
  double tot[CH_NACC];
  int scan[CH_WARPS];
  unsigned long long prof[5][CH_NPHASE];
};

// Bounding box (box fractions, minimum image around configuration 0) of the charged moved bead configurations, computed
// by one warp for itself: centre bc[3] and half widths bh[3] in every lane.
__device__ __forceinline__ void ch_bound_box(const ChSmem& sm, int n_e, int lane, float bc[3], float bh[3]) {
  const float4 c = sm.fe[0];
  float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f;
  for (int e = lane; e < n_e; e += 32) {
    const float4 f = sm.fe[e];
    float dx = f.x - c.x, dy = f.y - c.y, dz = f.z - c.z;
    dx -= mv_rintf(dx); dy -= mv_rintf(dy); dz -= mv_rintf(dz);
    lox = fminf(lox, dx); hix = fmaxf(hix, dx); loy = fminf(loy, dy); hiy = fmaxf(hiy, dy); loz = fminf(loz, dz); hiz = fmaxf(hiz, dz);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
    loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    loz = fminf(loz, __shfl_xor_sync(0xffffffffu, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(0xffffffffu, hiz, o));
  }
  // + slack for the FP32 rounding of the fractions; an extent of half a box or more bounds nothing
  const float hx = 0.5f * (hix - lox) + 2e-6f, hy = 0.5f * (hiy - loy) + 2e-6f, hz = 0.5f * (hiz - loz) + 2e-6f;
  bc[0] = c.x + 0.5f * (hix + lox); bc[1] = c.y + 0.5f * (hiy + loy); bc[2] = c.z + 0.5f * (hiz + loz);
  bh[0] = hx < 0.25f ? hx : 0.5f; bh[1] = hy < 0.25f ? hy : 0.5f; bh[2] = hz < 0.25f ? hz : 0.5f;
}

// One arm of Molecule::Pivot (molecule.cc:170-198 forward, :203-231 backward) in the reference's operation order, one
// warp: bead t of the arm (t = 1 .. La beads away from the pivot) is placed against the already placed bead t - 1 and the
// rest of the arm is dragged along.  The dependency bead -> bead is sequential by construction; every lane computes it
// redundantly, the dragged coordinates live in registers (lane l owns arm positions l + 1, l + 33, ...), so the only
// traffic per step is three shuffles that are off the critical path.
template <int DIR, bool TO_ROWS>
__device__ __forceinline__ void ch_pivot_arm(ChSmem& sm, int p, int glen, double msr, int lane) {
  const int La = DIR > 0 ? glen - 1 - p : p;
  const int rowbase = DIR > 0 ? 0 : glen - 1 - p;
  constexpr int NB = CH_MAXLEN / 32;
  double ox[NB], oy[NB], oz[NB];
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const int t = 32 * j + lane + 1;
    const int i = t <= La ? p + DIR * t : p;
    ox[j] = sm.cur[0][i]; oy[j] = sm.cur[1][i]; oz[j] = sm.cur[2][i];
  }
  double a[3] = {sm.cur[0][p], sm.cur[1][p], sm.cur[2][p]};
  double b[3] = {sm.cur[0][p + DIR], sm.cur[1][p + DIR], sm.cur[2][p + DIR]};
#pragma unroll
  for (int jb = 0; jb < NB; jb++) {
    if (32 * jb >= La) break;
    for (int r = 0; r < 32; r++) {
      const int t = 32 * jb + r + 1;
      if (t > La) break;
      const double4 row = sm.rv[rowbase + t - 1];
      // arm position t + 1 before this step's translation: lane r + 1 of this register block, or lane 0 of the next
      double n0, n1, n2;
      {
        constexpr int NXT = 0;   // (placeholder so that the block index below stays a compile-time constant)
        (void)NXT;
        const int jn = (jb + 1 < NB) ? jb + 1 : jb;
        const double sx_ = (r < 31) ? ox[jb] : ox[jn];
        const double sy_ = (r < 31) ? oy[jb] : oy[jn];
        const double sz_ = (r < 31) ? oz[jb] : oz[jn];
        const int src = (r + 1) & 31;
        n0 = __shfl_sync(0xffffffffu, sx_, src); n1 = __shfl_sync(0xffffffffu, sy_, src); n2 = __shfl_sync(0xffffffffu, sz_, src);
      }
      const double v[3] = {row.x, row.y, row.z};
      double m[3];
      pp_pivot_step(a, b, msr, v, row.w, m);
#pragma unroll
      for (int j = 0; j < NB; j++)
        if (j > jb || (j == jb && lane >= r)) { ox[j] = PP_ADD(ox[j], m[0]); oy[j] = PP_ADD(oy[j], m[1]); oz[j] = PP_ADD(oz[j], m[2]); }
      a[0] = PP_ADD(b[0], m[0]); a[1] = PP_ADD(b[1], m[1]); a[2] = PP_ADD(b[2], m[2]);
      b[0] = PP_ADD(n0, m[0]); b[1] = PP_ADD(n1, m[1]); b[2] = PP_ADD(n2, m[2]);
    }
  }
#pragma unroll
  for (int j = 0; j < NB; j++) {
    const int t = 32 * j + lane + 1;
    if (t <= La) {
      // TO_ROWS (pivot_mode 2): sm.trl holds the prefix-sum coordinates the energy phases are reading right now; the exact
      // ones go where this arm's rows were (every row has been consumed by now) and are copied over behind the decision
      if (TO_ROWS) sm.rv[rowbase + t - 1] = make_double4(ox[j], oy[j], oz[j], 0.0);
      else { const int i = p + DIR * t; sm.trl[0][i] = ox[j]; sm.trl[1][i] = oy[j]; sm.trl[2][i] = oz[j]; }
    }
  }
}

// The same arm as a prefix sum (pivot_mode 1): in exact arithmetic bead t sits at bead t - 1 + bond * unit(old_t -
// old_{t-1} + msr * v_t) — the placed neighbour cancels out of the direction — so the bond vectors are independent and
// the positions are their running sum.  Coordinates agree with the sequential form to a few ulp, not bit for bit.
template <int DIR>
__device__ __forceinline__ void ch_pivot_arm_prefix(ChSmem& sm, int p, int glen, double msr, int lane) {
  const int La = DIR > 0 ? glen - 1 - p : p;
  const int rowbase = DIR > 0 ? 0 : glen - 1 - p;
  const int C = (La + 31) / 32;   // consecutive arm positions per lane
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int q = 0; q < C; q++) {
    const int t = lane * C + q + 1;
    if (t <= La) {
      const int i = p + DIR * t;
      const double4 row = sm.rv[rowbase + t - 1];
      const double dx = (sm.cur[0][i] - sm.cur[0][i - DIR]) + msr * row.x, dy = (sm.cur[1][i] - sm.cur[1][i - DIR]) + msr * row.y,
                   dz = (sm.cur[2][i] - sm.cur[2][i - DIR]) + msr * row.z;
      const double nrm = row.w / sqrt(dx * dx + dy * dy + dz * dz);
      sx += nrm * dx; sy += nrm * dy; sz += nrm * dz;
      sm.trl[0][i] = sx; sm.trl[1][i] = sy; sm.trl[2][i] = sz;   // running sum inside the lane's stretch
    }
  }
  // exclusive scan of the lanes' totals
  double ex = sx, ey = sy, ez = sz;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double ux = __shfl_up_sync(0xffffffffu, ex, o), uy = __shfl_up_sync(0xffffffffu, ey, o), uz = __shfl_up_sync(0xffffffffu, ez, o);
    if (lane >= o) { ex += ux; ey += uy; ez += uz; }
  }
  const double bx = sm.cur[0][p] + (ex - sx), by = sm.cur[1][p] + (ey - sy), bz = sm.cur[2][p] + (ez - sz);
  for (int q = 0; q < C; q++) {
    const int t = lane * C + q + 1;
    if (t <= La) { const int i = p + DIR * t; sm.trl[0][i] += bx; sm.trl[1][i] += by; sm.trl[2][i] += bz; }
  }
}

__global__ void __launch_bounds__(CH_THREADS, CH_MIN_CTAS) k_chain(const PgChainArgs* __restrict__ all) {
  cgx::cluster_group cluster = cgx::this_cluster();
  const int G = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;

  extern __shared__ __align__(16) unsigned char ch_raw[];
  ChSmem& sm = *reinterpret_cast<ChSmem*>(ch_raw);
  {
    const int* src = reinterpret_cast<const int*>(all + blockIdx.x / G);
    int* dst = reinterpret_cast<int*>(&sm.A);
    for (int i = tid; i < (int)(sizeof(PgChainArgs) / 4); i += CH_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const PgChainArgs& A = sm.A;
  const PgMoveDev& P = A.D;
  for (int i = tid; i < CG_N; i += CH_THREADS) sm.mt.x[0][i] = A.mt_io[i];
  if (tid == 0) {
    sm.mt.cur = 0; sm.mt.p = (int)A.mt_io[CG_N]; sm.mt.need = 1;
    sm.err = 0; sm.stop = 0; sm.accept = 0;
  }
  __syncthreads();
  ch_mt_fix(sm.mt, tid);
  if (tid == 0 && sm.mt.p >= CG_N) cg_advance(sm.mt, 0);   // position 624: the first draw twists
  __syncthreads();
  ch_mt_fix(sm.mt, tid);

  // running totals (rank 0, thread 0): the reference's E_tot members
  double E_pair = 0, E_ewald = 0, E_bond = 0, E_ext = 0, E_real = 0, E_recip = 0;
  double n_inrange = 0.0;
  unsigned long long n_eval = 0ull, n_acc = 0ull;
  if (rank == 0 && tid == 0) {
    E_pair = A.state->E_pair; E_ewald = A.state->E_ewald; E_bond = A.state->E_bond; E_ext = A.state->E_ext;
    E_real = A.state->E_real; E_recip = A.state->E_recip;
  }
  if (tid == 0) sm.ovf_hw = __ldcg(A.ovf_n);   // overflow list high-water mark (rank 0 keeps every CTA's copy in step)

  const double Lx = P.box[0], Ly = P.box[1], Lz = P.box[2];
  const double iLx = P.inv_box[0], iLy = P.inv_box[1], iLz = P.inv_box[2];
  const float fLx = P.fbox[0], fLy = P.fbox[1], fLz = P.fbox[2];
  const double fdelta = 1e-6 * fmax(Lx, fmax(Ly, Lz));
  const float cutf = P.use_ewald ? mv_relax_f(P.rc2_relaxed, fdelta) : -1.0f;
  const int nk = P.use_ewald ? A.nk : 0;
  // Warp roles.  One CTA per chain (throughput): every warp takes its share of every phase in turn.  A cluster per chain
  // (latency): the four independent phases of the energy change — reciprocal space, real space, cell-grid LJ,
  // intra-molecular pairs — run side by side on disjoint warp groups of every CTA, so a step costs the longest of them
  // instead of their sum.
  const bool spec = (G > 1) && A.specialize;
  // (role indices are set per step below: in a pivot_mode 2 step warps 0 and 1 are busy with the exact arms)
  bool in_r, in_q, in_c, in_i;
  int rt, RT, gw, GW, ct, CT, it, IT;
  unsigned k_lo, k_hi;
  mv_share((unsigned)nk, (unsigned)G, (unsigned)rank, k_lo, k_hi);
  const int nq_tot = P.use_ewald ? A.nq_tot : 0;
  int nkt = 0;
  const int ne0 = A.kmax[0] + 1, ne1 = A.kmax[1] + 1, ne2 = A.kmax[2] + 1, ne = ne0 + ne1 + ne2;
  const int EC = CH_TAB / ne;
  if (EC < 1 && tid == 0) sm.err = CH_ERR_K;
  const int nbx = A.nc[0] >= 3 ? 3 : 1, nby = A.nc[1] >= 3 ? 3 : 1, nbz = A.nc[2] >= 3 ? 3 : 1;
  const int nnb = nbx * nby * nbz;
  const int ccap = A.cell_cap;
  __syncthreads();

  int step_i = 0;
  int par = 0;
  long long t_prev = 0;
  if (A.prof) {
    for (int i = tid; i < 5 * CH_NPHASE; i += CH_THREADS) (&sm.prof[0][0])[i] = 0ull;
    __syncthreads();
    t_prev = clock64();
  }
  const int skip = A.dbg_skip & 0xff, prof_rank = (A.dbg_skip >> 8) & 0xff, pivot_mode = A.pivot_mode;
  for (; step_i < A.max_steps; step_i++) {
    // ------------------------------------------------------------------ (1) the step's head: which move, which molecule
    if (warp == 0) {
      const uint32_t rw = cg_raw(sm.mt, lane);
      sm.pre_raw[lane] = rw;
      sm.pre_u[lane] = cg_uniform_of(rw);
      __syncwarp();
    }
    if (tid == 0) {
      CgStep d;
      const int* mf = A.mol_first;
      const int used = cg_step_header(sm.mt, A.cfg, A.chains, A.ions, [mf](int mol) { return mf[mol + 1] - mf[mol]; }, d,
                                      sm.pre_raw, sm.pre_u, 32);
      if (used > 560 && !sm.err) sm.err = CH_ERR_RNG;
      if (d.kind == CG_STOP_GC) sm.stop = 1;
      else cg_advance(sm.mt, used);
      if (d.kind >= 0) {
        sm.g0 = mf[d.mol]; sm.glen = mf[d.mol + 1] - mf[d.mol];
        if (sm.glen > CH_MAXLEN || sm.glen < 1 || (d.kind == CG_BEAD && sm.glen != 1)) sm.err = CH_ERR_LEN;
      }
      sm.step = d;
      sm.row = 0;
    }
    __syncthreads();
    if (sm.stop || sm.err) break;
    ch_mt_fix(sm.mt, tid);
    const int kind = sm.step.kind;
    if (kind == CG_NONE) {
      if (rank == 0 && tid == 0) { PgChainRec r; r.dE = 0.0; r.mol = -1; r.info = 0xff; A.log[step_i] = r; }
      __syncthreads();
      continue;
    }
    const int g0 = sm.g0, glen = sm.glen;
    int ovf_hw = sm.ovf_hw;
    CH_STAMP(0);
    // Warp roles of this step.  One CTA per chain: every warp takes its share of every phase in turn.  A cluster per
    // chain: the four independent parts of the energy change run side by side on disjoint warp groups of every CTA.
    // pivot_mode 2: warps 0 and 1 leave the pool (x0 = 2) — they build the exact arms while the others already evaluate
    // the energy change from the prefix-sum coordinates.
    int x0 = 0;
    if (kind == CG_PIVOT && pivot_mode == 2 && !(skip & 32)) {
      const int rt_threads = spec ? 64 : CH_THREADS - 64;
      if (((int)(k_hi - k_lo) + rt_threads - 1) / rt_threads <= CH_KPT) x0 = 2;   // else: this step runs as pivot_mode 0
    }
    {
      const int aW = CH_WARPS - x0;
      const int rW0 = x0, rWn = spec ? 4 - x0 : aW;
      const int qW0 = spec ? 4 : x0, qWn = spec ? 7 : aW;
      const int cW0 = spec ? 11 : x0, cWn = spec ? 4 : aW;
      const int iW0 = spec ? 15 : x0, iWn = spec ? 1 : aW;
      in_r = warp >= rW0 && warp < rW0 + rWn; in_q = warp >= qW0 && warp < qW0 + qWn;
      in_c = warp >= cW0 && warp < cW0 + cWn; in_i = warp >= iW0 && warp < iW0 + iWn;
      rt = (warp - rW0) * 32 + lane; RT = rWn * 32;                      // reciprocal space: thread of the CTA's k slice
      gw = rank * qWn + (warp - qW0); GW = G * qWn;                      // real space: warp over all charged partners
      ct = (rank * cWn + (warp - cW0)) * 32 + lane; CT = G * cWn * 32;   // cell units
      it = (rank * iWn + (warp - iW0)) * 32 + lane; IT = G * iWn * 32;   // intra-molecular pairs
      nkt = ((int)(k_hi - k_lo) + RT - 1) / RT;
    }
    const int wtid = tid - 32 * x0, WT = CH_THREADS - 32 * x0;   // thread index / count among the warps that evaluate

    // ------------------------------------------------------------------ (2) the molecule's current coordinates
    for (int i = tid; i < glen; i += CH_THREADS) {
      const double2 a = __ldcg(&A.xy[g0 + i]), c = __ldcg(&A.zq[g0 + i]);
      sm.cur[0][i] = a.x; sm.cur[1][i] = a.y; sm.cur[2][i] = c.x;
      sm.trl[0][i] = a.x; sm.trl[1][i] = a.y; sm.trl[2][i] = c.x;
      sm.gq[i] = c.y;
      sm.gtype[i] = A.type[g0 + i];
    }
    // ------------------------------------------------------------------ (3) pivot rows: len - 1 x (randSphere, bond length)
    if (kind == CG_PIVOT) {
      const int n_rows = sm.step.n_rows;
      for (;;) {
        const int row0 = sm.row;
        __syncthreads();   // everybody has read the row counter (and err) before thread 0 moves on
        if (row0 >= n_rows) break;
        if (A.cfg.vary_bond) {
          // the bond draw behind every accepted pair shifts the pairing of everything after it: one thread walks
          if (tid == 0) {
            int next = row0;
            const int used = cg_pivot_rows_serial(sm.mt, A.cfg, row0, n_rows, 400, reinterpret_cast<double*>(sm.rv), &next);
            if (used > 600) sm.err = CH_ERR_RNG;
            sm.row = next;
            cg_advance(sm.mt, used);
          }
          __syncthreads();
        } else {
          // rigid bonds: the stream is a sequence of candidate pairs, row r is the r-th accepted one: 256 pairs per pass
          bool acc = false;
          double v[3] = {0.0, 0.0, 0.0};
          if (tid < 256) acc = cg_sphere_pair(cg_uniform_of(cg_raw(sm.mt, 2 * tid)), cg_uniform_of(cg_raw(sm.mt, 2 * tid + 1)), v);
          const unsigned bal = __ballot_sync(0xffffffffu, acc);
          if (lane == 0) sm.scan[warp] = __popc(bal);
          __syncthreads();
          int before = 0, total = 0;
          for (int w = 0; w < 8; w++) { const int c = sm.scan[w]; if (w < warp) before += c; total += c; }
          const int need = n_rows - row0, mine = before + __popc(bal & lt);
          if (acc && mine < need) {
            sm.rv[row0 + mine] = make_double4(v[0], v[1], v[2], A.cfg.bond_len);
            if (mine == need - 1) sm.last_pair = tid;
          }
          __syncthreads();
          if (tid == 0) {
            if (total >= need) { cg_advance(sm.mt, 2 * (sm.last_pair + 1)); sm.row = n_rows; }
            else { cg_advance(sm.mt, 512); sm.row = row0 + total; }
          }
          __syncthreads();
        }
        ch_mt_fix(sm.mt, tid);
        if (sm.err) break;
      }
    }
    __syncthreads();
    if (sm.err) break;
    if (tid == 32) sm.uacc = cg_uniform_of(cg_raw(sm.mt, 0));   // off the critical path: read only behind the sums
    CH_STAMP(1);

    // ------------------------------------------------------------------ (4) trial coordinates (pg_propose_math.h)
    {
      const CgStep& d = sm.step;
      if (kind == CG_BEAD) {
        if (tid < 3) sm.trl[tid][0] = pp_bead_translate(sm.cur[tid][0], d.s, d.v[tid]);
      } else if (kind == CG_COM) {
        for (int i = tid; i < 3 * glen; i += CH_THREADS) {
          const int a = i / glen, g = i - a * glen;
          sm.trl[a][g] = pp_com_translate(sm.cur[a][g], d.v[a]);
        }
      } else if (kind == CG_REPT) {
        const int dir = d.i0, end = (dir > 0) ? glen - 1 : 0;
        for (int i = tid; i < 3 * glen; i += CH_THREADS) {
          const int a = i / glen, g = i - a * glen;
          sm.trl[a][g] = (g == end) ? pp_reptation_end(sm.cur[a][g], d.s, d.v[a], d.vlen) : sm.cur[a][g + dir];
        }
      } else if (kind == CG_CRANK) {
        // Molecule::Crankshaft, molecule.cc:239-265: beads strictly between the two axis beads turn about the axis.
        // The untouched beads keep trial == current: their "moved" pair terms and structure-factor terms are the
        // same arithmetic on the same numbers twice, i.e. exact zeros in every sum below.
        const int first = d.i0, last = min(d.i1, glen - 1);
        double* rot = reinterpret_cast<double*>(sm.rv);
        if (tid == 0) {
          const double pf[3] = {sm.cur[0][first], sm.cur[1][first], sm.cur[2][first]};
          const double pl[3] = {sm.cur[0][last], sm.cur[1][last], sm.cur[2][last]};
          double sn, cs;
          sincos(d.s, &sn, &cs);
          pp_crank_matrix(pf, pl, sn, cs, rot);
        }
        __syncthreads();
        for (int i = first + 1 + tid; i < last; i += CH_THREADS) {
          const double pf[3] = {sm.cur[0][first], sm.cur[1][first], sm.cur[2][first]};
          const double pos[3] = {sm.cur[0][i], sm.cur[1][i], sm.cur[2][i]};
          double out[3];
          pp_crank_apply(rot, pf, pos, out);
          sm.trl[0][i] = out[0]; sm.trl[1][i] = out[1]; sm.trl[2][i] = out[2];
        }
      } else if (!(skip & 32)) {   // CG_PIVOT: the two arms in two warps
        const int p = d.i0;
        if (pivot_mode == 0 || (pivot_mode == 2 && !x0)) {
          if (warp == 0 && p + 1 < glen) ch_pivot_arm<1, false>(sm, p, glen, d.s, lane);
          else if (warp == 1 && p > 0) ch_pivot_arm<-1, false>(sm, p, glen, d.s, lane);
        } else {
          if (warp == 0 && p + 1 < glen) ch_pivot_arm_prefix<1>(sm, p, glen, d.s, lane);
          else if (warp == 1 && p > 0) ch_pivot_arm_prefix<-1>(sm, p, glen, d.s, lane);
        }
      }
    }
    // charged moved beads (group-relative), in bead order
    if (warp == 2) {
      int cnt = 0;
      for (int base = 0; base < glen; base += 32) {
        const int i = base + lane;
        const bool c = P.use_ewald && i < glen && (kind != CG_BEAD || i == 0) && sm.gq[i] != 0.0;
        const unsigned b = __ballot_sync(0xffffffffu, c);
        if (c) sm.qidx[cnt + __popc(b & lt)] = i;
        cnt += __popc(b);
      }
      if (lane == 0) sm.nq = cnt;
    }
    __syncthreads();
    CH_STAMP(2);
    const int nq = sm.nq;
    if (x0 && warp < x0) {
      // pivot_mode 2: the exact arms (the reference's operation order) into the consumed rows, while the other warps go on
      const int p = sm.step.i0;
      if (warp == 0 && p + 1 < glen) ch_pivot_arm<1, true>(sm, p, glen, sm.step.s, lane);
      else if (warp == 1 && p > 0) ch_pivot_arm<-1, true>(sm, p, glen, sm.step.s, lane);
    } else {
      for (int e = wtid; e < 2 * nq; e += WT) {
        const int g = sm.qidx[e >> 1];
        const double(*c)[CH_MAXLEN] = (e & 1) ? sm.cur : sm.trl;
        sm.fe[e] = make_float4(mv_frac(c[0][g], iLx), mv_frac(c[1][g], iLy), mv_frac(c[2][g], iLz), 0.0f);
        sm.sq[e] = (e & 1) ? -sm.gq[g] : sm.gq[g];
      }
      if (x0) asm volatile("bar.sync 2, %0;" ::"r"(WT)); else __syncthreads();
    }

    double acc_pair = 0.0, acc_real = 0.0, acc_rec = 0.0, acc_ov = 0.0, w_sum = 0.0, b_sum = 0.0, w_out = 0.0, acc_cnt = 0.0;

    // ------------------------------------------------------------------ (5) reciprocal space: this CTA's k slice
    if (in_r && nq > 0 && nk > 0 && !(skip & 1)) {
      double dre[CH_KPT], dim[CH_KPT];
#pragma unroll
      for (int i = 0; i < CH_KPT; i++) { dre[i] = 0.0; dim[i] = 0.0; }
      for (int c0 = 0; c0 < 2 * nq; c0 += EC) {
        const int ec = min(EC, 2 * nq - c0);
        if (c0 > 0) { if (spec || x0) asm volatile("bar.sync 1, %0;" ::"r"(RT)); else __syncthreads(); }   // the previous chunk's tables are still being read
        for (int t = rt; t < 3 * ec; t += RT) {
          const int el = t / 3, ax = t - 3 * el, e = c0 + el;
          const int g = sm.qidx[e >> 1];
          double x = (e & 1) ? sm.cur[ax][g] : sm.trl[ax][g];
          x = pg_wrap_pos(x, P.ebox[ax], P.inv_ebox[ax], 1);
          double s1, c1;
          sincos(A.kunit[ax] * x, &s1, &c1);
          double2* row = sm.tab + el * ne + (ax == 0 ? 0 : (ax == 1 ? ne0 : ne0 + ne1));
          const int km = A.kmax[ax];
          row[0] = make_double2(1.0, 0.0);
          double cr = c1, sr = s1;
          if (km >= 1) row[1] = make_double2(c1, s1);
          for (int l = 2; l <= km; l++) {
            const double cn = cr * c1 - sr * s1, sn = sr * c1 + cr * s1;
            cr = cn; sr = sn;
            row[l] = make_double2(cr, sr);
          }
        }
        if (spec || x0) asm volatile("bar.sync 1, %0;" ::"r"(RT)); else __syncthreads();
#pragma unroll
        for (int i = 0; i < CH_KPT; i++) {
          const int k = (int)k_lo + rt + i * RT;
          if (i < nkt && k < (int)k_hi) {
            const int4 l = __ldg(&A.kl[k]);
            const int aly = abs(l.y), alz = abs(l.z);
            double ar = 0.0, ai = 0.0;
            for (int el = 0; el < ec; el++) {
              const double2* row = sm.tab + el * ne;
              const double2 a = row[l.x];
              double2 b = row[ne0 + aly], c = row[ne0 + ne1 + alz];
              if (l.y < 0) b.y = -b.y;
              if (l.z < 0) c.y = -c.y;
              const double abr = a.x * b.x - a.y * b.y, abi = a.x * b.y + a.y * b.x;
              const double w = sm.sq[c0 + el];
              ar += w * (abr * c.x - abi * c.y);
              ai += w * (abr * c.y + abi * c.x);
            }
            dre[i] += ar; dim[i] += ai;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < CH_KPT; i++) {
        const int k = (int)k_lo + rt + i * RT;
        if (i < nkt && k < (int)k_hi) {
          const double2 S = __ldcg(&A.S[k]);
          const double ek2 = __ldg(&A.ek2[k]);
          // never |S_new|^2 - |S_old|^2: 2 Re(conj(S) dS) + |dS|^2; x2 for the -k half
          acc_rec += 2.0 * ek2 * (2.0 * (S.x * dre[i] + S.y * dim[i]) + (dre[i] * dre[i] + dim[i] * dim[i]));
          __stcg(&A.dS[k], make_double2(dre[i], dim[i]));
        }
      }
      __syncwarp();
    }

    CH_STAMP(3);
    // ------------------------------------------------------------------ (6) real space: all charged partners, dealt to the warps
    // of the whole cluster with a stride (consecutive partners are beads of one chain: near or far together)
    if (in_q && nq > 0 && !(skip & 2)) {
      int qn = 0, nn = 0;
      const double rc2 = P.rc2_relaxed;
      float bc[3], bh[3];
      ch_bound_box(sm, 2 * nq, lane, bc, bh);
      // stage 1: partners whose distance from the moved beads' bounding box is within the cutoff (one test per partner);
      // stage 2 (CH_NEAR_PROCESS): FP32 filter of those against every configuration; stage 3 (CH_QUEUE_FLUSH): FP64
      for (int jb = 0; jb * GW < nq_tot; jb += 32) {
        const int j = (jb + lane) * GW + gw;
        bool near = false;
        if (j < nq_tot) {
          const float4 pf = __ldcg(&A.qfrac[j]);
          const int bead = __float_as_int(pf.w);
          float dx = pf.x - bc[0], dy = pf.y - bc[1], dz = pf.z - bc[2];
          dx -= mv_rintf(dx); dy -= mv_rintf(dy); dz -= mv_rintf(dz);
          dx = fmaxf(fabsf(dx) - bh[0], 0.0f) * fLx; dy = fmaxf(fabsf(dy) - bh[1], 0.0f) * fLy; dz = fmaxf(fabsf(dz) - bh[2], 0.0f) * fLz;
          // beads of the moved molecule itself are handled with the intra-molecular pairs (8)
          near = !(bead >= g0 && bead < g0 + glen) && fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= cutf;
        }
        const unsigned mn = __ballot_sync(0xffffffffu, near);
        if (near) sm.q_near[warp][nn + __popc(mn & lt)] = j;
        nn += __popc(mn);
        if (nn > CH_NEAR - 32) CH_NEAR_PROCESS();
      }
      if (nn > 0) CH_NEAR_PROCESS();
      if (qn > 0) CH_QUEUE_FLUSH();
    }

    CH_STAMP(4);
    // ------------------------------------------------------------------ (7) LJ / WCA through the cell grid
    if (in_c && P.pair_kind == 1 && !(skip & 4)) {
      const int n_mv = (kind == CG_BEAD) ? 1 : glen;
      // work units of one bead configuration: its 27 (or fewer) neighbour cells, then the overflow list in chunks
      const int per = nnb + (ovf_hw + CH_OVF_CHUNK - 1) / CH_OVF_CHUNK;
      const int U = n_mv * 2 * per;
      const double ljc2max = P.ljc2max;
      const PgDev* __restrict__ Pg = A.Pg;
      for (int u = ct; u < U; u += CT) {
        const int mc = u / per, cc = u - mc * per;
        const int m = mc >> 1, old = mc & 1;
        const double x = old ? sm.cur[0][m] : sm.trl[0][m], y = old ? sm.cur[1][m] : sm.trl[1][m],
                     z = old ? sm.cur[2][m] : sm.trl[2][m];
        const int tm = sm.gtype[m];
        // one candidate partner: exact FP64 minimum-image separation, the reference's r < rcut predicate
        auto lj_eval = [&](int j) {
          if (j < 0 || (j >= g0 && j < g0 + glen)) return;
          const double2 a = __ldcg(&A.xy[j]), c = __ldcg(&A.zq[j]);
          double dx = a.x - x, dy = a.y - y, dz = c.x - z;
          dx -= Lx * mv_rint(dx * iLx); dy -= Ly * mv_rint(dy * iLy); dz -= Lz * mv_rint(dz * iLz);
          const double r2 = dx * dx + dy * dy + dz * dz;
          if (r2 > ljc2max) return;
          const int tp = tm * PG_MAX_TYPES + A.type[j];
          if (r2 > Pg->lj_rcut2_relaxed[tp]) return;
          const double e = pg_pair_energy_r(*Pg, sqrt(r2), tp);
          acc_cnt += 1.0;
          if (old) acc_pair -= e;
          else { acc_pair += e; if (e >= PG_VLE) acc_ov += 1.0; }
        };
        if (cc < nnb) {
          int ix = ch_cell1(x, iLx, A.nc[0]), iy = ch_cell1(y, iLy, A.nc[1]), iz = ch_cell1(z, iLz, A.nc[2]);
          const int ox = (nbx == 3) ? (cc % 3) - 1 : 0;
          const int r1 = (nbx == 3) ? cc / 3 : cc;
          const int oy = (nby == 3) ? (r1 % 3) - 1 : 0;
          const int r2_ = (nby == 3) ? r1 / 3 : r1;
          const int oz = (nbz == 3) ? (r2_ % 3) - 1 : 0;
          ix += ox; iy += oy; iz += oz;
          if (ix < 0) ix += A.nc[0]; else if (ix >= A.nc[0]) ix -= A.nc[0];
          if (iy < 0) iy += A.nc[1]; else if (iy >= A.nc[1]) iy -= A.nc[1];
          if (iz < 0) iz += A.nc[2]; else if (iz >= A.nc[2]) iz -= A.nc[2];
          const int cidx = (ix * A.nc[1] + iy) * A.nc[2] + iz;
          const int4* cp = reinterpret_cast<const int4*>(A.cell_slots + (size_t)cidx * ccap);
          for (int q4 = 0; q4 < ccap; q4 += 4) {
            const int4 s0 = __ldcg(cp + (q4 >> 2));
            lj_eval(s0.x); lj_eval(s0.y); lj_eval(s0.z); lj_eval(s0.w);
          }
        } else {
          // the overflow list (beads whose cell was full when they arrived): normally a handful of entries
          const int o0 = (cc - nnb) * CH_OVF_CHUNK, o1 = min(o0 + CH_OVF_CHUNK, ovf_hw);
          for (int o = o0; o < o1; o++) lj_eval(__ldcg(&A.ovf[o]));
        }
      }
      __syncwarp();   // the lanes of a warp leave this loop at different times: reconverge before the next phase
    }

    // ------------------------------------------------------------------ (8) intra-molecular pairs (potential_pair.cc:157-178,
    // potential_ewald.cc:436-477): every pair of a chain move has a moved bead
    CH_STAMP(5);
    if (in_i && kind != CG_BEAD && glen > 1 && !(skip & 8)) {
      const int npairs = glen * (glen - 1) / 2;
      for (int p = it; p < npairs; p += IT) {
        int jj = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
        while (jj * (jj - 1) / 2 > p) jj--;
        while ((jj + 1) * jj / 2 <= p) jj++;
        const int g = p - jj * (jj - 1) / 2;
        double dxn = sm.trl[0][jj] - sm.trl[0][g], dyn = sm.trl[1][jj] - sm.trl[1][g], dzn = sm.trl[2][jj] - sm.trl[2][g];
        double dxo = sm.cur[0][jj] - sm.cur[0][g], dyo = sm.cur[1][jj] - sm.cur[1][g], dzo = sm.cur[2][jj] - sm.cur[2][g];
        dxn -= Lx * mv_rint(dxn * iLx); dyn -= Ly * mv_rint(dyn * iLy); dzn -= Lz * mv_rint(dzn * iLz);
        dxo -= Lx * mv_rint(dxo * iLx); dyo -= Ly * mv_rint(dyo * iLy); dzo -= Lz * mv_rint(dzo * iLz);
        const double r2n = dxn * dxn + dyn * dyn + dzn * dzn, r2o = dxo * dxo + dyo * dyo + dzo * dzo;
        const int tp = sm.gtype[g] * PG_MAX_TYPES + sm.gtype[jj];
        const double qq = P.use_ewald ? sm.gq[g] * sm.gq[jj] : 0.0;
        const double lim = fmax(P.pair_kind == 1 ? P.ljc2max : -1.0, (qq != 0.0) ? P.rc2_relaxed : -1.0);
        if (r2n <= lim) {
          acc_cnt += 1.0;
          const double2 en = mv_pair_inrange(A.Pg, r2n, qq, tp);
          if (en.x >= PG_VLE) acc_ov += 1.0;
          acc_pair += en.x; acc_real += en.y;
        }
        if (r2o <= lim) {
          acc_cnt += 1.0;
          const double2 eo = mv_pair_inrange(A.Pg, r2o, qq, tp);
          acc_pair -= eo.x; acc_real -= eo.y;
        }
      }
      __syncwarp();
    }

    CH_STAMP(6);
    // ------------------------------------------------------------------ (9) walls and bonds of the moved molecule (rank 0)
    if (rank == 0 && (P.ext_kind != 0 || P.bond_kind != 0)) {
      for (int g = tid; g < glen; g += CH_THREADS) {
        if (P.ext_kind != 0 && (kind != CG_BEAD || g == 0)) {
          const int t = sm.gtype[g];
          const double en = pg_wall_energy(*A.Pg, sm.trl[2][g], t);
          if (en >= PG_VLE) w_out = 1.0;
          w_sum += en - pg_wall_energy(*A.Pg, sm.cur[2][g], t);
        }
        if (P.bond_kind != 0 && g + 1 < glen)
          b_sum += pg_bond_energy(*A.Pg, sm.trl[0][g], sm.trl[1][g], sm.trl[2][g], sm.trl[0][g + 1], sm.trl[1][g + 1], sm.trl[2][g + 1]) -
                   pg_bond_energy(*A.Pg, sm.cur[0][g], sm.cur[1][g], sm.cur[2][g], sm.cur[0][g + 1], sm.cur[1][g + 1], sm.cur[2][g + 1]);
      }
    }

    CH_STAMP(7);
    // ------------------------------------------------------------------ (10) sums: warp -> CTA -> cluster, fixed order
    {
      double v[CH_NACC] = {acc_pair, acc_real, acc_rec, acc_ov, w_sum, b_sum, w_out, acc_cnt};
#pragma unroll
      for (int i = 0; i < CH_NACC; i++) v[i] = ch_warp_sum(v[i]);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < CH_NACC; i++) sm.red[warp][i] = v[i];
      }
      __syncthreads();
      if (tid < CH_NACC) {
        double s = 0.0;
        for (int w = 0; w < CH_WARPS; w++) s += sm.red[w][tid];
        if (G > 1) {
          for (int r = 0; r < G; r++) *cluster.map_shared_rank(&sm.part[par][rank][tid], r) = s;
        } else {
          sm.part[par][0][tid] = s;
        }
      }
      if (G > 1) cluster.sync(); else __syncthreads();
      if (tid < CH_NACC) {
        double s = 0.0;
        for (int r = 0; r < G; r++) s += sm.part[par][r][tid];
        sm.tot[tid] = s;
      }
      par ^= 1;
      __syncthreads();
    }

    CH_STAMP(8);
    // ------------------------------------------------------------------ (11) ForceField::EnergyDifference's orchestration and
    // the Metropolis test, in every CTA alike (same sums, same stream)
    double d_pair = 0, d_real = 0, d_recip = 0, d_ext = 0, d_bond = 0, d_ewald = 0, dE = 0;
    int stage = 0;
    if (tid == 0) {
      d_pair = sm.tot[0]; d_real = sm.tot[1];
      d_recip = P.use_ewald ? P.recip_pref * sm.tot[2] : 0.0;
      d_ext = sm.tot[4]; d_bond = sm.tot[5];
      if (sm.tot[6] > 0.0) d_ext = PG_VLE;   // potential_external.cc:107-110
      bool done = false;
      if (P.pair_kind != 0) { dE += d_pair; if (dE >= PG_VLE) { stage = 1; done = true; } }
      if (!done && P.ext_kind != 0) { dE += d_ext; if (dE >= PG_VLE) { stage = 2; done = true; } }
      if (!done) {
        if (P.use_ewald) { d_ewald = d_real + d_recip; dE += d_ewald; }
        if (P.bond_kind != 0) dE += d_bond;
      }
      int accept = 0;
      if (dE < PG_VLE) {                     // simulation.cc:327-332: the variate is drawn only here
        const double uacc = sm.uacc;
        cg_advance(sm.mt, 1);
        accept = uacc < exp(-P.beta * dE);
      }
      sm.accept = accept;
      sm.dE = dE;
      if (rank == 0) {
        PgChainRec r; r.dE = dE; r.mol = sm.step.mol; r.info = (kind & 0xff) | (accept << 8) | (stage << 16);
        A.log[step_i] = r;
        n_inrange += sm.tot[7]; n_eval++; n_acc += (unsigned long long)accept;
        if (accept) {
          E_pair += d_pair; E_ewald += d_ewald; E_bond += d_bond; E_ext += d_ext; E_real += d_real; E_recip += d_recip;
        }
      }
    }
    __syncthreads();
    ch_mt_fix(sm.mt, tid);
    if (x0 && (sm.accept || A.trial_log)) {
      // pivot_mode 2: what becomes the state (and what the tests read back) are the exact arms, parked in the rows
      const int p = sm.step.i0;
      for (int i = tid; i < glen; i += CH_THREADS)
        if (i != p) {
          const double4 r = sm.rv[i > p ? i - p - 1 : glen - 2 - i];
          sm.trl[0][i] = r.x; sm.trl[1][i] = r.y; sm.trl[2][i] = r.z;
        }
      __syncthreads();
    }
    if (A.trial_log && rank == 0)
      for (int i = tid; i < 3 * glen; i += CH_THREADS) {
        const int g = i / 3, a = i - 3 * g;
        A.trial_log[((size_t)step_i * A.trial_stride + g) * 3 + a] = sm.trl[a][g];
      }
    CH_STAMP(9);

    // ------------------------------------------------------------------ (12) FinalizeEnergies: an accepted move becomes the state
    if (sm.accept && !(skip & 16)) {
      if (in_r && nq > 0 && nk > 0) {
#pragma unroll
        for (int i = 0; i < CH_KPT; i++) {
          const int k = (int)k_lo + rt + i * RT;
          if (i < nkt && k < (int)k_hi) {
            double2 S = __ldcg(&A.S[k]);
            const double2 d = __ldcg(&A.dS[k]);
            S.x += d.x; S.y += d.y;
            __stcg(&A.S[k], S);
          }
        }
      }
      if (rank == 0) {
        for (int i = tid; i < glen; i += CH_THREADS) {
          if (kind == CG_BEAD && i != 0) continue;
          const int jg = g0 + i;
          const double x = sm.trl[0][i], y = sm.trl[1][i], z = sm.trl[2][i], q = sm.gq[i];
          __stcg(&A.xy[jg], make_double2(x, y));
          __stcg(&A.zq[jg], make_double2(z, q));
          if (q != 0.0 && P.use_ewald) {
            const int s = A.qslot[jg];
            __stcg(&A.qpos[2 * s], make_double2(x, y));
            __stcg(&A.qpos[2 * s + 1], make_double2(z, q));
            __stcg(&A.qfrac[s], make_float4(mv_frac(x, iLx), mv_frac(y, iLy), mv_frac(z, iLz), __int_as_float(jg)));
          }
        }
        // cell grid: one warp, 32 beads at a time, every lane files its own bead: lanes that enter the same cell
        // (__match_any_sync) take its free slots in bead order, whoever finds none goes to the overflow list
        if (warp == 0 && P.pair_kind == 1) {
          const int n_mv = (kind == CG_BEAD) ? 1 : glen;
          int hw = ovf_hw;
          bool full = false;
          for (int base = 0; base < n_mv; base += 32) {
            const int i = base + lane;
            const bool active = i < n_mv;
            const int bead = g0 + i;
            int newc = -1, oldc = -1;
            if (active) {
              newc = (ch_cell1(sm.trl[0][i], iLx, A.nc[0]) * A.nc[1] + ch_cell1(sm.trl[1][i], iLy, A.nc[1])) * A.nc[2] +
                     ch_cell1(sm.trl[2][i], iLz, A.nc[2]);
              oldc = __ldcg(&A.bead_cell[bead]);
            }
            const bool changed = active && newc != oldc;
            if (changed) {
              const int os = __ldcg(&A.bead_slot[bead]);
              if (os < ccap) __stcg(&A.cell_slots[(size_t)oldc * ccap + os], -1);
              else __stcg(&A.ovf[os - ccap], -1);
            }
            __syncwarp();   // the freed slots are visible to the lanes that look for one
            const unsigned grp = __match_any_sync(0xffffffffu, changed ? newc : -1 - lane);
            unsigned empty = 0u;
            if (changed) {
              const int4* cp = reinterpret_cast<const int4*>(A.cell_slots + (size_t)newc * ccap);
              for (int q4 = 0; q4 < ccap; q4 += 4) {
                const int4 sv = __ldcg(cp + (q4 >> 2));
                empty |= ((sv.x < 0 ? 1u : 0u) | (sv.y < 0 ? 2u : 0u) | (sv.z < 0 ? 4u : 0u) | (sv.w < 0 ? 8u : 0u)) << q4;
              }
            }
            const int r = __popc(grp & lt), n_empty = __popc(empty);
            const bool spill = changed && r >= n_empty;
            const unsigned ms = __ballot_sync(0xffffffffu, spill);
            if (changed) {
              if (!spill) {
                const int slot = __fns(empty, 0, r + 1);
                __stcg(&A.cell_slots[(size_t)newc * ccap + slot], bead);
                __stcg(&A.bead_slot[bead], slot);
              } else {
                const int idx = hw + __popc(ms & lt);
                if (idx < CH_OVF_CAP) { __stcg(&A.ovf[idx], bead); __stcg(&A.bead_slot[bead], ccap + idx); }
                else full = true;
              }
              __stcg(&A.bead_cell[bead], newc);
            }
            hw = min(hw + __popc(ms), CH_OVF_CAP);
            __syncwarp();
          }
          full = __any_sync(0xffffffffu, full);
          if (lane == 0) {
            if (hw != ovf_hw) __stcg(A.ovf_n, hw);
            for (int r = 0; r < G; r++) {
              *cluster.map_shared_rank(&sm.ovf_hw, r) = hw;
              if (full) *cluster.map_shared_rank(&sm.err, r) = (int)CH_ERR_OVERFLOW;
            }
          }
        }
      }
      // (the cluster barrier's release / acquire pair orders these global writes for the other CTAs)
      if (G > 1) cluster.sync(); else __syncthreads();
    }
    __syncthreads();
    CH_STAMP(10);
    if (A.prof) {
      __syncthreads();
      if (tid == 0 && rank == prof_rank) {
        long long prev = t_prev;
        for (int ph = 0; ph <= 10; ph++) {
          long long mx = prev;
          for (int w = 0; w < CH_WARPS; w++) mx = sm.wst[ph][w] > mx ? sm.wst[ph][w] : mx;
          sm.prof[kind][ph] += (unsigned long long)(mx - prev);
          prev = mx;
        }
        sm.prof[kind][CH_NPHASE - 1] += 1ull;   // moves of this kind
        t_prev = clock64();
      }
      __syncthreads();
    }
  }

  // ------------------------------------------------------------------ epilogue
  __syncthreads();
  if (A.prof && rank == prof_rank)
    for (int i = tid; i < 5 * CH_NPHASE; i += CH_THREADS) A.prof[i] = (&sm.prof[0][0])[i];
  if (rank == 0) {
    for (int i = tid; i < CG_N; i += CH_THREADS) A.mt_io[i] = sm.mt.x[sm.mt.cur][i];
    if (tid == 0) {
      A.mt_io[CG_N] = (uint32_t)sm.mt.p;
      PgState* st = A.state;
      st->E_pair = E_pair; st->E_ewald = E_ewald; st->E_bond = E_bond; st->E_ext = E_ext; st->E_real = E_real; st->E_recip = E_recip;
      A.out[0] = step_i; A.out[1] = sm.stop; A.out[3] = sm.ovf_hw;
      if (A.counters) { A.counters[0] += (unsigned long long)n_inrange; A.counters[1] += n_eval; A.counters[2] += n_acc; }
      if (sm.err) A.out[2] = sm.err;
    }
  }
  if (G > 1) cluster.sync();
}
