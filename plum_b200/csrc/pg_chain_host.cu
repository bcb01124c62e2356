// plum_b200 — host side of the device-resident Markov chain (k_chain, pg_chain.cu): the pg_chain_* entry points of
// include/plum_b200.h.  Included at the end of pg_engine.cu.
//
// The resident spatial structures (cell grid, compact charged list) are BUILT here on the host, deterministically (beads
// in index order), whenever another path has changed the coordinates or the molecule table behind the chain's back
// (`ch.valid == false`); inside a chain they are kept up to date by the kernel itself.

namespace {

const double kChainCellMargin = 1.001;   // cell edge >= 1.001 x the largest (relaxed) LJ cutoff

void chain_free(pg_engine* h) {
  PgChainHost& c = h->ch;
  cudaFree(c.d_mol_first); cudaFree(c.d_chains); cudaFree(c.d_ions);
  cudaFree(c.d_cell_slots); cudaFree(c.d_ovf); cudaFree(c.d_ovf_n); cudaFree(c.d_bead_cell); cudaFree(c.d_bead_slot);
  cudaFree(c.d_qslot); cudaFree(c.d_qpos); cudaFree(c.d_qfrac);
  if (c.h_pin) cudaFreeHost(c.h_pin);
  cudaFree(c.d_pack);
  cudaFree(c.d_mt); cudaFree(c.d_log); cudaFree(c.d_trial_log); cudaFree(c.d_out); cudaFree(c.d_args); cudaFree(c.d_prof); cudaFree(c.d_counters);
  c = PgChainHost();
}

int chain_kernel_setup(pg_engine* h, int cluster) {
  static bool smem_done = false, np_done = false;
  if (!smem_done) {
    PG_CUDA(h, cudaFuncSetAttribute(k_chain512, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem512)));
    PG_CUDA(h, cudaFuncSetAttribute(k_chain448, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChSmem448)));
    PG_CUDA(h, cudaFuncSetAttribute(k_chain448, cudaFuncAttributePreferredSharedMemoryCarveout, 100));   // two CTAs per SM
    smem_done = true;
  }
  if (cluster > 8 && !np_done) {
    PG_CUDA(h, cudaFuncSetAttribute(k_chain512, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    np_done = true;
  }
  return PG_OK;
}

inline int chain_cell_of(const PgChainHost& c, const PgDev& P, double x, double y, double z) {
  return (ch_cell1(x, P.inv_box[0], c.nc[0]) * c.nc[1] + ch_cell1(y, P.inv_box[1], c.nc[1])) * c.nc[2] +
         ch_cell1(z, P.inv_box[2], c.nc[2]);
}
inline float chain_frac(double x, double invL) {
  const double s = x * invL;
  return (float)(s - rint(s));
}

// (Re)build everything the kernel keeps resident from the engine's current coordinates and molecule table.
int chain_build(pg_engine* h) {
  PgChainHost& c = h->ch;
  const PgDev& P = h->P;
  int rc = flush_commit(h);
  if (rc) return rc;
  const int n = h->n, n_mol = h->n_mol;
  std::vector<double2> hxy(std::max(n, 1)), hzq(std::max(n, 1));
  if (n > 0) {
    PG_CUDA(h, cudaMemcpyAsync(hxy.data(), h->xy, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(hzq.data(), h->zq, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
  }
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  // molecule table, movable chains and ions (simulation.cc:255-268: molecules behind the phantoms, chains have > 1 bead)
  std::vector<int> chains, ions;
  int max_len = 1;
  for (int m = c.phantom; m < n_mol; m++) {
    const int len = h->mol_first[m + 1] - h->mol_first[m];
    (len > 1 ? chains : ions).push_back(m);
    max_len = std::max(max_len, len);
  }
  if (max_len > CH_MAXLEN) { h->err = "chain: a movable molecule is longer than CH_MAXLEN"; return PG_ERR_CAPACITY; }
  c.max_len = max_len;
  c.cfg.n_chain = (int)chains.size();
  c.cfg.n_ion = (int)ions.size();
  const size_t mol_need = (size_t)n_mol + 2;
  if (mol_need > c.mol_cap) {
    cudaFree(c.d_mol_first); cudaFree(c.d_chains); cudaFree(c.d_ions);
    c.d_mol_first = c.d_chains = c.d_ions = nullptr; c.mol_cap = 0;
    const size_t cap = mol_need * 2;
    PG_CUDA(h, cudaMalloc((void**)&c.d_mol_first, sizeof(int) * cap));
    PG_CUDA(h, cudaMalloc((void**)&c.d_chains, sizeof(int) * cap));
    PG_CUDA(h, cudaMalloc((void**)&c.d_ions, sizeof(int) * cap));
    c.mol_cap = cap;
  }
  PG_CUDA(h, cudaMemcpyAsync(c.d_mol_first, h->mol_first.data(), sizeof(int) * (size_t)(n_mol + 1), cudaMemcpyHostToDevice, h->stream));
  if (!chains.empty()) PG_CUDA(h, cudaMemcpyAsync(c.d_chains, chains.data(), sizeof(int) * chains.size(), cudaMemcpyHostToDevice, h->stream));
  if (!ions.empty()) PG_CUDA(h, cudaMemcpyAsync(c.d_ions, ions.data(), sizeof(int) * ions.size(), cudaMemcpyHostToDevice, h->stream));
  // cell grid: the finest split with cells no smaller than the largest LJ cutoff (<= 64 cells per axis, <= 2^18 cells);
  // slots per cell: the smallest of 4 / 8 / 16 / 32 that leaves at most 0.1 % of the beads to the overflow list
  for (int a = 0; a < 3; a++) c.nc[a] = 1;
  if (P.pair_kind == PG_PAIR_TRUNCATED_LJ && P.lj_rcut2_relaxed_max > 0) {
    const double edge = sqrt(P.lj_rcut2_relaxed_max) * kChainCellMargin;
    int lim = CH_NC_MAX;
    for (;;) {
      for (int a = 0; a < 3; a++) {
        int k = std::min((int)floor(P.box[a] / edge), lim);
        c.nc[a] = (k >= 3) ? k : 1;   // fewer than 3 cells: the axis is not split (27 neighbours would alias)
      }
      if ((long long)c.nc[0] * c.nc[1] * c.nc[2] <= CH_CELLS_MAX || lim <= 3) break;
      lim--;
    }
  }
  const size_t n_cells = (size_t)c.nc[0] * c.nc[1] * c.nc[2];
  std::vector<int> cell_of(std::max(n, 1), 0), count(n_cells, 0);
  if (P.pair_kind == PG_PAIR_TRUNCATED_LJ)
    for (int i = 0; i < n; i++) {
      cell_of[i] = chain_cell_of(c, P, hxy[i].x, hxy[i].y, hzq[i].x);
      count[cell_of[i]]++;
    }
  int cap = 4;
  for (; cap < CH_CELL_CAP_MAX; cap *= 2) {
    long long over = 0;
    for (size_t k = 0; k < n_cells; k++) over += std::max(0, count[k] - cap);
    if (over * 1000 <= (long long)n) break;
  }
  c.cell_cap = cap;
  if (n_cells * (size_t)cap > c.cell_cap_words || !c.d_cell_slots) {
    cudaFree(c.d_cell_slots); c.d_cell_slots = nullptr; c.cell_cap_words = 0;
    PG_CUDA(h, cudaMalloc((void**)&c.d_cell_slots, sizeof(int) * n_cells * (size_t)cap));
    c.cell_cap_words = n_cells * (size_t)cap;
  }
  if (!c.d_ovf) {
    PG_CUDA(h, cudaMalloc((void**)&c.d_ovf, sizeof(int) * CH_OVF_CAP));
    PG_CUDA(h, cudaMalloc((void**)&c.d_ovf_n, sizeof(int)));
  }
  if (!c.d_counters) {
    PG_CUDA(h, cudaMalloc((void**)&c.d_counters, sizeof(unsigned long long) * 4));
    PG_CUDA(h, cudaMemset(c.d_counters, 0, sizeof(unsigned long long) * 4));
  }
  if (!c.d_mt) {
    PG_CUDA(h, cudaMalloc((void**)&c.d_mt, sizeof(uint32_t) * (CG_N + 8)));
    PG_CUDA(h, cudaMemset(c.d_mt, 0, sizeof(uint32_t) * (CG_N + 8)));
    PG_CUDA(h, cudaMalloc((void**)&c.d_out, sizeof(int) * 4));
    PG_CUDA(h, cudaMemset(c.d_out, 0, sizeof(int) * 4));
  }
  const size_t bead_need = (size_t)std::max(h->cap, 1);
  if (bead_need > c.bead_cap) {
    cudaFree(c.d_bead_cell); cudaFree(c.d_bead_slot); cudaFree(c.d_qslot);
    c.d_bead_cell = c.d_bead_slot = c.d_qslot = nullptr; c.bead_cap = 0;
    PG_CUDA(h, cudaMalloc((void**)&c.d_bead_cell, sizeof(int) * bead_need));
    PG_CUDA(h, cudaMalloc((void**)&c.d_bead_slot, sizeof(int) * bead_need));
    PG_CUDA(h, cudaMalloc((void**)&c.d_qslot, sizeof(int) * bead_need));
    c.bead_cap = bead_need;
  }
  std::vector<int> slots((size_t)cap * n_cells, -1), ovf(CH_OVF_CAP, -1), bead_cell(std::max(n, 1), 0),
      bead_slot(std::max(n, 1), 0), qslot(std::max(n, 1), -1);
  std::fill(count.begin(), count.end(), 0);
  int ovf_n = 0;
  if (P.pair_kind == PG_PAIR_TRUNCATED_LJ) {
    for (int i = 0; i < n; i++) {
      const int cell = cell_of[i];
      bead_cell[i] = cell;
      if (count[cell] < cap) {
        bead_slot[i] = count[cell];
        slots[(size_t)cell * cap + count[cell]++] = i;
      } else {
        if (ovf_n >= CH_OVF_CAP) { h->err = "chain: cell overflow list is full (system too dense for the cell grid)"; return PG_ERR_CAPACITY; }
        bead_slot[i] = cap + ovf_n;
        ovf[ovf_n++] = i;
      }
    }
  }
  // compact charged list
  std::vector<double2> qpos;
  std::vector<float4> qfrac;
  if (P.use_ewald) {
    for (int i = 0; i < n; i++)
      if (hzq[i].y != 0.0) {
        qslot[i] = (int)qfrac.size();
        qpos.push_back(hxy[i]);
        qpos.push_back(hzq[i]);
        float4 f;
        f.x = chain_frac(hxy[i].x, P.inv_box[0]); f.y = chain_frac(hxy[i].y, P.inv_box[1]); f.z = chain_frac(hzq[i].x, P.inv_box[2]);
        memcpy(&f.w, &i, sizeof(int));
        qfrac.push_back(f);
      }
  }
  c.nq_tot = (int)qfrac.size();
  const size_t q_need = (size_t)std::max(h->cap, 1);   // every bead could be charged
  if (q_need > c.q_cap) {
    cudaFree(c.d_qpos); cudaFree(c.d_qfrac); c.d_qpos = nullptr; c.d_qfrac = nullptr; c.q_cap = 0;
    PG_CUDA(h, cudaMalloc((void**)&c.d_qpos, sizeof(double2) * 2 * q_need));
    PG_CUDA(h, cudaMalloc((void**)&c.d_qfrac, sizeof(float4) * q_need));
    c.q_cap = q_need;
  }
  PG_CUDA(h, cudaMemcpyAsync(c.d_cell_slots, slots.data(), sizeof(int) * slots.size(), cudaMemcpyHostToDevice, h->stream));
  PG_CUDA(h, cudaMemcpyAsync(c.d_ovf, ovf.data(), sizeof(int) * CH_OVF_CAP, cudaMemcpyHostToDevice, h->stream));
  PG_CUDA(h, cudaMemcpyAsync(c.d_ovf_n, &ovf_n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (n > 0) {
    PG_CUDA(h, cudaMemcpyAsync(c.d_bead_cell, bead_cell.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(c.d_bead_slot, bead_slot.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(c.d_qslot, qslot.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
  }
  if (c.nq_tot > 0) {
    PG_CUDA(h, cudaMemcpyAsync(c.d_qpos, qpos.data(), sizeof(double2) * qpos.size(), cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(h, cudaMemcpyAsync(c.d_qfrac, qfrac.data(), sizeof(float4) * qfrac.size(), cudaMemcpyHostToDevice, h->stream));
  }
  PG_CUDA(h, cudaStreamSynchronize(h->stream));   // the host vectors go out of scope
  // reciprocal space: largest |l| per axis and the unit wave numbers, with the reference's expression (l = 1)
  for (int a = 0; a < 3; a++) {
    c.kmax[a] = 0;
    c.kunit[a] = 1 * 2 * kPi / P.ebox[a];
  }
  for (int k = 0; k < h->nk; k++)
    for (int a = 0; a < 3; a++) c.kmax[a] = std::max(c.kmax[a], std::abs(h->h_kl[4 * k + a]));
  c.valid = true;
  return PG_OK;
}

int chain_reserve_io(pg_engine* h, int max_steps) {
  PgChainHost& c = h->ch;
  if ((size_t)max_steps > c.log_cap || !c.d_log) {
    cudaFree(c.d_log); c.d_log = nullptr; c.log_cap = 0;
    const size_t cap = std::max<size_t>((size_t)max_steps, 1024);
    PG_CUDA(h, cudaMalloc((void**)&c.d_log, sizeof(PgChainRec) * cap));
    c.log_cap = cap;
  }
  if (c.want_trials) {
    const size_t need = (size_t)max_steps * (size_t)c.max_len * 3;
    if (need > c.tl_cap || !c.d_trial_log) {
      cudaFree(c.d_trial_log); c.d_trial_log = nullptr; c.tl_cap = 0;
      PG_CUDA(h, cudaMalloc((void**)&c.d_trial_log, sizeof(double) * std::max<size_t>(need, 1)));
      c.tl_cap = std::max<size_t>(need, 1);
    }
  }
  return PG_OK;
}

void chain_fill_args(pg_engine* h, int max_steps, PgChainArgs& A) {
  PgChainHost& c = h->ch;
  memset(&A, 0, sizeof(A));
  fill_move_dev(h, A.D);
  A.Pg = h->d_P;
  A.xy = h->xy; A.zq = h->zq; A.type = h->type; A.n = h->n;
  A.mol_first = c.d_mol_first; A.chains = c.d_chains; A.ions = c.d_ions;
  A.cfg = c.cfg;
  A.kl = reinterpret_cast<const int4*>(h->d_kl); A.ek2 = h->d_ek2; A.S = h->d_S; A.dS = h->d_dS; A.nk = h->P.use_ewald ? h->nk : 0;
  for (int a = 0; a < 3; a++) { A.kmax[a] = c.kmax[a]; A.kunit[a] = c.kunit[a]; A.nc[a] = c.nc[a]; }
  A.cell_cap = c.cell_cap;
  A.nq_tot = c.nq_tot; A.qpos = c.d_qpos; A.qfrac = c.d_qfrac; A.qslot = c.d_qslot;
  A.cell_slots = c.d_cell_slots; A.ovf = c.d_ovf; A.ovf_n = c.d_ovf_n; A.bead_cell = c.d_bead_cell; A.bead_slot = c.d_bead_slot;
  A.mt_io = c.d_mt; A.log = c.d_log;
  A.trial_log = c.want_trials ? c.d_trial_log : nullptr;
  A.trial_stride = c.max_len;
  A.state = h->d_state; A.out = c.d_out; A.max_steps = max_steps; A.pivot_mode = c.pivot_mode;
  // side-by-side phases need the CTA's k slice to fit the four reciprocal-space warps
  {
    const int nkc = (A.nk + c.cluster - 1) / c.cluster;
    static const bool no_spec = [] { const char* e = getenv("PLUM_B200_CHAIN_NO_SPEC"); return e && e[0] == '1'; }();
    A.specialize = (c.cluster > 1 && !no_spec && (nkc + 127) / 128 <= CH_KPT) ? 1 : 0;
  }
  A.prof = c.d_prof;
  A.dbg_skip = c.dbg_skip;
  A.counters = c.d_counters;
}

// What a chain launch needs from every engine it carries.
int chain_prepare(pg_engine* h, int max_steps) {
  PgChainHost& c = h->ch;
  if (!c.configured) { h->err = "pg_chain_configure has not been called"; return PG_ERR_STATE; }
  if (c.inflight || h->mc_inflight || h->inflight) { h->err = "another batch or trial is in flight"; return PG_ERR_STATE; }
  if (h->pending) { h->err = "previous trial not committed"; return PG_ERR_STATE; }
  if (max_steps < 0) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  int rc = flush_commit(h);
  if (rc) return rc;
  if (!c.valid) { rc = chain_build(h); if (rc) return rc; }
  rc = chain_reserve_io(h, std::max(max_steps, 1));
  if (rc) return rc;
  const int nk = h->P.use_ewald ? h->nk : 0;
  if (((nk + c.cluster - 1) / c.cluster + 448 - 1) / 448 > CH_KPT) { h->err = "chain: too many k vectors per CTA (raise the cluster size)"; return PG_ERR_CAPACITY; }
  if ((c.kmax[0] + c.kmax[1] + c.kmax[2] + 3) > CH_TAB) { h->err = "chain: k table does not fit"; return PG_ERR_CAPACITY; }
  return chain_kernel_setup(h, c.cluster);
}

// What a fleet hands back, gathered per chain into one block (pg_chain_run_multi_io): generator, status words, step
// log, coordinates as [n_beads][3].  Grid (chains, 16), 256 threads.
__global__ void __launch_bounds__(256) k_chain_pack(const PgChainArgs* __restrict__ args, char* __restrict__ stage, size_t per,
                                                    int log_steps, int nb) {
  const PgChainArgs& A = args[blockIdx.x];
  char* dst = stage + per * (size_t)blockIdx.x;
  const int t = blockIdx.y * 256 + threadIdx.x, T = gridDim.y * 256;
  uint32_t* d_rng = reinterpret_cast<uint32_t*>(dst);
  for (int w = t; w < 625; w += T) d_rng[w] = A.mt_io[w];
  int* d_out = reinterpret_cast<int*>(dst + 640 * sizeof(uint32_t));
  for (int w = t; w < 4; w += T) d_out[w] = A.out[w];
  int4* d_log = reinterpret_cast<int4*>(dst + 640 * sizeof(uint32_t) + 16);
  const int4* s_log = reinterpret_cast<const int4*>(A.log);
  for (int w = t; w < log_steps; w += T) d_log[w] = s_log[w];
  double* d_xyz = reinterpret_cast<double*>(dst + 640 * sizeof(uint32_t) + 16 + sizeof(PgChainRec) * (size_t)log_steps);
  for (int b = t; b < nb; b += T) {
    const double2 a = A.xy[b], c = A.zq[b];
    d_xyz[3 * b] = a.x; d_xyz[3 * b + 1] = a.y; d_xyz[3 * b + 2] = c.x;
  }
}

// Which build of k_chain a launch takes (pg_chain.cu): clusters and fleets of at most one chain per SM run the
// 512-thread kernel (shortest step); more chains than SMs run the 448-thread kernel, two CTAs per SM.
// PLUM_B200_CHAIN_THREADS=448 / 512 forces one of them for single-CTA chains (tests, A/B measurements).
int chain_threads_for(const pg_engine* lead, int n_chains, int cluster) {
  if (cluster > 1) return 512;
  const char* env = getenv("PLUM_B200_CHAIN_THREADS");   // read per launch: tests switch it between chains
  const int forced = env ? atoi(env) : 0;
  if (forced == 448 || forced == 512) return forced;
  return (n_chains > std::max(1, lead->n_sm)) ? 448 : 512;
}

int chain_launch(pg_engine* lead, const PgChainArgs* d_args, int n_chains, int cluster) {
  const int threads = chain_threads_for(lead, n_chains, cluster);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(n_chains * cluster), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (threads == 448) ? sizeof(ChSmem448) : sizeof(ChSmem512);
  cfg.stream = lead->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (threads == 448) PG_CUDA(lead, cudaLaunchKernelEx(&cfg, k_chain448, d_args));
  else PG_CUDA(lead, cudaLaunchKernelEx(&cfg, k_chain512, d_args));
  lead->launches++;
  return PG_OK;
}

const char* chain_err_text(int e) {
  switch (e) {
    case CH_ERR_RNG: return "chain kernel: a rejection loop ran past the random-stream window";
    case CH_ERR_LEN: return "chain kernel: molecule length not supported";
    case CH_ERR_KIND: return "chain kernel: move kind not offered on the device (crankshaft)";
    case CH_ERR_OVERFLOW: return "chain kernel: cell overflow list is full";
    case CH_ERR_K: return "chain kernel: reciprocal-space capacity exceeded";
    default: return "chain kernel: unknown error";
  }
}

}  // namespace

extern "C" {

int pg_chain_configure(pg_engine* h, const pg_chain_config* cfg) {
  if (!h || !cfg) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  if (!h->fast) {
    h->err = "pg_chain_*: needs a single-image system (3 periodic axes, one box, real-space cutoff < L/2); use pg_mc_* / pg_delta_e";
    return PG_ERR_INVALID;
  }
  if (cfg->cluster_ctas < 0 || cfg->cluster_ctas > CH_GMAX) { h->err = "pg_chain_*: cluster_ctas must be 0 (default) or 1..16"; return PG_ERR_INVALID; }
  if (cfg->phantom < 0 || cfg->gc_freq < 0) return PG_ERR_INVALID;
  c.cluster = cfg->cluster_ctas > 0 ? cfg->cluster_ctas : 1;
  if (!c.configured || c.phantom != cfg->phantom) c.valid = false;   // the movable-molecule lists depend on it
  c.phantom = cfg->phantom;
  c.cfg.gc_freq = cfg->gc_freq;
  c.cfg.vary_bond = cfg->vary_bond ? 1 : 0;
  c.cfg.move_size = cfg->move_size;
  c.cfg.bond_len = cfg->bond_len;
  for (int i = 0; i < 5; i++) c.cfg.prob[i] = cfg->move_prob[i];
  c.want_trials = cfg->keep_trials != 0;
  if (cfg->pivot_mode < 0 || cfg->pivot_mode > 2) { h->err = "pg_chain_*: pivot_mode must be 0, 1 or 2"; return PG_ERR_INVALID; }
  c.pivot_mode = cfg->pivot_mode;
  c.configured = true;
  return PG_OK;
}

int pg_chain_set_rng(pg_engine* h, const uint32_t* state624, int position) {
  if (!h || !state624 || position < 0 || position > CG_N) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  if (!c.d_mt) {
    PG_CUDA(h, cudaMalloc((void**)&c.d_mt, sizeof(uint32_t) * (CG_N + 8)));
    PG_CUDA(h, cudaMalloc((void**)&c.d_out, sizeof(int) * 4));
    PG_CUDA(h, cudaMemset(c.d_out, 0, sizeof(int) * 4));
  }
  uint32_t buf[CG_N + 1];
  memcpy(buf, state624, sizeof(uint32_t) * CG_N);
  buf[CG_N] = (uint32_t)position;
  PG_CUDA(h, cudaMemcpyAsync(c.d_mt, buf, sizeof(buf), cudaMemcpyHostToDevice, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  return PG_OK;
}

int pg_chain_get_rng(pg_engine* h, uint32_t* state624, int* position) {
  if (!h || !state624 || !position) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  if (!c.d_mt) { h->err = "no random stream has been set"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  uint32_t buf[CG_N + 1];
  PG_CUDA(h, cudaMemcpyAsync(buf, c.d_mt, sizeof(buf), cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  memcpy(state624, buf, sizeof(uint32_t) * CG_N);
  *position = (int)buf[CG_N];
  return PG_OK;
}

int pg_chain_begin(pg_engine* h, int max_steps) {
  if (!h) return PG_ERR_INVALID;
  int rc = chain_prepare(h, max_steps);
  if (rc) return rc;
  PgChainHost& c = h->ch;
  if (c.args_cap < 1) {
    cudaFree(c.d_args); c.d_args = nullptr; c.args_cap = 0;
    PG_CUDA(h, cudaMalloc((void**)&c.d_args, sizeof(PgChainArgs)));
    c.args_cap = 1;
  }
  PgChainArgs A;
  chain_fill_args(h, max_steps, A);
  PG_CUDA(h, cudaMemsetAsync(c.d_out, 0, sizeof(int) * 4, h->stream));
  PG_CUDA(h, cudaMemcpyAsync(c.d_args, &A, sizeof(A), cudaMemcpyHostToDevice, h->stream));   // pageable: staged before return
  PG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  rc = chain_launch(h, c.d_args, 1, c.cluster);
  if (rc) return rc;
  PG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  c.inflight = true;
  c.max_steps = max_steps;
  return PG_OK;
}

int pg_chain_end(pg_engine* h, int* n_done, int* stop_kind, pg_chain_step* steps, float* elapsed_ms) {
  if (!h) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  if (!c.inflight) { h->err = "no chain in flight"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  c.inflight = false;
  int out[4] = {0, 0, 0, 0};
  PG_CUDA(h, cudaMemcpyAsync(out, c.d_out, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (elapsed_ms) PG_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
  if (out[2] != 0) {
    h->err = chain_err_text(out[2]);
    c.valid = false;
    return (out[2] == CH_ERR_OVERFLOW || out[2] == CH_ERR_K || out[2] == CH_ERR_LEN) ? PG_ERR_CAPACITY : PG_ERR_STATE;
  }
  if (out[0] < 0 || out[0] > c.max_steps) { h->err = "chain: inconsistent step count"; return PG_ERR_STATE; }
  if (n_done) *n_done = out[0];
  if (stop_kind) *stop_kind = out[1];
  if (steps && out[0] > 0) {
    static_assert(sizeof(pg_chain_step) == sizeof(PgChainRec), "pg_chain_step layout");
    PG_CUDA(h, cudaMemcpy(steps, c.d_log, sizeof(PgChainRec) * (size_t)out[0], cudaMemcpyDeviceToHost));
  }
  return PG_OK;
}

int pg_chain_run(pg_engine* h, int max_steps, int* n_done, int* stop_kind, pg_chain_step* steps, float* elapsed_ms) {
  int rc = pg_chain_begin(h, max_steps);
  if (rc) return rc;
  return pg_chain_end(h, n_done, stop_kind, steps, elapsed_ms);
}

// Many independent replicas in ONE launch (one cluster each) on the first engine's stream: a single host thread keeps
// the whole GPU busy.  All engines must live on the same device and use the same cluster size.
int pg_chain_run_multi(pg_engine** hs, int n, int max_steps, int* n_done, float* elapsed_ms) {
  if (!hs || n <= 0 || !hs[0]) return PG_ERR_INVALID;
  pg_engine* lead = hs[0];
  const int cluster = lead->ch.cluster;
  for (int i = 0; i < n; i++) {
    if (!hs[i]) return PG_ERR_INVALID;
    if (hs[i]->device != lead->device || hs[i]->ch.cluster != cluster) { lead->err = "pg_chain_run_multi: engines must share device and cluster size"; return PG_ERR_INVALID; }
    int rc = chain_prepare(hs[i], max_steps);
    if (rc) { if (hs[i] != lead) lead->err = hs[i]->err; return rc; }
  }
  PgChainHost& c = lead->ch;
  if (c.args_cap < (size_t)n) {
    cudaFree(c.d_args); c.d_args = nullptr; c.args_cap = 0;
    PG_CUDA(lead, cudaMalloc((void**)&c.d_args, sizeof(PgChainArgs) * (size_t)n));
    c.args_cap = (size_t)n;
  }
  std::vector<PgChainArgs> args((size_t)n);
  for (int i = 0; i < n; i++) {
    chain_fill_args(hs[i], max_steps, args[i]);
    // the replicas' own streams may still hold earlier work (uploads, initialisation)
    if (hs[i] != lead) PG_CUDA(lead, cudaStreamSynchronize(hs[i]->stream));
    PG_CUDA(lead, cudaMemsetAsync(hs[i]->ch.d_out, 0, sizeof(int) * 4, lead->stream));
  }
  PG_CUDA(lead, cudaMemcpyAsync(c.d_args, args.data(), sizeof(PgChainArgs) * (size_t)n, cudaMemcpyHostToDevice, lead->stream));
  PG_CUDA(lead, cudaEventRecord(lead->ev0, lead->stream));
  int rc = chain_launch(lead, c.d_args, n, cluster);
  if (rc) return rc;
  PG_CUDA(lead, cudaEventRecord(lead->ev1, lead->stream));
  std::vector<int> out(4 * (size_t)n);
  for (int i = 0; i < n; i++)
    PG_CUDA(lead, cudaMemcpyAsync(out.data() + 4 * i, hs[i]->ch.d_out, sizeof(int) * 4, cudaMemcpyDeviceToHost, lead->stream));
  PG_CUDA(lead, cudaStreamSynchronize(lead->stream));
  if (elapsed_ms) PG_CUDA(lead, cudaEventElapsedTime(elapsed_ms, lead->ev0, lead->ev1));
  for (int i = 0; i < n; i++) {
    hs[i]->ch.max_steps = max_steps;
    if (out[4 * i + 2] != 0) {
      hs[i]->err = chain_err_text(out[4 * i + 2]);
      hs[i]->ch.valid = false;
      lead->err = hs[i]->err;
      return PG_ERR_STATE;
    }
    if (n_done) n_done[i] = out[4 * i];
  }
  return PG_OK;
}

// The whole boundary crossing of a batch in ONE call and ONE synchronisation (the end-to-end path of a driver that keeps
// the generators and the beads on the host): generator states down, one launch, step logs, generator states and (if xyz
// is not NULL) the accepted coordinates up — all through one pinned staging block of the first engine.
//   rng_io   [n][625]  in: state words + position of every replica's std::mt19937; out: the state behind the last step
//   rng_up   [n] (may be NULL = all): 0 = replica i continues from the state resident on the device (the caller's
//            generator has not been touched since the last call handed it back)
//   steps    [n][max_steps] records (may be NULL);  xyz [n] pointers to [n_beads][3] (may be NULL, entries may be NULL)
int pg_chain_run_multi_io(pg_engine** hs, int n, int max_steps, uint32_t* rng_io, const uint8_t* rng_up, pg_chain_step* steps,
                          double* const* xyz, int* n_done, float* elapsed_ms) {
  if (!hs || n <= 0 || !hs[0] || !rng_io) return PG_ERR_INVALID;
  pg_engine* lead = hs[0];
  const int cluster = lead->ch.cluster;
  const int nb = lead->n;
  for (int i = 0; i < n; i++) {
    if (!hs[i]) return PG_ERR_INVALID;
    if (hs[i]->device != lead->device || hs[i]->ch.cluster != cluster) { lead->err = "pg_chain_run_multi_io: engines must share device and cluster size"; return PG_ERR_INVALID; }
    if (xyz && hs[i]->n != nb) { lead->err = "pg_chain_run_multi_io: engines must hold the same number of beads"; return PG_ERR_INVALID; }
    if ((!rng_up || rng_up[i]) && rng_io[(size_t)i * 625 + 624] > CG_N) { lead->err = "pg_chain_run_multi_io: generator position out of range"; return PG_ERR_INVALID; }
    int rc = chain_prepare(hs[i], max_steps);
    if (rc) { if (hs[i] != lead) lead->err = hs[i]->err; return rc; }
  }
  PgChainHost& c = lead->ch;
  // what comes back, per chain and in this order: rng 625 u32 (padded to 640) | out 4 int | log max_steps x 16 B |
  // xyz nb x 24 B.  A small kernel gathers the pieces of all chains into one device block (k_chain_pack), ONE copy
  // brings it into pinned memory, a few host threads hand it to the caller's arrays.
  const size_t rng_b = 640 * sizeof(uint32_t), out_b = 16, log_b = steps ? sizeof(PgChainRec) * (size_t)std::max(max_steps, 1) : 0;
  const size_t pos_b = xyz ? ((sizeof(double) * 3 * (size_t)std::max(nb, 1) + 15) & ~(size_t)15) : 0;
  const size_t per = rng_b + out_b + log_b + pos_b;
  if (per * (size_t)n > c.pin_cap) {
    if (c.h_pin) cudaFreeHost(c.h_pin);
    cudaFree(c.d_pack);
    c.h_pin = nullptr; c.d_pack = nullptr; c.pin_cap = 0;
    PG_CUDA(lead, cudaHostAlloc((void**)&c.h_pin, per * (size_t)n, cudaHostAllocDefault));
    PG_CUDA(lead, cudaMalloc((void**)&c.d_pack, per * (size_t)n));
    c.pin_cap = per * (size_t)n;
  }
  if (c.args_cap < (size_t)n) {
    cudaFree(c.d_args); c.d_args = nullptr; c.args_cap = 0;
    PG_CUDA(lead, cudaMalloc((void**)&c.d_args, sizeof(PgChainArgs) * (size_t)n));
    c.args_cap = (size_t)n;
  }
  std::vector<PgChainArgs> args((size_t)n);
  for (int i = 0; i < n; i++) {
    char* base = c.h_pin + per * (size_t)i;
    chain_fill_args(hs[i], max_steps, args[i]);
    if (hs[i] != lead) PG_CUDA(lead, cudaStreamSynchronize(hs[i]->stream));
    if (!rng_up || rng_up[i]) {
      memcpy(base, rng_io + (size_t)i * 625, 625 * sizeof(uint32_t));
      PG_CUDA(lead, cudaMemcpyAsync(hs[i]->ch.d_mt, base, 625 * sizeof(uint32_t), cudaMemcpyHostToDevice, lead->stream));
    }
    PG_CUDA(lead, cudaMemsetAsync(hs[i]->ch.d_out, 0, sizeof(int) * 4, lead->stream));
  }
  PG_CUDA(lead, cudaMemcpyAsync(c.d_args, args.data(), sizeof(PgChainArgs) * (size_t)n, cudaMemcpyHostToDevice, lead->stream));
  PG_CUDA(lead, cudaEventRecord(lead->ev0, lead->stream));
  int rc = chain_launch(lead, c.d_args, n, cluster);
  if (rc) return rc;
  PG_CUDA(lead, cudaEventRecord(lead->ev1, lead->stream));
  k_chain_pack<<<dim3((unsigned)n, 16), 256, 0, lead->stream>>>(c.d_args, c.d_pack, per, (int)(log_b / sizeof(PgChainRec)), xyz ? nb : 0);
  lead->launches++;
  PG_CUDA(lead, cudaGetLastError());
  PG_CUDA(lead, cudaMemcpyAsync(c.h_pin, c.d_pack, per * (size_t)n, cudaMemcpyDeviceToHost, lead->stream));
  PG_CUDA(lead, cudaStreamSynchronize(lead->stream));
  if (elapsed_ms) PG_CUDA(lead, cudaEventElapsedTime(elapsed_ms, lead->ev0, lead->ev1));
  for (int i = 0; i < n; i++) {
    const char* base = c.h_pin + per * (size_t)i;
    const int* out = reinterpret_cast<const int*>(base + rng_b);
    hs[i]->ch.max_steps = max_steps;
    if (out[2] != 0) {
      hs[i]->err = chain_err_text(out[2]);
      hs[i]->ch.valid = false;
      lead->err = hs[i]->err;
      return PG_ERR_STATE;
    }
    if (out[0] < 0 || out[0] > max_steps) { lead->err = "chain: inconsistent step count"; return PG_ERR_STATE; }
    if (n_done) n_done[i] = out[0];
  }
  // pinned block -> the caller's arrays: memory-bound host work, spread over a few threads when a whole fleet comes back
  auto unpack = [&](int i0, int i1) {
    for (int i = i0; i < i1; i++) {
      const char* base = c.h_pin + per * (size_t)i;
      const int done = reinterpret_cast<const int*>(base + rng_b)[0];
      memcpy(rng_io + (size_t)i * 625, base, 625 * sizeof(uint32_t));
      if (steps && done > 0) memcpy(steps + (size_t)i * max_steps, base + rng_b + out_b, sizeof(PgChainRec) * (size_t)done);
      if (xyz && xyz[i]) memcpy(xyz[i], base + rng_b + out_b + log_b, sizeof(double) * 3 * (size_t)nb);
    }
  };
  int T = 1;
  if (xyz && (long long)n * nb >= 1000000) T = (int)std::min<unsigned>(std::min<unsigned>(8u, (unsigned)n), std::max(1u, std::thread::hardware_concurrency() / 2));
  if (T <= 1) {
    unpack(0, n);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(unpack, (int)((long long)n * t / T), (int)((long long)n * (t + 1) / T));
    unpack(0, (int)((long long)n / T));
    for (auto& x : th) x.join();
  }
  return PG_OK;
}

// The log of the last chain of this engine (also after pg_chain_run_multi).
int pg_chain_steps(pg_engine* h, int first, int count, pg_chain_step* steps) {
  if (!h || !steps || first < 0 || count < 0) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  if ((size_t)(first + count) > c.log_cap) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  if (count > 0) PG_CUDA(h, cudaMemcpy(steps, c.d_log + first, sizeof(PgChainRec) * (size_t)count, cudaMemcpyDeviceToHost));
  return PG_OK;
}

int pg_chain_trial_xyz(pg_engine* h, int step, double* xyz, int n_beads) {
  if (!h || !xyz || step < 0 || n_beads < 0) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  if (!c.want_trials || !c.d_trial_log || step >= c.max_steps || n_beads > c.max_len) { h->err = "trial coordinates were not kept"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaMemcpy(xyz, c.d_trial_log + (size_t)step * c.max_len * 3, sizeof(double) * 3 * (size_t)n_beads, cudaMemcpyDeviceToHost));
  return PG_OK;
}

// Instrumentation (not part of the ABI header; tools/chain_probe.py): per-phase clock sums of the next chains of this
// engine ([5 kinds][CH_NPHASE] unsigned 64-bit, read back with pgx_chain_prof_read) and a mask of phases to leave out.
int pgx_chain_prof(pg_engine* h, int enable, int skip_mask) {
  if (!h) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  PG_CUDA(h, cudaSetDevice(h->device));
  if (enable && !c.d_prof) {
    PG_CUDA(h, cudaMalloc((void**)&c.d_prof, sizeof(unsigned long long) * 5 * CH_NPHASE));
    PG_CUDA(h, cudaMemset(c.d_prof, 0, sizeof(unsigned long long) * 5 * CH_NPHASE));
  }
  if (!enable && c.d_prof) { cudaFree(c.d_prof); c.d_prof = nullptr; }
  c.dbg_skip = skip_mask;
  return PG_OK;
}
int pgx_chain_prof_read(pg_engine* h, unsigned long long* out) {
  if (!h || !out || !h->ch.d_prof) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  PG_CUDA(h, cudaMemcpy(out, h->ch.d_prof, sizeof(unsigned long long) * 5 * CH_NPHASE, cudaMemcpyDeviceToHost));
  return PG_OK;
}

// Work counters of this engine's chains since the last reset: pair configurations evaluated inside a cutoff, steps that
// evaluated an energy change, accepted steps.
int pg_chain_counters(pg_engine* h, uint64_t* out3, int reset) {
  if (!h || !out3) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  out3[0] = out3[1] = out3[2] = 0;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  if (!c.d_counters) return PG_OK;
  PG_CUDA(h, cudaSetDevice(h->device));
  unsigned long long v[4];
  PG_CUDA(h, cudaMemcpy(v, c.d_counters, sizeof(v), cudaMemcpyDeviceToHost));
  out3[0] = v[0]; out3[1] = v[1]; out3[2] = v[2];
  if (reset) PG_CUDA(h, cudaMemset(c.d_counters, 0, sizeof(unsigned long long) * 4));
  return PG_OK;
}

// Consistency of the resident structures with the resident coordinates (tests): every bead filed exactly once, in the
// cell its coordinates map to; no stale entries; the compact charged records equal the bead arrays.
int pg_chain_check(pg_engine* h, int* n_bad) {
  if (!h || !n_bad) return PG_ERR_INVALID;
  PgChainHost& c = h->ch;
  *n_bad = 0;
  if (c.inflight) { h->err = "a chain is in flight"; return PG_ERR_STATE; }
  if (!c.valid) { h->err = "chain structures are not built"; return PG_ERR_STATE; }
  PG_CUDA(h, cudaSetDevice(h->device));
  const PgDev& P = h->P;
  const int n = h->n;
  const size_t n_cells = (size_t)c.nc[0] * c.nc[1] * c.nc[2];
  std::vector<double2> hxy(std::max(n, 1)), hzq(std::max(n, 1)), qpos(2 * (size_t)std::max(c.nq_tot, 1));
  std::vector<float4> qfrac(std::max(c.nq_tot, 1));
  const int cap = c.cell_cap;
  std::vector<int> slots((size_t)cap * n_cells), ovf(CH_OVF_CAP), bead_cell(std::max(n, 1)), bead_slot(std::max(n, 1)), qslot(std::max(n, 1));
  int ovf_n = 0;
  PG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (n > 0) {
    PG_CUDA(h, cudaMemcpy(hxy.data(), h->xy, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost));
    PG_CUDA(h, cudaMemcpy(hzq.data(), h->zq, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost));
    PG_CUDA(h, cudaMemcpy(bead_cell.data(), c.d_bead_cell, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    PG_CUDA(h, cudaMemcpy(bead_slot.data(), c.d_bead_slot, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    PG_CUDA(h, cudaMemcpy(qslot.data(), c.d_qslot, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
  }
  PG_CUDA(h, cudaMemcpy(slots.data(), c.d_cell_slots, sizeof(int) * slots.size(), cudaMemcpyDeviceToHost));
  PG_CUDA(h, cudaMemcpy(ovf.data(), c.d_ovf, sizeof(int) * CH_OVF_CAP, cudaMemcpyDeviceToHost));
  PG_CUDA(h, cudaMemcpy(&ovf_n, c.d_ovf_n, sizeof(int), cudaMemcpyDeviceToHost));
  if (c.nq_tot > 0) {
    PG_CUDA(h, cudaMemcpy(qpos.data(), c.d_qpos, sizeof(double2) * 2 * (size_t)c.nq_tot, cudaMemcpyDeviceToHost));
    PG_CUDA(h, cudaMemcpy(qfrac.data(), c.d_qfrac, sizeof(float4) * (size_t)c.nq_tot, cudaMemcpyDeviceToHost));
  }
  int bad = 0;
  std::string first;
  auto flag = [&](const std::string& what) { if (bad++ == 0) first = what; };
  if (P.pair_kind == PG_PAIR_TRUNCATED_LJ) {
    std::vector<int> seen(std::max(n, 1), 0);
    for (size_t s = 0; s < slots.size(); s++) {
      const int b = slots[s];
      if (b < 0) continue;
      if (b >= n) { flag("cell slot holds a bead index out of range"); continue; }
      seen[b]++;
      if (bead_cell[b] != (int)(s / cap) || bead_slot[b] != (int)(s % cap)) flag("cell slot disagrees with the bead's own record, bead " + std::to_string(b));
    }
    if (ovf_n < 0 || ovf_n > CH_OVF_CAP) flag("overflow count out of range");
    for (int o = 0; o < std::min(std::max(ovf_n, 0), CH_OVF_CAP); o++) {
      const int b = ovf[o];
      if (b < 0) continue;
      if (b >= n) { flag("overflow entry out of range"); continue; }
      seen[b]++;
      if (bead_slot[b] != cap + o) flag("overflow entry disagrees with the bead's own record");
    }
    for (int i = 0; i < n; i++) {
      if (seen[i] != 1) flag("bead " + std::to_string(i) + " is filed " + std::to_string(seen[i]) + " times");
      if (bead_cell[i] != chain_cell_of(c, P, hxy[i].x, hxy[i].y, hzq[i].x)) flag("bead " + std::to_string(i) + " sits in the wrong cell");
    }
  }
  if (P.use_ewald) {
    int nq = 0;
    for (int i = 0; i < n; i++) {
      if (hzq[i].y == 0.0) { if (qslot[i] != -1) flag("neutral bead has a charged-list slot"); continue; }
      const int s = qslot[i];
      if (s != nq) flag("charged-list slot out of order");
      nq++;
      if (s < 0 || s >= c.nq_tot) continue;
      int idx;
      memcpy(&idx, &qfrac[s].w, sizeof(int));
      if (idx != i) flag("charged record points at another bead");
      if (memcmp(&qpos[2 * s], &hxy[i], sizeof(double2)) || memcmp(&qpos[2 * s + 1], &hzq[i], sizeof(double2))) flag("charged record differs from the bead arrays, bead " + std::to_string(i));
      // (equal up to a whole box length: a coordinate at exactly half a cell may round either way)
      auto same = [](float a, float b) { const float d = fabsf(a - b); return d <= 1e-6f || fabsf(d - 1.0f) <= 1e-6f; };
      if (!same(qfrac[s].x, chain_frac(hxy[i].x, P.inv_box[0])) || !same(qfrac[s].y, chain_frac(hxy[i].y, P.inv_box[1])) ||
          !same(qfrac[s].z, chain_frac(hzq[i].x, P.inv_box[2]))) flag("FP32 box fraction differs, bead " + std::to_string(i));
    }
    if (nq != c.nq_tot) flag("charged count differs");
  }
  *n_bad = bad;
  if (bad) h->err = "pg_chain_check: " + first;
  return PG_OK;
}

}  // extern "C"
