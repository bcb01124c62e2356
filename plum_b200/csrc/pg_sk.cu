// plum_b200 — the full structure factor S(k) = sum_j q_j e^{i k.r_j} of the resident configuration
// (PotentialEwald::EnergyInitialization in structure-factor form, src/force_field/potential_ewald.cc:176-230; k tables of
// src/force_field/potential_ewald_coul.cc:77-115): energy initialisation, drift reset of the incrementally updated S(k),
// and the k-sharded recompute across GPUs.
//
//   k_sk_block   grid (tiles of k columns, chunks of the COMPACT list of charged beads).  A CTA stages 64 charged beads at a
//                time as per-axis phase tables e^{i l theta} — ONE sincos per bead and axis, then complex recurrences over
//                l <= kmax — in shared memory; a thread owns up to 16 consecutive lz of one (lx, ly) column: per bead one
//                complex product for the column and ONE table load + 4 FMA per k vector (instead of a sincos per
//                (bead, k)).  One writer per (chunk, k).
//   k_sk_finish  adds the chunks in index order (a k slice computed alone is bit-identical to the same rows of the full
//                result, whatever the slicing) and, when the recompute is sharded over GPUs, is the collective: it
//                stores the finished slice straight into every peer's S(k) through peer-mapped pointers (NVLink), raises
//                a per-rank flag in the peers' flag blocks behind a system fence, and waits for the other ranks' flags —
//                no NCCL call, no host in the loop.  An entry handshake (`ready` flags) keeps a fast rank from writing into
//                a peer that is still using its S(k).
#ifndef PLUM_B200_PG_SK_CU_
#define PLUM_B200_PG_SK_CU_

#define SK_THREADS 64       // k_sk_finish
#define SK_ITHREADS 128     // k_sk_block: work items (k columns) per CTA
#define SK_SEG 16           // k vectors (consecutive lz of one (lx, ly) column) one thread owns
#define SK_BEADS 32
#define SK_MAX_PEERS 16

// One work item of k_sk_block: `n` consecutive entries k0 .. k0 + n - 1 of the k list, all with the same (lx, ly) and
// lz = lz0, lz0 + 1, ... (the list is in the reference's cube order: lx outer, ly, lz inner — potential_ewald_coul.cc:89-115).
struct __align__(16) PgSkItem {
  int lx, ly, lz0, k0;
};

struct PgSkArgs {
  const double2* xy; const double2* zq; const int* qidx; int nq;
  const PgSkItem* items; const int* item_n; int item_first, item_count;
  int k_first, k_count;
  int kmax[3]; double kunit[3]; double ebox[3], inv_ebox[3]; int pbc[3];
  int chunk;            // charged beads per chunk (a multiple of SK_BEADS; a function of nq only)
  double2* partial;     // [n_chunks][k_count]
};

// Shared-memory phase tables of one staged bead: q e^{i l theta_x} for l = 0 .. kmax_x, then e^{i l theta_y} for
// l = -kmax_y .. kmax_y, then e^{i l theta_z} for l = -kmax_z .. kmax_z (negative l stored explicitly: no conjugation in
// the inner loop).
__global__ void __launch_bounds__(SK_ITHREADS) k_sk_block(const PgSkArgs A) {
  extern __shared__ __align__(16) unsigned char sk_raw[];
  double2* tab = reinterpret_cast<double2*>(sk_raw);          // [SK_BEADS][ne]
  const int tid = threadIdx.x;
  const int ne0 = A.kmax[0] + 1, ne1 = 2 * A.kmax[1] + 1, ne2 = 2 * A.kmax[2] + 1, ne = ne0 + ne1 + ne2;
  const int item = blockIdx.x * SK_ITHREADS + tid;
  const bool act = item < A.item_count;
  PgSkItem I = {0, 0, 0, 0};
  int s_lo = 0, s_hi = 0;   // entries [s_lo, s_hi) of the item lie in the k slice
  if (act) {
    I = A.items[A.item_first + item];
    const int n = A.item_n[A.item_first + item];
    s_lo = max(0, A.k_first - I.k0);
    s_hi = min(n, A.k_first + A.k_count - I.k0);
  }
  double re[SK_SEG], im[SK_SEG];
#pragma unroll
  for (int s_ = 0; s_ < SK_SEG; s_++) { re[s_] = 0.0; im[s_] = 0.0; }
  const int c0 = blockIdx.y * A.chunk, c1 = min(A.nq, c0 + A.chunk);
  for (int b0 = c0; b0 < c1; b0 += SK_BEADS) {
    const int nb = min(SK_BEADS, c1 - b0);
    __syncthreads();
    for (int t = tid; t < 3 * nb; t += SK_ITHREADS) {
      const int bead = t / 3, ax = t - 3 * bead;
      const int j = __ldg(&A.qidx[b0 + bead]);
      const double2 c = A.zq[j];
      double x;
      if (ax == 2) x = c.x;
      else { const double2 a = A.xy[j]; x = ax == 0 ? a.x : a.y; }
      x = pg_wrap_pos(x, A.ebox[ax], A.inv_ebox[ax], A.pbc[ax]);
      double s1, c1_;
      sincos(A.kunit[ax] * x, &s1, &c1_);
      const int km = A.kmax[ax];
      // x: row[l] for l >= 0, scaled by the charge; y, z: centre entry at km, row[km + l] and row[km - l]
      double2* row = tab + bead * ne + (ax == 0 ? 0 : (ax == 1 ? ne0 + km : ne0 + ne1 + km));
      const double w = (ax == 0) ? c.y : 1.0;
      row[0] = make_double2(w, 0.0);
      double cr = c1_, sr = s1;
      for (int q = 1; q <= km; q++) {
        row[q] = make_double2(w * cr, w * sr);
        if (ax != 0) row[-q] = make_double2(cr, -sr);
        const double cn = cr * c1_ - sr * s1, sn = sr * c1_ + cr * s1;
        cr = cn; sr = sn;
      }
    }
    __syncthreads();
    if (s_hi > s_lo) {
      const int oy = ne0 + A.kmax[1] + I.ly, oz = ne0 + ne1 + A.kmax[2] + I.lz0;
      for (int bead = 0; bead < nb; bead++) {
        const double2* row = tab + bead * ne;
        const double2 a = row[I.lx], b = row[oy];
        const double abr = a.x * b.x - a.y * b.y, abi = a.x * b.y + a.y * b.x;
#pragma unroll
        for (int s_ = 0; s_ < SK_SEG; s_++) {
          if (s_ >= s_lo && s_ < s_hi) {
            const double2 c = row[oz + s_];
            re[s_] += abr * c.x - abi * c.y;
            im[s_] += abr * c.y + abi * c.x;
          }
        }
      }
    }
  }
#pragma unroll
  for (int s_ = 0; s_ < SK_SEG; s_++)
    if (s_ >= s_lo && s_ < s_hi)
      A.partial[(size_t)blockIdx.y * A.k_count + (I.k0 + s_ - A.k_first)] = make_double2(re[s_], im[s_]);
}

struct PgSkPeers {
  int world, rank;
  double2* S[SK_MAX_PEERS];          // every rank's S(k) (own entry: the local buffer)
  unsigned int* flags[SK_MAX_PEERS]; // every rank's flag block: [0 .. world) ready, [SK_MAX_PEERS .. SK_MAX_PEERS + world) done
  unsigned int seq;
};

// Flags are read and written with system-scope atomics: they are resolved at the memory's point of coherence, whichever
// GPU (or die) the polling kernel runs on.
__device__ __forceinline__ unsigned int sk_ld_sys(unsigned int* p) { return atomicAdd_system(p, 0u); }
__device__ __forceinline__ void sk_st_sys(unsigned int* p, unsigned int v) { atomicMax_system(p, v); }

// Entry handshake of a sharded recompute: "this rank no longer reads its S(k)" — ordered behind the rank's earlier work
// by the stream.
__global__ void k_sk_ready(PgSkPeers Pe) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < Pe.world; r++) sk_st_sys(Pe.flags[r] + Pe.rank, Pe.seq);
    __threadfence_system();
  }
}

// out: where the finished slice goes locally (the engine's S(k) + k_first, or a caller's buffer).  With peers the slice is
// also stored into every other rank's S(k); the last CTA then raises this rank's `done` flag everywhere and waits until
// every rank has raised its own here.
__global__ void __launch_bounds__(SK_THREADS) k_sk_finish(const double2* __restrict__ partial, int n_chunks, int k_count,
                                                          int k_first, double2* out, PgSkPeers Pe, unsigned int* counter) {
  const int kk = blockIdx.x * SK_THREADS + threadIdx.x;
  const bool peers = Pe.world > 1;
  if (peers && threadIdx.x < Pe.world) {
    // nobody may still be reading the S(k) we are about to overwrite (bounded wait: a rank that never shows up must
    // not hang the device — the host turns counter[1] into PG_ERR_TIMEOUT)
    unsigned int spins = 0;
    while (sk_ld_sys(Pe.flags[Pe.rank] + threadIdx.x) < Pe.seq) {
      __nanosleep(200);
      if (++spins > (1u << 23)) { counter[1] = 1u + threadIdx.x; break; }
    }
  }
  if (peers) __syncthreads();
  if (kk < k_count) {
    double re = 0.0, im = 0.0;
    int c = 0;
    for (; c + 8 <= n_chunks; c += 8) {   // eight loads in flight, added in chunk order
      double2 v8[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v8[u] = partial[(size_t)(c + u) * k_count + kk];
#pragma unroll
      for (int u = 0; u < 8; u++) { re += v8[u].x; im += v8[u].y; }
    }
    for (; c < n_chunks; c++) {
      const double2 v1 = partial[(size_t)c * k_count + kk];
      re += v1.x; im += v1.y;
    }
    const double2 v = make_double2(re, im);
    out[kk] = v;
    if (peers)
      for (int r = 0; r < Pe.world; r++)
        if (r != Pe.rank) Pe.S[r][k_first + kk] = v;
  }
  if (!peers) return;
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *counter = 0u;
  if (threadIdx.x == 0) {
    __threadfence_system();
    for (int r = 0; r < Pe.world; r++) sk_st_sys(Pe.flags[r] + SK_MAX_PEERS + Pe.rank, Pe.seq);
    __threadfence_system();
  }
  if (threadIdx.x < Pe.world) {
    unsigned int spins = 0;
    while (sk_ld_sys(Pe.flags[Pe.rank] + SK_MAX_PEERS + threadIdx.x) < Pe.seq) {
      __nanosleep(200);
      if (++spins > (1u << 23)) { counter[1] = 101u + threadIdx.x; break; }
    }
    __threadfence_system();
  }
}

#endif
