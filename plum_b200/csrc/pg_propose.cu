// plum_b200 — k_propose: trial coordinates of Plum's move generators built on the device
// from the resident coordinates.
//
// The random numbers are drawn on the host, ahead of time, in the reference's draw order
// (plum_b200/host/mc_propose.h): nothing a generator draws depends on the coordinates, only
// what it does with the draws.  One CTA per Markov-chain step; a launch covers a "wave" of
// consecutive steps that move pairwise different molecules, so their proposals are
// independent of each other's outcome and only the step in front of the wave (whose decision
// is still pending in PgState::accept) has to be looked at through the overlay.
//
//   BEAD        Molecule::BeadTranslate   src/molecules/molecule.cc:103-134
//   COM         Molecule::COMTranslate    :136-153
//   PIVOT       Molecule::Pivot           :155-237   (sequential by construction: bead i is placed
//               relative to the already placed bead i∓1 and drags the rest of its arm along; the
//               two arms are independent and run in two warps, every lane of a warp computes the
//               same step and the lanes share the arm's translation)
//   REPTATION   Molecule::RandomReptation :268-312
//   CRANKSHAFT  Molecule::Crankshaft      :239-265   (Eigen's AngleAxisd::toRotationMatrix order; the sine and
//               cosine of the angle are evaluated by the host generator's libm and travel in the descriptor)
//
// Arithmetic: pg_propose_math.h (separately rounded operations in the reference's order, so the
// coordinates are bit-identical to the host generator's and to the reference's).
#include "pg_kernels.cuh"
#include "pg_propose_math.h"

#define PP_THREADS 64
#define PP_MAXLEN 512   // longest molecule a device-side proposal handles (shared-memory staging)

enum { PP_BEAD = 0, PP_COM = 1, PP_PIVOT = 2, PP_CRANK = 3, PP_REPT = 4 };

// Device copy of one step's descriptor.
struct __align__(16) PgPropDev {
  int g0, glen, kind, i0;
  int off, rv_off, pad0, pad1;   // off: bead offset of the step's slot in the packed trial buffer
  double s, vx, vy, vz, vlen, pad2;
};

struct PgProposeArgs {
  const double2* xy;
  const double2* zq;
  const PgPropDev* moves;
  const double4* rv;
  double* trial;         // packed trial buffer ([beads][3]); step m writes rows [off, off + glen)
  int first;             // block b builds step first + b
  int prev;              // the step whose decision is pending when this wave starts, or -1
  const PgState* state;
  const int* stop;       // batch stopped (a step returned dE >= 1e8): nothing left to do
};

__global__ void __launch_bounds__(PP_THREADS) k_propose(const PgProposeArgs A) {
  if (__ldcg(A.stop)) return;
  __shared__ double sx[PP_MAXLEN], sy[PP_MAXLEN], sz[PP_MAXLEN];
  __shared__ double4 srv[PP_MAXLEN];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PgPropDev d = A.moves[A.first + blockIdx.x];
  const int len = d.glen;
  // current coordinates of the molecule, through the overlay of the pending step
  int pg0 = 0, pg1 = 0;
  const double* pt = nullptr;
  if (A.prev >= 0 && __ldcg(&A.state->accept)) {
    const PgPropDev p = A.moves[A.prev];
    pg0 = p.g0; pg1 = p.g0 + p.glen; pt = A.trial + 3 * (size_t)p.off;
  }
  for (int g = tid; g < len; g += PP_THREADS) {
    const int jg = d.g0 + g;
    if (jg >= pg0 && jg < pg1) {
      sx[g] = pt[3 * (jg - pg0)]; sy[g] = pt[3 * (jg - pg0) + 1]; sz[g] = pt[3 * (jg - pg0) + 2];
    } else {
      const double2 a = A.xy[jg], c = A.zq[jg];
      sx[g] = a.x; sy[g] = a.y; sz[g] = c.x;
    }
  }
  if (d.kind == PP_PIVOT)
    for (int r = tid; r < len - 1; r += PP_THREADS) srv[r] = A.rv[d.rv_off + r];
  __syncthreads();
  double* out = A.trial + 3 * (size_t)d.off;

  if (d.kind == PP_BEAD) {
    // only bead 0 moves (molecule.cc:114-119)
    for (int g = tid; g < len; g += PP_THREADS) {
      double x = sx[g], y = sy[g], z = sz[g];
      if (g == 0) { x = pp_bead_translate(x, d.s, d.vx); y = pp_bead_translate(y, d.s, d.vy); z = pp_bead_translate(z, d.s, d.vz); }
      out[3 * g] = x; out[3 * g + 1] = y; out[3 * g + 2] = z;
    }
  } else if (d.kind == PP_COM) {
    for (int g = tid; g < len; g += PP_THREADS) {
      out[3 * g] = pp_com_translate(sx[g], d.vx);
      out[3 * g + 1] = pp_com_translate(sy[g], d.vy);
      out[3 * g + 2] = pp_com_translate(sz[g], d.vz);
    }
  } else if (d.kind == PP_REPT) {
    const int dir = d.i0, end = (dir > 0) ? len - 1 : 0;
    for (int g = tid; g < len; g += PP_THREADS) {
      double x, y, z;
      if (g == end) {
        x = pp_reptation_end(sx[g], d.s, d.vx, d.vlen);
        y = pp_reptation_end(sy[g], d.s, d.vy, d.vlen);
        z = pp_reptation_end(sz[g], d.s, d.vz, d.vlen);
      } else {
        x = sx[g + dir]; y = sy[g + dir]; z = sz[g + dir];
      }
      out[3 * g] = x; out[3 * g + 1] = y; out[3 * g + 2] = z;
    }
  } else if (d.kind == PP_CRANK) {
    // Molecule::Crankshaft, molecule.cc:239-265: sin / cos of the angle come with the descriptor (vx, vy), the axis
    // and the rotation matrix from the resident coordinates in Eigen's operation order (pg_propose_math.h)
    const int first = d.i0, last = min(d.rv_off, len - 1);
    const double pf[3] = {sx[first], sy[first], sz[first]};
    const double pl[3] = {sx[last], sy[last], sz[last]};
    double rot[9];
    pp_crank_matrix(pf, pl, d.vx, d.vy, rot);
    for (int g = tid; g < len; g += PP_THREADS) {
      double o[3] = {sx[g], sy[g], sz[g]};
      if (g > first && g < last) {
        const double pos[3] = {sx[g], sy[g], sz[g]};
        pp_crank_apply(rot, pf, pos, o);
      }
      out[3 * g] = o[0]; out[3 * g + 1] = o[1]; out[3 * g + 2] = o[2];
    }
  } else if (d.kind == PP_PIVOT) {
    const int p = d.i0;
    const double msr = d.s;
    if (warp == 0 && p + 1 < len) {
      // forward arm: beads p+1 .. len-1, rows 0 .. len-2-p
      double a[3] = {sx[p], sy[p], sz[p]};
      double b[3] = {sx[p + 1], sy[p + 1], sz[p + 1]};
      for (int i = p + 1; i < len; i++) {
        const double4 r = srv[i - (p + 1)];
        double nb[3] = {0.0, 0.0, 0.0};
        if (i + 1 < len) { nb[0] = sx[i + 1]; nb[1] = sy[i + 1]; nb[2] = sz[i + 1]; }
        __syncwarp();   // every lane has read its copy before any lane's update loop writes these rows
        const double v[3] = {r.x, r.y, r.z};
        double m[3];
        pp_pivot_step(a, b, msr, v, r.w, m);
        for (int j = i + lane; j < len; j += 32) {
          sx[j] = PP_ADD(sx[j], m[0]); sy[j] = PP_ADD(sy[j], m[1]); sz[j] = PP_ADD(sz[j], m[2]);
        }
        __syncwarp();
        a[0] = PP_ADD(b[0], m[0]); a[1] = PP_ADD(b[1], m[1]); a[2] = PP_ADD(b[2], m[2]);
        b[0] = PP_ADD(nb[0], m[0]); b[1] = PP_ADD(nb[1], m[1]); b[2] = PP_ADD(nb[2], m[2]);
      }
    } else if (warp == 1 && p > 0) {
      // backward arm: beads p-1 .. 0, rows len-1-p .. len-2
      const int row0 = len - 1 - p;
      double a[3] = {sx[p], sy[p], sz[p]};
      double b[3] = {sx[p - 1], sy[p - 1], sz[p - 1]};
      for (int i = p - 1; i >= 0; i--) {
        const double4 r = srv[row0 + (p - 1 - i)];
        double nb[3] = {0.0, 0.0, 0.0};
        if (i > 0) { nb[0] = sx[i - 1]; nb[1] = sy[i - 1]; nb[2] = sz[i - 1]; }
        __syncwarp();
        const double v[3] = {r.x, r.y, r.z};
        double m[3];
        pp_pivot_step(a, b, msr, v, r.w, m);
        for (int j = i - lane; j >= 0; j -= 32) {
          sx[j] = PP_ADD(sx[j], m[0]); sy[j] = PP_ADD(sy[j], m[1]); sz[j] = PP_ADD(sz[j], m[2]);
        }
        __syncwarp();
        a[0] = PP_ADD(b[0], m[0]); a[1] = PP_ADD(b[1], m[1]); a[2] = PP_ADD(b[2], m[2]);
        b[0] = PP_ADD(nb[0], m[0]); b[1] = PP_ADD(nb[1], m[1]); b[2] = PP_ADD(nb[2], m[2]);
      }
    }
    __syncthreads();
    for (int g = tid; g < len; g += PP_THREADS) { out[3 * g] = sx[g]; out[3 * g + 1] = sy[g]; out[3 * g + 2] = sz[g]; }
  }
}
