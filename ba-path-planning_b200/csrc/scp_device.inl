// scp_device.inl -- the SCP solver for one scenario, executed by one CTA.
//
// Written "phase style": every SCP_PHASE block is a data-parallel loop over the
// CTA's threads and phases are separated by SCP_SYNC.  With SCP_EMU defined the
// same source compiles as plain C++ where a phase is a sequential loop over
// thread ids (tests/ only -- used to debug the kernel logic without a GPU; the
// product never loads it).
//
// What it replaces in the reference (src/path_planning/solvers/scp.py):
//   setup_scenario      _precompute_constraint_matrices bounds      :205-257
//   factor_operator     OSQP setup / KKT factorisation              :360, :442
//   admm_run            problem.solve()                             :362, :445
//   forward_rows        _accelerations_to_positions_velocities      :559-595
//   build_candidates    _add_collision_constraints (matrix free)    :453-557
//   gate_and_minsep     _fast_check_avoidance_constraints           :597-615
//   solve_scenario      generate_trajectories                       :131-180
//
// Decision vector: a[i][k][axis] (scp.py:15-26); internally one "agent-axis"
// q = 2 i + axis owns a K-vector stored contiguously: x[q*K + k].
//
// QP solver: ADMM on  min sum a^2  s.t.  A a = z, z in C, with
//   * the 2 terminal equalities per agent-axis handled exactly inside the
//     x-update (affine constraint of f),
//   * box rows (jerk, acc, vel, pos) in the (v = z + y/rho) form: one double of
//     state per row, z = clip(v), y = rho (v - z),
//   * collision rows eta.(p_i - p_j) >= bound split on per-agent copies of the
//     two positions; the copy multipliers stay on span{(eta,-eta)} so one scalar
//     lam >= 0 per row is the whole state, and the x-update operator is one
//     K x K matrix shared by every agent-axis of the scenario:
//        M = (2+sigma) I + rho (D'RjD + Ra + V'RvV + S'RpS) + rho c S'RcS
//     (c = copies = max candidate rows per (agent, step)).
// Only rows whose linearisation-point distance is below R + margin are carried
// ("candidates"); after convergence every dropped row is verified and the
// candidate set enlarged if one is violated, so the result is the minimiser of
// the full QP.  DESIGN.md derives all of this.

#ifndef SCP_DEVICE_INL
#define SCP_DEVICE_INL

#include <math.h>
#include <stdint.h>

#include "../../include/scp_b200.h"

#ifdef SCP_EMU
#define SCP_NANOS() 0LL
#define SCP_CLOCK() 0LL
#define SCP_DEV inline
#define SCP_PHASE(c) for (int tid = 0; tid < (c).nthreads; ++tid)
#define SCP_SYNC(c) ((void)0)
#define SCP_FMAX(a, b) ((a) > (b) ? (a) : (b))
#define SCP_FMIN(a, b) ((a) < (b) ? (a) : (b))
#else
#define SCP_CLOCK() clock64()
#define SCP_NANOS() scp_globaltimer()
#define SCP_DEV __device__ __forceinline__
#define SCP_PHASE(c) for (int tid = (c).tid0 + threadIdx.x, _once = 1; _once; _once = 0)
#define SCP_SYNC(c) scp_team_sync((c).team)
#define SCP_FMAX(a, b) fmax(a, b)
#define SCP_FMIN(a, b) fmin(a, b)
#endif

#ifndef SCP_EMU
#include <cooperative_groups.h>
__device__ __forceinline__ long long scp_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
// Barrier of the team that solves one scenario: the CTA (team of 1) or the whole cooperative grid.
__device__ __forceinline__ void scp_team_sync(int team) {
  if (team > 1) cooperative_groups::this_grid().sync(); else __syncthreads();
}
#endif

namespace scp {

constexpr int CH = 8;           // scan chunk length
constexpr int RED = 1024;       // reduction column stride for a one-CTA team (>= its threads)
// team scratch `sh`: 4 reduction columns of stride rs | 16 misc | 1056 second-level reduction
constexpr int SH_EXTRA = 1072;
#ifndef SCP_EMU
__host__ __device__
#endif
inline size_t sh_doubles(size_t rs) { return 4 * rs + SH_EXTRA; }

struct Tables {                 // constant per (K, h, weights); unit rho
  const double* B1;             // K*K  D'RjD + Ra + V'RvV + S'RpS
  const double* B2;             // K*K  S'RcS (one copy)
  const double* rj;             // K
  const double* ra;             // K
  const double* rv;             // K
  const double* rp;             // K
  const double* rc;             // K
};

// Offsets (in doubles / ints) of the per-slot scratch.
struct Layout {
  size_t Minv, N0, Qm, x, xprev, rhs, vj, va, vv, vp, posrow, velrow, P, Pbar, F, FY, off, deq, mu;
  size_t c_eta, c_bound, lam, scr, red, n_double;
  size_t xt, Pt, yj, ya, yv, yp, plam, pL, pG, prhs, py, pb_, pex, pey, pcv, pcp;   // polish (doubles)
  size_t pmark, pcmark, ptype, pq, pj2, pk, psgn, pdec, pcdec, ppos, pcpos, pcown, pid, chg, pslot, pdirty, pdl;   // polish (ints)
  size_t pscore, pcscore;                                               // polish (doubles)
  int pcap;
  size_t cnt, coff, c_j, flags, n_int;
  size_t cap;
};

inline
#ifndef SCP_EMU
__host__ __device__
#endif
Layout make_layout(int N, int K) {
  Layout L;
  size_t Q = 2 * (size_t)N, QK = Q * K, o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 1) & ~(size_t)1; return r; };
  L.Minv = take((size_t)K * K);
  L.N0 = take(2 * (size_t)K);
  L.Qm = take(2 * (size_t)K);
  L.x = take(QK); L.xprev = take(QK); L.rhs = take(QK);
  L.vj = take(QK); L.va = take(QK); L.vv = take(QK); L.vp = take(QK);
  L.posrow = take(QK); L.velrow = take(QK);
  L.P = take(QK); L.Pbar = take(QK); L.F = take(QK); L.FY = take(QK);
  L.off = take(QK); L.deq = take(2 * Q); L.mu = take(2 * Q);
  L.cap = (size_t)N * (size_t)K * (size_t)(N > 1 ? N - 1 : 1);
  L.c_eta = take(2 * L.cap); L.c_bound = take(L.cap); L.lam = take((size_t)N * N * K);
  L.scr = take(3 * Q * ((size_t)(K + CH - 1) / CH) + 2 * Q + 2100 + 2 * 1024);
  L.red = take(4 * RED);
  L.pcap = 12 * N + 128; if (L.pcap > 1024) L.pcap = 1024;
  L.xt = take(QK); L.Pt = take(QK); L.yj = take(QK); L.ya = take(QK); L.yv = take(QK); L.yp = take(QK);
  L.plam = take(L.cap);
  L.pL = take((size_t)L.pcap * L.pcap); L.pG = take((size_t)L.pcap * L.pcap); L.prhs = take(L.pcap); L.py = take(L.pcap); L.pb_ = take(L.pcap);
  L.pex = take(L.pcap); L.pey = take(L.pcap); L.pcv = take(L.pcap); L.pcp = take(L.pcap);
  L.pscore = take(4 * QK); L.pcscore = take(L.cap);
  L.n_double = o;
  size_t p = 0;
  auto takei = [&](size_t n) { size_t r = p; p += (n + 3) & ~(size_t)3; return r; };
  L.cnt = takei((size_t)N * K + 1); L.coff = takei((size_t)N * K + 1); L.c_j = takei(L.cap); L.flags = takei(((size_t)N * N * K + 3) / 4);
  L.pmark = takei(4 * QK); L.pcmark = takei(L.cap); L.pdec = takei(4 * QK); L.pcdec = takei(L.cap);
  L.ppos = takei(4 * QK); L.pcpos = takei(L.cap); L.pcown = takei(L.cap); L.pid = takei(L.pcap); L.chg = takei(L.pcap);
  L.pslot = takei(L.pcap); L.pdirty = takei(L.pcap); L.pdl = takei(L.pcap);
  L.ptype = takei(L.pcap); L.pq = takei(L.pcap); L.pj2 = takei(L.pcap); L.pk = takei(L.pcap); L.psgn = takei(L.pcap);
  L.n_int = p;
  return L;
}

struct Params {                 // one per launch (kernel parameter / constant bank)
  scp_b200_problem pb;
  Tables tb;
  Layout L;
};

struct Ctx {
  int nthreads;                 // threads of the team solving this scenario
  int team, tid0;               // CTAs in the team, first team-thread id of this CTA
  int np;                       // workers of the serial-prefix phases (<= 512)
  int rs;                       // stride of the reduction columns in `sh` (>= nthreads)
  double* sh;                   // scratch visible to the whole team: shared memory (team of 1) or global
  int N, K, Q;                  // Q = 2N agent-axes
  const Params* g;
  double* wd;                   // slot scratch (doubles)
  int* wi;                      // slot scratch (ints)
  double* sm;                   // shared scratch: [0,4*RED) reductions, then optional Nmat copy
  const double* nmat;           // where P2 reads the operator from (shared or global)
  int nmat_in_smem;
  double *a_x, *a_rhs, *a_vj, *a_va, *a_vv, *a_vp, *a_P, *a_F;   // hot per-agent-axis arrays: shared memory when they fit
  // per scenario
  const double *p0, *v0, *pf, *vf;
  double *acc, *pos, *vel;
  scp_b200_record* rec;
  int fused_epl;                // 0: phase-style iterations only; 2/4: warp-fused iteration with that many steps per lane
  double* fused_rows;           // scratch rows in shared memory (nwarps x K doubles >= 2K + 4N)
  int all_hot;                  // every hot array lives in shared memory
  int mma_ok;                   // the tensor-pipe iteration pays (all hot arrays in shared memory)
  // The polish factors its Gram matrix in shared memory: the hot arrays P, F, x, vp, vv, vj, va (everything but rhs)
  // are parked in their global homes for the duration of an attempt and the a_* pointers follow them.
  int pol_no_smem;              // a set outgrew the shared-memory region: later attempts of this scenario use the slot scratch
  double* pol_A;                // where Ainv of the running attempt lives (parked shared region or slot scratch)
  double* pol_smem;             // start of the parked region (null: nothing to park -- team kernel, test build)
  size_t pol_smem_doubles;
  double* hot_s[7];             // shared-memory copies (null when the array lives in global memory anyway)
  double* hot_g[7];             // global homes
  int polish_rounds;            // add/drop rounds over all attempts of this scenario
  int pol_valid, pol_n, pol_use_col, pol_col_stale;   // polish list/inverse state carried between attempts
  int pol_max_failed;           // failed polish attempts this subproblem may spend (halves with every unsolved subproblem of the scenario)
  long long t_pbuild, t_psolve, t_peval, t_papply;
  long long t_admm, t_polish;   // SM clock cycles spent in ADMM iterations / polish attempts
  long long t_fused, t_colx, t_chk;   // SCP_PROFILE_SPLIT builds only
  // block-uniform solver state
  double rho;
  int copies;
  int ncand;
};

// ------------------------------------------------------------------ reductions
// Phase-style block reductions over a value every thread contributes.
// kind 0: max, 1: sum.  Result returned to all threads.
#ifdef SCP_EMU
#define SCP_REDUCE_FN inline
#define SCP_PHASE_ARGS(nt, tid0_) for (int tid = 0; tid < (nt); ++tid)
#define SCP_SYNC_ARGS(team_) ((void)0)
#else
#define SCP_REDUCE_FN __device__ __noinline__
#define SCP_PHASE_ARGS(nt, tid0_) for (int tid = (tid0_) + threadIdx.x, _once = 1; _once; _once = 0)
#define SCP_SYNC_ARGS(team_) scp_team_sync(team_)
#endif
// A real (non-inlined) function: it has ~50 call sites and takes only scalars, so nothing of the caller's context is
// forced into memory.
template <int kind>
SCP_REDUCE_FN double reduce_finish_impl(double* sh, int rs, int nthreads, int team, int tid0, int slot) {
  // every team thread has written red[tid]; two strided levels (<= 1024, then 32), result to all threads
  double* red = sh + (size_t)slot * rs;
  double* aux = sh + (size_t)4 * rs + 16;          // 1024 + 32
  const int n1 = nthreads < 1024 ? nthreads : 1024;
  SCP_PHASE_ARGS(nthreads, tid0) {
    if (tid < n1) {
      double a = kind ? 0.0 : -INFINITY;
      for (int e = tid; e < nthreads; e += n1) a = kind ? a + red[e] : SCP_FMAX(a, red[e]);
      aux[tid] = a;
    }
  }
  SCP_SYNC_ARGS(team);
  SCP_PHASE_ARGS(nthreads, tid0) {
    if (tid < 32) {
      double a = kind ? 0.0 : -INFINITY;
      for (int e = tid; e < n1; e += 32) a = kind ? a + aux[e] : SCP_FMAX(a, aux[e]);
      aux[1024 + tid] = a;
    }
  }
  SCP_SYNC_ARGS(team);
  double r = kind ? 0.0 : -INFINITY;
  for (int e = 0; e < 32; ++e) r = kind ? r + aux[1024 + e] : SCP_FMAX(r, aux[1024 + e]);
  SCP_SYNC_ARGS(team);
  return r;
}
SCP_DEV double reduce_finish(Ctx& c, int slot, int kind) {
  return kind ? reduce_finish_impl<1>(c.sh, c.rs, c.nthreads, c.team, c.tid0, slot)
              : reduce_finish_impl<0>(c.sh, c.rs, c.nthreads, c.team, c.tid0, slot);
}

SCP_DEV double clampd(double v, double lo, double hi) { return SCP_FMIN(SCP_FMAX(v, lo), hi); }

// ------------------------------------------------------------------ scenario setup
// off[q][k] = p0 + h (k+1) v0 : constant part of position state k+1 (scp.py:246-247)
// deq[q] = (vf - v0, pf - off[K-1])  terminal equalities (scp.py:223-224, 256-257)
SCP_DEV void setup_scenario(Ctx& c) {
  const int K = c.K;
  double* off = c.wd + c.g->L.off;
  double* deq = c.wd + c.g->L.deq;
  double* x = c.a_x;
  double* F = c.a_F;
  double* FY = c.wd + c.g->L.FY;
  double* mu = c.wd + c.g->L.mu;
  const double h = c.g->pb.time_step;
  SCP_PHASE(c) {
    for (int e = tid; e < c.Q * K; e += c.nthreads) {
      int q = e / K, k = e - q * K;
      off[e] = c.p0[q] + h * (double)(k + 1) * c.v0[q];
      x[e] = 0.0;
      F[e] = 0.0; FY[e] = 0.0;          // scratch is not zero-initialised by the caller
    }
    for (int t = tid; t < 2 * c.Q; t += c.nthreads) mu[t] = 0.0;
    for (int t = tid; t <= c.N * K; t += c.nthreads) (c.wi + c.g->L.coff)[t] = 0;
    for (int q = tid; q < c.Q; q += c.nthreads) {
      deq[2 * q + 0] = c.vf[q] - c.v0[q];
      deq[2 * q + 1] = c.pf[q] - (c.p0[q] + h * (double)K * c.v0[q]);
    }
  }
  SCP_SYNC(c);
}

// ------------------------------------------------------------------ operator
// Minv <- inverse of M(rho, copies) by in-place Gauss-Jordan (SPD, no pivoting),
// then the equality-constrained solution operator
//   x = Nmat r + N0 d,   mu = Qm r - G d,   Nmat = Minv - Minv C'(C Minv C')^-1 C Minv
// C = [h 1' ; S[K-1,:]] (terminal velocity / position rows).
SCP_DEV void factor_operator(Ctx& c) {
  const int K = c.K;
  double* M = c.wd + c.g->L.Minv;
  double* N0 = c.wd + c.g->L.N0;
  double* Qm = c.wd + c.g->L.Qm;
  double* colb = c.sh;             // K  (K <= RED assumed for the scratch; checked on host)
  double* rowb = c.sh + c.rs;      // K
  double* mc = c.sh + 2 * (size_t)c.rs;   // 2K : Minv C'
  const double rho = c.rho, sig = c.g->pb.sigma, cp = (double)c.copies;
  const double h = c.g->pb.time_step;
  SCP_PHASE(c) {
    for (int e = tid; e < K * K; e += c.nthreads) {
      int r = e / K, cc = e - r * K;
      M[e] = rho * (c.g->tb.B1[e] + cp * c.g->tb.B2[e]) + (r == cc ? 2.0 + sig : 0.0);
    }
  }
  SCP_SYNC(c);
  for (int p = 0; p < K; ++p) {
    SCP_PHASE(c) {
      for (int e = tid; e < K; e += c.nthreads) { colb[e] = M[e * K + p]; rowb[e] = M[p * K + e]; }
    }
    SCP_SYNC(c);
    const double ip = 1.0 / colb[p];
    SCP_PHASE(c) {
      for (int e = tid; e < K * K; e += c.nthreads) {
        int r = e / K, cc = e - r * K;
        double v;
        if (r == p) v = (cc == p) ? ip : rowb[cc] * ip;
        else if (cc == p) v = -colb[r] * ip;
        else v = M[e] - colb[r] * rowb[cc] * ip;
        M[e] = v;
      }
    }
    SCP_SYNC(c);
  }
  // mc[k][e] = sum_j Minv[k][j] C[e][j]
  SCP_PHASE(c) {
    for (int k = tid; k < K; k += c.nthreads) {
      double a0 = 0.0, a1 = 0.0;
      for (int j = 0; j < K; ++j) {
        double m = M[k * K + j];
        a0 += m * h;
        a1 += m * (h * h * ((double)(K - 1 - j) + 0.5));
      }
      mc[2 * k] = a0; mc[2 * k + 1] = a1;
    }
  }
  SCP_SYNC(c);
  double h00 = 0, h01 = 0, h11 = 0;   // H = C mc (2x2, symmetric); every thread computes it
  for (int k = 0; k < K; ++k) {
    double cv = h, cpk = h * h * ((double)(K - 1 - k) + 0.5);
    h00 += cv * mc[2 * k]; h01 += cv * mc[2 * k + 1]; h11 += cpk * mc[2 * k + 1];
  }
  const double det = h00 * h11 - h01 * h01;
  const double g00 = h11 / det, g01 = -h01 / det, g11 = h00 / det;
  SCP_PHASE(c) {
    for (int k = tid; k < K; k += c.nthreads) {
      double n0 = mc[2 * k] * g00 + mc[2 * k + 1] * g01;
      double n1 = mc[2 * k] * g01 + mc[2 * k + 1] * g11;
      N0[2 * k] = n0; N0[2 * k + 1] = n1;          // N0 = mc G
      Qm[k] = n0; Qm[K + k] = n1;                  // Qm = G mc' (G symmetric)
    }
    if (tid == 0) { double* gg = c.sh + 4 * (size_t)c.rs; gg[0] = g00; gg[1] = g01; gg[2] = g11; }
  }
  SCP_SYNC(c);
  SCP_PHASE(c) {
    for (int e = tid; e < K * K; e += c.nthreads) {
      int r = e / K, cc = e - r * K;
      M[e] -= mc[2 * r] * Qm[cc] + mc[2 * r + 1] * Qm[K + cc];
    }
  }
  SCP_SYNC(c);
  if (c.nmat_in_smem) {
    double* dst = c.sm + sh_doubles(RED);
    SCP_PHASE(c) { for (int e = tid; e < K * K; e += c.nthreads) dst[e] = M[e]; }
    SCP_SYNC(c);
    c.nmat = dst;
  } else {
    c.nmat = M;
  }
}

// ------------------------------------------------------------------ forward rows
// rows of A x for every agent-axis: jerk (D x), vel (h cumsum), pos (h^2 (cumsum2 - cumsum/2)).
// mode 0: initialise v := A x (y = 0)            (OSQP warm start with x only, scp.py:443)
// mode 1: ADMM update v := A x + (v - clip(v))
// Also writes posrow/velrow and the positions P[q][k], k = 1..K-1 (P[q][0] = p0).
SCP_DEV void forward_rows(Ctx& c, int mode, int store_rows = 1) {
  const int K = c.K, nch = (K + CH - 1) / CH;
  const double h = c.g->pb.time_step, ih = 1.0 / h;
  double* x = c.a_x;
  double* s1 = c.wd + c.g->L.scr;            // chunk totals
  double* s2 = s1 + (size_t)c.Q * nch;
  double *vj = c.a_vj, *va = c.a_va, *vv = c.a_vv, *vp = c.a_vp;
  double *posrow = c.wd + c.g->L.posrow, *velrow = c.wd + c.g->L.velrow, *P = c.a_P, *off = c.wd + c.g->L.off;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double lo[2] = {c.g->pb.space[0], c.g->pb.space[1]}, hi[2] = {c.g->pb.space[2], c.g->pb.space[3]};
  SCP_PHASE(c) {
    for (int t = tid; t < c.Q * nch; t += c.nthreads) {
      int q = t / nch, ch = t - q * nch;
      int k0 = ch * CH, k1 = k0 + CH < K ? k0 + CH : K;
      double a1 = 0, a2 = 0;
      for (int k = k0; k < k1; ++k) { a1 += x[q * K + k]; a2 += a1; }
      s1[t] = a1; s2[t] = a2;
    }
  }
  SCP_SYNC(c);
  SCP_PHASE(c) {
    for (int t = tid; t < c.Q * nch; t += c.nthreads) {
      int q = t / nch, ch = t - q * nch;
      int k0 = ch * CH, k1 = k0 + CH < K ? k0 + CH : K;
      double c1 = 0, c2 = 0;
      for (int cc = 0; cc < ch; ++cc) { c2 += s2[q * nch + cc] + (double)CH * c1; c1 += s1[q * nch + cc]; }
      const double v0q = c.v0[q], p0q = c.p0[q];
      const double lv = -vl - v0q, uv = vl - v0q;
      const int ax = q & 1;
      for (int k = k0; k < k1; ++k) {
        const int e = q * K + k;
        const double xk = x[e];
        c1 += xk; c2 += c1;
        const double rv_ = h * c1, rp_ = h * h * (c2 - 0.5 * c1);
        const double offe = p0q + h * (double)(k + 1) * v0q;      // == off[e]
        if (store_rows) { velrow[e] = rv_; posrow[e] = rp_; }
        if (k + 1 < K) P[q * K + k + 1] = offe + rp_;
        if (mode == 0) {
          va[e] = xk;
          if (k < K - 1) { vj[e] = (x[e + 1] - xk) * ih; vv[e] = rv_; vp[e] = rp_; }
        } else {
          double v = va[e]; va[e] = xk + (v - clampd(v, -al, al));
          if (k < K - 1) {
            v = vj[e]; vj[e] = (x[e + 1] - xk) * ih + (v - clampd(v, -jl, jl));
            v = vv[e]; vv[e] = rv_ + (v - clampd(v, lv, uv));
            v = vp[e]; vp[e] = rp_ + (v - clampd(v, lo[ax] - offe, hi[ax] - offe));
          }
        }
      }
      if (ch == 0) P[q * K] = c.p0[q];
    }
  }
  SCP_SYNC(c);
}

// ------------------------------------------------------------------ transpose scans
// out[m] = base(m) + (D'wj)[m] + wa[m] + h sum_{k>=m} wv[k] + h^2 sum_{k>=m} (k-m+1/2) wp[k]
// mode 0 (rhs of the x-update):  w = rho r (2 clip(v) - v); wp += rho c rc posrow + F; base = sigma x
// mode 1 (dual residual):        w = rho r (v - clip(v)) = y; wp -= FY; base = 2 x + C'mu
//   in mode 1 the result is written to rhs and max|.| terms are left to the caller.
// mode 2 (polish):               w = y read from the dense arrays yj/ya/yv/yp, wp -= FY; base = 0
// mode 3 (infeasibility test):   w = dy = rho r (A x - clip(v)) (the last multiplier step), wp -= FD; base = 0
SCP_DEV void transpose_rows(Ctx& c, int mode) {
  const int K = c.K, nch = (K + CH - 1) / CH;
  const double h = c.g->pb.time_step, ih = 1.0 / h, rho = c.rho;
  double* x = c.a_x;
  double* out = c.a_rhs;
  double *vj = c.a_vj, *va = c.a_va, *vv = c.a_vv, *vp = c.a_vp;
  double *posrow = c.wd + c.g->L.posrow, *off = c.wd + c.g->L.off, *mu = c.wd + c.g->L.mu;
  double* Fm = (mode == 0) ? c.a_F : (mode == 3 ? c.wd + c.g->L.Pt : c.wd + c.g->L.FY);
  const double* velrow3 = c.wd + c.g->L.velrow;
  const double* posrow3 = c.wd + c.g->L.posrow;
  const double* Pcur = c.a_P;
  if (mode == 2) { vj = c.wd + c.g->L.yj; va = c.wd + c.g->L.ya; vv = c.wd + c.g->L.yv; vp = c.wd + c.g->L.yp; }
  double* t1 = c.wd + c.g->L.scr;          // chunk totals
  double* t2 = t1 + (size_t)c.Q * nch;
  double* t3 = t2 + (size_t)c.Q * nch;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double lo[2] = {c.g->pb.space[0], c.g->pb.space[1]}, hi[2] = {c.g->pb.space[2], c.g->pb.space[3]};
  const double cpr = (double)c.copies * rho;
  const double sig = c.g->pb.sigma;
  for (int pass = 0; pass < 2; ++pass) {
    SCP_PHASE(c) {
      for (int t = tid; t < c.Q * nch; t += c.nthreads) {
        int q = t / nch, ch = t - q * nch;
        int k0 = ch * CH, k1 = k0 + CH < K ? k0 + CH : K;
        const double v0q = c.v0[q], p0q = c.p0[q];
        const double lv = -vl - v0q, uv = vl - v0q;
        const int ax = q & 1;
        double r1v = 0, r1p = 0, r2p = 0;
        if (pass == 1) {
          for (int cc = nch - 1; cc > ch; --cc) {
            int len = (cc * CH + CH < K ? CH : K - cc * CH);
            r2p += t3[q * nch + cc] + (double)len * r1p;
            r1p += t2[q * nch + cc];
            r1v += t1[q * nch + cc];
          }
        }
        for (int k = k1 - 1; k >= k0; --k) {
          const int e = q * K + k;
          double wv = 0, wp = 0;
          if (k < K - 1) {
            double v = vv[e], z = clampd(v, lv, uv);
            wv = mode == 2 ? v : rho * c.g->tb.rv[k] * (mode == 0 ? 2 * z - v : (mode == 3 ? velrow3[e] - z : v - z));
            const double offe = p0q + h * (double)(k + 1) * v0q;   // == off[e]
            v = vp[e]; z = clampd(v, lo[ax] - offe, hi[ax] - offe);
            wp = mode == 2 ? v : rho * c.g->tb.rp[k] * (mode == 0 ? 2 * z - v : (mode == 3 ? posrow3[e] - z : v - z));
            // force on position state k+1 <-> row k; posrow[e] = P[k+1] - off[e]
            if (mode == 0) wp += cpr * c.g->tb.rc[k] * (Pcur[q * K + k + 1] - offe) + Fm[q * K + k + 1];
            else wp -= Fm[q * K + k + 1];
          }
          r1v += wv; r1p += wp; r2p += r1p;
          if (pass == 1) {
            double v = va[e], z = clampd(v, -al, al);
            double o = mode == 2 ? v : rho * c.g->tb.ra[k] * (mode == 0 ? 2 * z - v : (mode == 3 ? x[e] - z : v - z));
            double wj0 = 0, wj1 = 0;   // wj[k-1], wj[k]
            if (k >= 1) { v = vj[e - 1]; z = clampd(v, -jl, jl); wj0 = mode == 2 ? v : rho * c.g->tb.rj[k - 1] * (mode == 0 ? 2 * z - v : (mode == 3 ? (x[e] - x[e - 1]) * ih - z : v - z)); }
            if (k < K - 1) { v = vj[e]; z = clampd(v, -jl, jl); wj1 = mode == 2 ? v : rho * c.g->tb.rj[k] * (mode == 0 ? 2 * z - v : (mode == 3 ? (x[e + 1] - x[e]) * ih - z : v - z)); }
            o += (wj0 - wj1) * ih + h * r1v + h * h * (r2p - 0.5 * r1p);
            if (mode == 0) o += sig * x[e];
            else if (mode == 1) o += 2.0 * x[e] + h * mu[2 * q] + h * h * ((double)(K - 1 - k) + 0.5) * mu[2 * q + 1];
            out[e] = o;
          }
        }
        if (pass == 0) { t1[t] = r1v; t2[t] = r1p; t3[t] = r2p; }
      }
    }
    SCP_SYNC(c);
  }
}

// ------------------------------------------------------------------ x-update
SCP_DEV void x_update(Ctx& c, int want_mu) {
  const int K = c.K;
  double* x = c.a_x;
  const double* rhs = c.a_rhs;
  const double* N0 = c.wd + c.g->L.N0;
  const double* Qm = c.wd + c.g->L.Qm;
  const double* deq = c.wd + c.g->L.deq;
  double* mu = c.wd + c.g->L.mu;
  const double* Nm = c.nmat;
  SCP_PHASE(c) {
    for (int e = tid; e < c.Q * K; e += c.nthreads) {
      int q = e / K, k = e - q * K;
      const double* r = rhs + (size_t)q * K;
      double a = N0[2 * k] * deq[2 * q] + N0[2 * k + 1] * deq[2 * q + 1];
      for (int j = 0; j < K; ++j) a += Nm[j * K + k] * r[j];   // Nmat symmetric: column read is coalesced
      x[e] = a;
    }
    if (want_mu) {
      const double* gg = c.sh + 4 * (size_t)c.rs;
      const double g00 = gg[0], g01 = gg[1], g11 = gg[2];
      for (int t = tid; t < 2 * c.Q; t += c.nthreads) {
        int q = t >> 1, ee = t & 1;
        const double* r = rhs + (size_t)q * K;
        double a = 0;
        for (int j = 0; j < K; ++j) a += Qm[ee * K + j] * r[j];
        double gd = ee == 0 ? g00 * deq[2 * q] + g01 * deq[2 * q + 1] : g01 * deq[2 * q] + g11 * deq[2 * q + 1];
        mu[t] = a - gd;
      }
    }
  }
  SCP_SYNC(c);
}

#ifndef SCP_EMU
// ------------------------------------------------------------------ fused iteration (GPU only)
// One ADMM iteration for the agent-axis part with ONE warp per agent-axis q and NO block barrier
// inside: lanes hold EPL consecutive-by-32 steps (k = lane + 32 e), suffix/prefix sums run as warp
// shuffles, the K x K operator is applied from shared memory with the right-hand side broadcast from
// a per-warp shared row.  Same arithmetic as transpose_rows(0) + x_update + forward_rows(1); used for
// the iterations without a residual check when the hot arrays live in shared memory.
template <int EPL>
__device__ __forceinline__ void warp_suffix_sum(double (&v)[EPL], int lane) {
  // inclusive suffix sum over the sequence k = lane + 32 e (e major): out[k] = sum_{k' >= k} in[k']
  double carry = 0.0;
#pragma unroll
  for (int e = EPL - 1; e >= 0; --e) {
    double s = v[e];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double o = __shfl_down_sync(0xffffffffu, s, d);
      if (lane + d < 32) s += o;
    }
    s += carry;
    v[e] = s;
    carry = __shfl_sync(0xffffffffu, s, 0);
  }
}

template <int EPL>
__device__ __forceinline__ void warp_prefix_sum(double (&v)[EPL], int lane) {
  double carry = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    double s = v[e];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double o = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d) s += o;
    }
    s += carry;
    v[e] = s;
    carry = __shfl_sync(0xffffffffu, s, 31);
  }
}

template <int EPL>
__device__ __forceinline__ void admm_iter_fused(Ctx& c, double* rhs_rows /* nwarps x K in shared */) {
  const int K = c.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const double h = c.g->pb.time_step, ih = 1.0 / h, rho = c.rho, sig = c.g->pb.sigma;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double cpr = (double)c.copies * rho;
  const double* Nm = c.nmat;
  const double* N0 = c.wd + c.g->L.N0;
  const double* deq = c.wd + c.g->L.deq;
  double* myrhs = rhs_rows + warp * K;
  double trj[EPL], tra[EPL], trv[EPL], trp[EPL], trc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int k = lane + 32 * e;
    const bool in = k < K;
    trj[e] = in ? rho * c.g->tb.rj[k] : 0.0; tra[e] = in ? rho * c.g->tb.ra[k] : 0.0;
    trv[e] = in ? rho * c.g->tb.rv[k] : 0.0; trp[e] = in ? rho * c.g->tb.rp[k] : 0.0;
    trc[e] = in ? cpr * c.g->tb.rc[k] : 0.0;
  }
  for (int q = warp; q < c.Q; q += nwarps) {
    const double v0q = c.v0[q], p0q = c.p0[q];
    const double lv = -vl - v0q, uv = vl - v0q;
    const double plo = c.g->pb.space[q & 1], phi = c.g->pb.space[2 + (q & 1)];
    double *x = c.a_x + q * K, *vj = c.a_vj + q * K, *va = c.a_va + q * K, *vv = c.a_vv + q * K, *vp = c.a_vp + q * K;
    double *P = c.a_P + q * K, *F = c.a_F + q * K;
    double xo[EPL], sj[EPL], sa[EPL], sv[EPL], sp[EPL], wj[EPL], wv[EPL], wp[EPL], wa[EPL], off[EPL];
    // ---- rows -> weighted reflections
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int k = lane + 32 * e;
      xo[e] = 0; sj[e] = sa[e] = sv[e] = sp[e] = 0; wj[e] = wv[e] = wp[e] = wa[e] = 0; off[e] = 0;
      if (k < K) {
        xo[e] = x[k];
        double v = va[k], z = clampd(v, -al, al);
        sa[e] = v - z; wa[e] = tra[e] * (2 * z - v);
        if (k < K - 1) {
          v = vj[k]; z = clampd(v, -jl, jl); sj[e] = v - z; wj[e] = trj[e] * (2 * z - v);
          v = vv[k]; z = clampd(v, lv, uv); sv[e] = v - z; wv[e] = trv[e] * (2 * z - v);
          off[e] = p0q + h * (double)(k + 1) * v0q;
          v = vp[k]; z = clampd(v, plo - off[e], phi - off[e]); sp[e] = v - z;
          wp[e] = trp[e] * (2 * z - v) + trc[e] * (P[k + 1] - off[e]) + F[k + 1];
        }
      }
    }
    // ---- transpose: D'wj + wa + V'wv + S'wp
    double r1v[EPL], r1p[EPL], r2p[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) { r1v[e] = wv[e]; r1p[e] = wp[e]; }
    warp_suffix_sum<EPL>(r1v, lane);
    warp_suffix_sum<EPL>(r1p, lane);
#pragma unroll
    for (int e = 0; e < EPL; ++e) r2p[e] = r1p[e];
    warp_suffix_sum<EPL>(r2p, lane);
    double last = 0.0;   // wj[k-1] across the lane/element boundary
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int k = lane + 32 * e;
      double prev = __shfl_up_sync(0xffffffffu, wj[e], 1);
      if (lane == 0) prev = last;                    // element 32e - 1 lives in lane 31 of the previous element row
      last = __shfl_sync(0xffffffffu, wj[e], 31);
      if (k < K)
        myrhs[k] = sig * xo[e] + (prev - wj[e]) * ih + wa[e] + h * r1v[e] + h * h * (r2p[e] - 0.5 * r1p[e]);
    }
    __syncwarp();
    // ---- x = Nmat rhs + N0 d
    double xn[EPL];
    const double d0 = deq[2 * q], d1 = deq[2 * q + 1];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int k = lane + 32 * e;
      xn[e] = (k < K) ? N0[2 * k] * d0 + N0[2 * k + 1] * d1 : 0.0;
    }
    {
      // column walk of the symmetric operator: lane owns columns lane, lane+32 (clamped, masked at the end)
      const double* np0 = Nm + lane;
      const double* np1 = Nm + ((lane + 32 < K) ? lane + 32 : lane);
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
      int j = 0;
      for (; j + 1 < K; j += 2) {
        const double r0 = myrhs[j], r1 = myrhs[j + 1];
        a0 += np0[0] * r0; a1 += np1[0] * r0;
        b0 += np0[K] * r1; b1 += np1[K] * r1;
        np0 += 2 * K; np1 += 2 * K;
      }
      if (j < K) { const double r0 = myrhs[j]; a0 += np0[0] * r0; a1 += np1[0] * r0; }
      if (lane < K) xn[0] += a0 + b0;
      if (EPL > 1 && lane + 32 < K) xn[1] += a1 + b1;
    }
    __syncwarp();
    // ---- forward rows and v update
    double c1[EPL], c2[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) c1[e] = xn[e];
    warp_prefix_sum<EPL>(c1, lane);
#pragma unroll
    for (int e = 0; e < EPL; ++e) c2[e] = c1[e];
    warp_prefix_sum<EPL>(c2, lane);
    double first_next = 0.0;   // x[k+1] across the boundary: element row e+1, lane 0
#pragma unroll
    for (int e = EPL - 1; e >= 0; --e) {
      const int k = lane + 32 * e;
      double nxt = __shfl_down_sync(0xffffffffu, xn[e], 1);
      if (lane == 31) nxt = first_next;
      first_next = __shfl_sync(0xffffffffu, xn[e], 0);
      if (k < K) {
        x[k] = xn[e];
        va[k] = xn[e] + sa[e];
        if (k < K - 1) {
          const double rv_ = h * c1[e], rp_ = h * h * (c2[e] - 0.5 * c1[e]);
          vj[k] = (nxt - xn[e]) * ih + sj[e];
          vv[k] = rv_ + sv[e];
          vp[k] = rp_ + sp[e];
          P[k + 1] = off[e] + rp_;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ tensor-pipe iteration (GPU only)
// The same ADMM iteration as admm_iter_fused, organised as three block phases so that the ONE GEMM-shaped contraction
// of the path -- x = N rhs for all 2N agent-axes of the scenario, [K x K] . [K x 2N] --
// runs on the fp64 tensor pipe (mma.sync m8n8k4, SASS DMMA) instead of one broadcast-FMA matvec per warp:
//   A  one warp per agent-axis: rows -> weighted reflections -> transposed operators -> rhs[q][.]   (shared memory)
//   B  tiles of 8 steps x 8 agent-axes over the warps, four accumulator tiles in flight per warp: x = N rhs + N0 d
//   C  one warp per agent-axis: forward rows of the new x, v-updates, new positions
// Lane L owns the ADJACENT steps 2L, 2L+1 (K <= 64), so a scan over the steps is one local add plus ONE 5-stage warp
// scan; the double scans S' w and S x are written as two independent first-moment scans (sum w, sum (k+1) w), so the
// three suffix scans of phase A -- and the two prefix scans of phase C -- advance together instead of one after another.
// SH: every hot array, the operator and the rhs rows live in shared memory -- they are then indexed off the dynamic
// shared array so that the compiler emits LDS/STS with 32-bit addresses instead of generic 64-bit loads.
__device__ __forceinline__ void scp_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool SH>
struct HotView {     // one hot array: shared-memory offset (SH) or generic pointer
  double* p; int o;
  __device__ __forceinline__ HotView(double* base, const double* smem0) : p(base), o(SH ? (int)(base - smem0) : 0) {}
  __device__ __forceinline__ double ld(int i) const {
    extern __shared__ double smem[];
    return SH ? smem[o + i] : p[i];
  }
  __device__ __forceinline__ void st(int i, double v) const {
    extern __shared__ double smem[];
    if (SH) smem[o + i] = v; else p[i] = v;
  }
};

template <bool SH>
__device__ __forceinline__ void admm_iter_mma(Ctx& c) {
  extern __shared__ double smem[];
  const int K = c.K, Q = c.Q, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const double h = c.g->pb.time_step, ih = 1.0 / h, rho = c.rho, sig = c.g->pb.sigma;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double cpr = (double)c.copies * rho;
  const HotView<SH> X(c.a_x, smem), RHS(c.a_rhs, smem), VJ(c.a_vj, smem), VA(c.a_va, smem), VV(c.a_vv, smem),
      VP(c.a_vp, smem), PP(c.a_P, smem), FF(c.a_F, smem);
  const HotView<true> NM(const_cast<double*>(c.nmat), smem), SCR(c.fused_rows, smem);   // fused path: always shared
  const int k0 = 2 * lane, k1 = k0 + 1;
  const bool in0 = k0 < K, in1 = k1 < K, r0 = k0 < K - 1, r1 = k1 < K - 1;
  // ---- phase A
  {
    // N0 (2K) and d (2Q) for phase B's epilogue: staged in the scratch rows while the warps are busy with the scans
    const double* N0 = c.wd + c.g->L.N0;
    const double* deq = c.wd + c.g->L.deq;
    for (int e = threadIdx.x; e < 2 * K + 2 * Q; e += blockDim.x) SCR.st(e, e < 2 * K ? N0[e] : deq[e - 2 * K]);
    const double trj0 = r0 ? rho * c.g->tb.rj[k0] : 0.0, trj1 = r1 ? rho * c.g->tb.rj[k1] : 0.0;
    const double tra0 = in0 ? rho * c.g->tb.ra[k0] : 0.0, tra1 = in1 ? rho * c.g->tb.ra[k1] : 0.0;
    const double trv0 = r0 ? rho * c.g->tb.rv[k0] : 0.0, trv1 = r1 ? rho * c.g->tb.rv[k1] : 0.0;
    const double trp0 = r0 ? rho * c.g->tb.rp[k0] : 0.0, trp1 = r1 ? rho * c.g->tb.rp[k1] : 0.0;
    const double trc0 = r0 ? cpr * c.g->tb.rc[k0] : 0.0, trc1 = r1 ? cpr * c.g->tb.rc[k1] : 0.0;
    for (int q = warp; q < Q; q += nwarps) {
      const int b = q * K;
      const double v0q = c.v0[q], p0q = c.p0[q];
      const double lv = -vl - v0q, uv = vl - v0q;
      const double plo = c.g->pb.space[q & 1], phi = c.g->pb.space[2 + (q & 1)];
      double xo0 = 0, xo1 = 0, wa0 = 0, wa1 = 0, wj0 = 0, wj1 = 0, wv0 = 0, wv1 = 0, wp0 = 0, wp1 = 0;
      // every v array is overwritten with s = v - z (the scaled multiplier) here; phase C adds the new row to it
      if (in0) {
        xo0 = X.ld(b + k0);
        const double v = VA.ld(b + k0), z = clampd(v, -al, al);
        wa0 = tra0 * (2 * z - v); VA.st(b + k0, v - z);
      }
      if (in1) {
        xo1 = X.ld(b + k1);
        const double v = VA.ld(b + k1), z = clampd(v, -al, al);
        wa1 = tra1 * (2 * z - v); VA.st(b + k1, v - z);
      }
      if (r0) {
        double v = VJ.ld(b + k0), z = clampd(v, -jl, jl); wj0 = trj0 * (2 * z - v); VJ.st(b + k0, v - z);
        v = VV.ld(b + k0); z = clampd(v, lv, uv); wv0 = trv0 * (2 * z - v); VV.st(b + k0, v - z);
        const double off = p0q + h * (double)(k0 + 1) * v0q;
        v = VP.ld(b + k0); z = clampd(v, plo - off, phi - off); VP.st(b + k0, v - z);
        wp0 = trp0 * (2 * z - v) + trc0 * (PP.ld(b + k0 + 1) - off) + FF.ld(b + k0 + 1);
      }
      if (r1) {
        double v = VJ.ld(b + k1), z = clampd(v, -jl, jl); wj1 = trj1 * (2 * z - v); VJ.st(b + k1, v - z);
        v = VV.ld(b + k1); z = clampd(v, lv, uv); wv1 = trv1 * (2 * z - v); VV.st(b + k1, v - z);
        const double off = p0q + h * (double)(k1 + 1) * v0q;
        v = VP.ld(b + k1); z = clampd(v, plo - off, phi - off); VP.st(b + k1, v - z);
        wp1 = trp1 * (2 * z - v) + trc1 * (PP.ld(b + k1 + 1) - off) + FF.ld(b + k1 + 1);
      }
      // transposed operators: D'wj + wa + V'wv + S'wp; suffix sums of wv, wp and (k+1) wp over the steps
      double sv = wv0 + wv1, s0 = wp0 + wp1, s1 = (double)(k0 + 1) * wp0 + (double)(k1 + 1) * wp1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double a = __shfl_down_sync(0xffffffffu, sv, d), b0 = __shfl_down_sync(0xffffffffu, s0, d),
                     b1 = __shfl_down_sync(0xffffffffu, s1, d);
        if (lane + d < 32) { sv += a; s0 += b0; s1 += b1; }
      }
      double pj = __shfl_up_sync(0xffffffffu, wj1, 1);      // wj[k0 - 1]
      if (lane == 0) pj = 0.0;
      const double sv1 = sv - wv0, s01 = s0 - wp0, s11 = s1 - (double)(k0 + 1) * wp0;
      if (in0) RHS.st(b + k0, sig * xo0 + (pj - wj0) * ih + wa0 + h * sv + h * h * (s1 - ((double)k0 + 0.5) * s0));
      if (in1) RHS.st(b + k1, sig * xo1 + (wj0 - wj1) * ih + wa1 + h * sv1 + h * h * (s11 - ((double)k1 + 0.5) * s01));
    }
  }
  __syncthreads();
  // ---- phase B: x[q][k] = sum_j N[k][j] rhs[q][j] + N0[k] . d[q] on the fp64 tensor pipe
  {
    const int MT = (K + 7) >> 3, NT = (Q + 7) >> 3, total = MT * NT;
    const int g = lane >> 2, t = lane & 3;
    for (int tg = warp; tg < total; tg += 4 * nwarps) {
      int arow[4], brow[4];
      double c0[4], c1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tile = tg + i * nwarps;
        const int mt = tile % MT, nt = tile / MT;
        const int m = mt * 8 + g, n = nt * 8 + g;
        arow[i] = (tile < total && m < K ? m : K - 1) * K;
        brow[i] = (tile < total && n < Q ? n : Q - 1) * K;
        c0[i] = 0.0; c1[i] = 0.0;
      }
      // operands of step kk + 4 are fetched before the products of step kk issue
      double a[4], bb[4];
      {
        const int kc = t < K ? t : K - 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = NM.ld(arow[i] + kc); bb[i] = RHS.ld(brow[i] + kc); }
      }
      for (int kk = 0; kk < K; kk += 4) {
        const bool in = kk + t < K;
        const int kn = kk + 4 + t, kc = kn < K ? kn : K - 1;
        double a2[4], b2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a2[i] = NM.ld(arow[i] + kc); b2[i] = RHS.ld(brow[i] + kc); }
#pragma unroll
        for (int i = 0; i < 4; ++i) scp_dmma(c0[i], c1[i], in ? a[i] : 0.0, in ? bb[i] : 0.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = a2[i]; bb[i] = b2[i]; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tile = tg + i * nwarps;
        if (tile < total) {
          const int mt = tile % MT, nt = tile / MT;
          const int k = mt * 8 + g, q = nt * 8 + 2 * t;
          if (k < K) {
            const double n0a = SCR.ld(2 * k), n0b = SCR.ld(2 * k + 1);
            if (q < Q) X.st(q * K + k, c0[i] + n0a * SCR.ld(2 * K + 2 * q) + n0b * SCR.ld(2 * K + 2 * q + 1));
            if (q + 1 < Q) X.st((q + 1) * K + k, c1[i] + n0a * SCR.ld(2 * K + 2 * q + 2) + n0b * SCR.ld(2 * K + 2 * q + 3));
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- phase C: forward rows of the new x, v <- A x + (v - z), new positions
  for (int q = warp; q < Q; q += nwarps) {
    const int b = q * K;
    const double v0q = c.v0[q], p0q = c.p0[q];
    const double xn0 = in0 ? X.ld(b + k0) : 0.0, xn1 = in1 ? X.ld(b + k1) : 0.0;
    double t0 = xn0 + xn1, t1 = (double)k0 * xn0 + (double)k1 * xn1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double a0 = __shfl_up_sync(0xffffffffu, t0, d), a1 = __shfl_up_sync(0xffffffffu, t1, d);
      if (lane >= d) { t0 += a0; t1 += a1; }
    }
    const double nx = __shfl_down_sync(0xffffffffu, xn0, 1);     // x[k1 + 1]
    // inclusive prefix sums at k1 are (t0, t1); at k0 remove the k1 term
    const double c10 = t0 - xn1, c11 = t0;
    const double c20 = (double)(k0 + 1) * c10 - (t1 - (double)k1 * xn1), c21 = (double)(k1 + 1) * c11 - t1;
    if (in0) {
      VA.st(b + k0, xn0 + VA.ld(b + k0));
      if (r0) {
        const double rv_ = h * c10, rp_ = h * h * (c20 - 0.5 * c10), off = p0q + h * (double)(k0 + 1) * v0q;
        VJ.st(b + k0, (xn1 - xn0) * ih + VJ.ld(b + k0));
        VV.st(b + k0, rv_ + VV.ld(b + k0));
        VP.st(b + k0, rp_ + VP.ld(b + k0));
        PP.st(b + k0 + 1, off + rp_);
      }
    }
    if (in1) {
      VA.st(b + k1, xn1 + VA.ld(b + k1));
      if (r1) {
        const double rv_ = h * c11, rp_ = h * h * (c21 - 0.5 * c11), off = p0q + h * (double)(k1 + 1) * v0q;
        VJ.st(b + k1, (nx - xn1) * ih + VJ.ld(b + k1));
        VV.st(b + k1, rv_ + VV.ld(b + k1));
        VP.st(b + k1, rp_ + VP.ld(b + k1));
        PP.st(b + k1 + 1, off + rp_);
      }
    }
  }
}
#endif

// ------------------------------------------------------------------ collision rows
// One thread per (k, i), k = 1..K-1: walks its candidate rows, updates the row
// multipliers (both owners of a row update their own copy identically) and sums
// the force on p_i[k].  Returns (through red slot 0) max |dlam|/rho_c = primal
// residual of the copy rows when `want_res`.
SCP_DEV void collision_rows(Ctx& c, int want_res) {
  const int K = c.K, N = c.N;
  const double* P = c.a_P;
  double* F = c.a_F;
  double* FY = c.wd + c.g->L.FY;
  const int* coff = c.wi + c.g->L.coff;
  const int* cj = c.wi + c.g->L.c_j;
  const double* ceta = c.wd + c.g->L.c_eta;
  const double* cb = c.wd + c.g->L.c_bound;
  double* lam = c.wd + c.g->L.lam;        // dense [k][i][j]; the two owners keep identical copies
  double* dlam = c.wd + c.g->L.plam;      // last multiplier step per carried entry (check iterations)
  double* FD = c.wd + c.g->L.Pt;          // sum_j dlam eta (check iterations; Pt is free outside the polish)
  double* red = c.sh;
  const double rho = c.rho;
  SCP_PHASE(c) {
    double worst = 0.0;
    for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
      int k = 1 + t / N, i = t - (k - 1) * N;
      const double pix = P[(2 * i) * K + k], piy = P[(2 * i + 1) * K + k];
      const double rc = rho * c.g->tb.rc[k - 1];
      double fx = 0, fy = 0, yx = 0, yy = 0, dx_ = 0, dy_ = 0;
      for (int s = coff[k * N + i]; s < coff[k * N + i + 1]; ++s) {
        const int j = cj[s];
        const double ex = ceta[2 * s], ey = ceta[2 * s + 1];
        const double g = ex * (pix - P[(2 * j) * K + k]) + ey * (piy - P[(2 * j + 1) * K + k]);
        const size_t li = ((size_t)k * N + i) * N + j;
        const double l0 = lam[li];
        const double l1 = SCP_FMAX(0.0, l0 + 0.5 * rc * (cb[s] - g));
        lam[li] = l1;
        const double f = 2.0 * l1 - l0;
        fx += f * ex; fy += f * ey;
        yx += l1 * ex; yy += l1 * ey;
        worst = SCP_FMAX(worst, fabs(l1 - l0) / rc);
        if (want_res) { dlam[s] = l1 - l0; dx_ += (l1 - l0) * ex; dy_ += (l1 - l0) * ey; }
      }
      F[(2 * i) * K + k] = fx; F[(2 * i + 1) * K + k] = fy;
      if (want_res) {
        FY[(2 * i) * K + k] = yx; FY[(2 * i + 1) * K + k] = yy;
        FD[(2 * i) * K + k] = dx_; FD[(2 * i + 1) * K + k] = dy_;
      }
    }
    if (want_res) red[tid] = worst;
  }
  SCP_SYNC(c);
}

#ifndef SCP_EMU
// The collision rows of an iteration WITHOUT a residual check: the same update as collision_rows(c, 0) with the loads of
// up to four candidate rows of a (step, agent) issued together (they come from L2: partner index, normal, bound, the
// multiplier) before the first one is used, positions / forces through shared memory when they live there, and nothing
// computed that only the residual check needs.
template <bool SH>
__device__ __forceinline__ void collision_rows_fast(Ctx& c) {
  extern __shared__ double smem[];
  const int K = c.K, N = c.N;
  const HotView<SH> PP(c.a_P, smem), FF(c.a_F, smem);
  const int* __restrict__ coff = c.wi + c.g->L.coff;
  const int* __restrict__ cj = c.wi + c.g->L.c_j;
  const double2* __restrict__ ceta = reinterpret_cast<const double2*>(c.wd + c.g->L.c_eta);
  const double* __restrict__ cb = c.wd + c.g->L.c_bound;
  double* lam = c.wd + c.g->L.lam;
  const double hrho = 0.5 * c.rho;
  const int items = (K - 1) * N, nt = blockDim.x;
  const int dk = nt / N, di = nt - dk * N;
  int k = 1 + (int)threadIdx.x / N, i = (int)threadIdx.x - (k - 1) * N;
  for (int tI = threadIdx.x; tI < items; tI += nt) {
    const int s0 = coff[k * N + i], s1 = coff[k * N + i + 1];
    const double pix = PP.ld((2 * i) * K + k), piy = PP.ld((2 * i + 1) * K + k);
    const double hrc = hrho * c.g->tb.rc[k - 1];
    double* lrow = lam + ((size_t)k * N + i) * N;
    double fx = 0.0, fy = 0.0;
    for (int s = s0; s < s1; s += 4) {
      int j[4]; double2 e[4]; double bd[4], l0[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int su = s + u < s1 ? s + u : s;
        j[u] = cj[su]; e[u] = ceta[su]; bd[u] = cb[su];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) l0[u] = lrow[j[u]];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (s + u < s1) {
          const double g = e[u].x * (pix - PP.ld((2 * j[u]) * K + k)) + e[u].y * (piy - PP.ld((2 * j[u] + 1) * K + k));
          const double l1 = SCP_FMAX(0.0, l0[u] + hrc * (bd[u] - g));
          lrow[j[u]] = l1;
          const double f = 2.0 * l1 - l0[u];
          fx += f * e[u].x; fy += f * e[u].y;
        }
      }
    }
    FF.st((2 * i) * K + k, fx); FF.st((2 * i + 1) * K + k, fy);
    k += dk; i += di;
    if (i >= N) { i -= N; ++k; }
  }
  __syncthreads();
}
#endif

// Candidate rows.  flags[k][i][j] != 0 marks row (k,i,j) as carried by the ADMM.
// mark_near_rows: start of a subproblem -- rows whose linearisation-point distance
//   is below R + margin (and their multipliers zeroed: OSQP starts every
//   subproblem from y = 0, scp.py:441-443).
// build_candidates: CSR over the flagged rows, per owner (k,i): partner j, eta
//   oriented towards the owner (scp.py:498-509), bound = R + (eta.d - dist)
//   (scp.py:547-549).  Degenerate pairs (dist < 1e-6) get the deterministic
//   direction (+-1, 0) and dist := 1 (the reference draws a random one, :503-507).
// verify_rows: every row of the full QP evaluated at the current positions; a
//   violated row that is not carried is flagged (multiplier 0) and counted.
SCP_DEV void mark_near_rows(Ctx& c, double margin, int keep_lam) {
  const int K = c.K, N = c.N;
  const double* Pb = c.wd + c.g->L.Pbar;
  unsigned char* flags = (unsigned char*)(c.wi + c.g->L.flags);
  double* lam = c.wd + c.g->L.lam;
  double* F = c.a_F;
  double* FY = c.wd + c.g->L.FY;
  const double R = c.g->pb.min_distance;
  const double r2 = (R + margin) * (R + margin);
  SCP_PHASE(c) {
    for (int t = tid; t < K * N; t += c.nthreads) {
      int k = t / N, i = t - k * N;
      const double pix = Pb[(2 * i) * K + k], piy = Pb[(2 * i + 1) * K + k];
      for (int j = 0; j < N; ++j) {
        double dx = pix - Pb[(2 * j) * K + k], dy = piy - Pb[(2 * j + 1) * K + k];
        int near = (k >= 1) && (j != i) && (dx * dx + dy * dy < r2);
        // warm duals: a row carried by the previous subproblem keeps its multiplier (same minimiser, fewer iterations)
        if (near && !(keep_lam && flags[(size_t)t * N + j])) lam[(size_t)t * N + j] = 0.0;
        flags[(size_t)t * N + j] = (unsigned char)near;
      }
    }
    for (int e = tid; e < c.Q * K; e += c.nthreads) { F[e] = 0.0; FY[e] = 0.0; }
  }
  SCP_SYNC(c);
}

SCP_DEV void build_candidates(Ctx& c) {
  const int K = c.K, N = c.N;
  const double* Pb = c.wd + c.g->L.Pbar;
  const unsigned char* flags = (const unsigned char*)(c.wi + c.g->L.flags);
  int* cnt = c.wi + c.g->L.cnt;
  int* coff = c.wi + c.g->L.coff;
  int* cj = c.wi + c.g->L.c_j;
  double* ceta = c.wd + c.g->L.c_eta;
  double* cb = c.wd + c.g->L.c_bound;
  const double R = c.g->pb.min_distance;
  const int total = K * N;
  SCP_PHASE(c) {
    for (int t = tid; t < total; t += c.nthreads) {
      int n = 0;
      for (int j = 0; j < N; ++j) n += flags[(size_t)t * N + j];
      cnt[t] = n;
    }
  }
  SCP_SYNC(c);
  // exclusive scan of cnt -> coff (two-level, phase style), max -> copies
  const int np = c.np;
  const int per = (total + np - 1) / np;
  int* part = (int*)(c.sh + 2 * (size_t)c.rs);        // np ints
  int* pmax = part + 512;
  SCP_PHASE(c) {
    if (tid < np) {
      int s = 0, m = 0;
      for (int e = tid * per; e < total && e < (tid + 1) * per; ++e) { s += cnt[e]; m = cnt[e] > m ? cnt[e] : m; }
      part[tid] = s; pmax[tid] = m;
    }
  }
  SCP_SYNC(c);
  SCP_PHASE(c) {
    if (tid < np) {
      int base = 0;
      for (int e = 0; e < tid; ++e) base += part[e];
      for (int e = tid * per; e < total && e < (tid + 1) * per; ++e) { coff[e] = base; base += cnt[e]; }
      if (tid == np - 1) coff[total] = base;
    }
  }
  SCP_SYNC(c);
  int mx = 0;
  for (int e = 0; e < np; ++e) mx = pmax[e] > mx ? pmax[e] : mx;
  c.copies = mx;
  c.ncand = coff[total];
  c.pol_col_stale = 1;             // the candidate set changed: collision rows of the polish list refer to the old CSR
  SCP_SYNC(c);
  SCP_PHASE(c) {
    for (int t = tid; t < total; t += c.nthreads) {
      int k = t / N, i = t - k * N;
      if (cnt[t] == 0) continue;
      int s = coff[t];
      const double pix = Pb[(2 * i) * K + k], piy = Pb[(2 * i + 1) * K + k];
      for (int j = 0; j < N; ++j) {
        if (!flags[(size_t)t * N + j]) continue;
        double dx = pix - Pb[(2 * j) * K + k], dy = piy - Pb[(2 * j + 1) * K + k];
        double dist = hypot(dx, dy), ex, ey;
        if (dist < 1e-6) { ex = i < j ? 1.0 : -1.0; ey = 0.0; cb[s] = R + (ex * dx + ey * dy - 1.0); }
        else { ex = dx / dist; ey = dy / dist; cb[s] = R + ((ex * dx + ey * dy) - dist); }
        cj[s] = j; ceta[2 * s] = ex; ceta[2 * s + 1] = ey;
        ++s;
      }
    }
  }
  SCP_SYNC(c);
}

SCP_DEV int verify_rows(Ctx& c, double tol) {
  const int K = c.K, N = c.N;
  const double* Pb = c.wd + c.g->L.Pbar;
  const double* P = c.a_P;
  unsigned char* flags = (unsigned char*)(c.wi + c.g->L.flags);
  double* lam = c.wd + c.g->L.lam;
  const double R = c.g->pb.min_distance;
  double* red = c.sh;
  SCP_PHASE(c) {
    double bad = 0.0;
    for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
      int k = 1 + t / N, i = t - (k - 1) * N;
      const double bx = Pb[(2 * i) * K + k], by = Pb[(2 * i + 1) * K + k];
      const double px = P[(2 * i) * K + k], py = P[(2 * i + 1) * K + k];
      for (int j = i + 1; j < N; ++j) {
        const size_t fij = ((size_t)k * N + i) * N + j, fji = ((size_t)k * N + j) * N + i;
        if (flags[fij]) continue;
        double dx = bx - Pb[(2 * j) * K + k], dy = by - Pb[(2 * j + 1) * K + k];
        double dist = hypot(dx, dy), ex, ey, bound;
        if (dist < 1e-6) { ex = 1.0; ey = 0.0; bound = R + (dx - 1.0); }
        else { ex = dx / dist; ey = dy / dist; bound = R + ((ex * dx + ey * dy) - dist); }
        double g = ex * (px - P[(2 * j) * K + k]) + ey * (py - P[(2 * j + 1) * K + k]);
        if (g < bound - tol) {
          flags[fij] = 1; flags[fji] = 1; lam[fij] = 0.0; lam[fji] = 0.0;
          bad += 1.0;
        }
      }
    }
    red[tid] = bad;
  }
  SCP_SYNC(c);
  return (int)reduce_finish(c, 0, 1);
}

// _fast_check_avoidance_constraints (scp.py:597-615) on positions P (k = 0..K-1):
// min separation and the first row in scan order (k-major, i<j) below R - margin.
SCP_DEV void gate_and_minsep(Ctx& c, double* minsep, long long* first_row, double* first_dist) {
  const int K = c.K, N = c.N;
  const double* P = c.a_P;
  const double thr = c.g->pb.min_distance - c.g->pb.feas_margin;
  double* red = c.sh;            // slot 0: -min distance ; slot 1: -(first row index) ; slot 2: dist of it
  const long long npairs = (long long)N * (N - 1) / 2;
  SCP_PHASE(c) {
    double mn = INFINITY, fr = INFINITY, fd = 0.0;
    for (int t = tid; t < K * N; t += c.nthreads) {
      int k = t / N, i = t - k * N;
      const double px = P[(2 * i) * K + k], py = P[(2 * i + 1) * K + k];
      for (int j = i + 1; j < N; ++j) {
        double dx = px - P[(2 * j) * K + k], dy = py - P[(2 * j + 1) * K + k];
        // np.linalg.norm of a 2-vector: sqrt(dx^2 + dy^2)
        double d = sqrt(dx * dx + dy * dy);
        mn = SCP_FMIN(mn, d);
        if (d < thr) {
          double row = (double)((long long)k * npairs + ((long long)i * (2 * N - i - 1)) / 2 + (j - i - 1));
          if (row < fr) { fr = row; fd = d; }
        }
      }
    }
    red[tid] = -mn; red[c.rs + tid] = -fr; red[2 * (size_t)c.rs + tid] = fd;
  }
  SCP_SYNC(c);
  // first row: max of -fr, then the distance carried by the thread(s) that hold it
  const double best = reduce_finish(c, 1, 0);
  SCP_PHASE(c) { if (!(red[c.rs + tid] == best)) red[2 * (size_t)c.rs + tid] = -INFINITY; }
  SCP_SYNC(c);
  const double bd = (best == -INFINITY) ? 0.0 : reduce_finish(c, 2, 0);
  double m = reduce_finish(c, 0, 0);
  *minsep = -m;
  *first_row = (best == -INFINITY) ? -1 : (long long)(-best);
  *first_dist = bd;
}

// ------------------------------------------------------------------ polish
// Active-set refinement ("polish", OSQP paper section 4, with add/drop rounds).  Given a
// guess W of the active rows it solves  min sum||x||^2  s.t.  C x_q = d_q, u_r.x = b_r (r in W)
// exactly:  x = x_d - 1/2 Pi A_W' y,  (A_W Pi A_W') y = 2 (A_W x_d - b_W),
// Pi = I - C'(CC')^-1 C, x_d = C'(CC')^-1 d.  All Gram entries are closed forms of the row
// type and step (rows are D_k, e_k, V_k, S_k or eta-weighted S_k pairs), so G is assembled
// without touching a K-vector.  Then every carried row is checked (violated -> add, wrong
// multiplier sign -> drop) until W is stable: the result satisfies the KKT conditions of the
// carried QP to rounding, i.e. it IS its minimiser.
struct PRow { int type, q, j, k; double ex, ey, cv, cp; };   // cv,cp = C u for the row's library vector   // type 0 jerk,1 acc,2 vel,3 pos,4 collision(i=q,j, state k)

SCP_DEV double lib_dot(int t1, int k1, int t2, int k2, double h) {
  if (t1 > t2) { int t = t1; t1 = t2; t2 = t; t = k1; k1 = k2; k2 = t; }
  const double h2 = h * h;
  if (t1 == 0) {
    if (t2 == 0) return ((k1 == k2) ? 2.0 : ((k1 == k2 + 1 || k1 + 1 == k2) ? -1.0 : 0.0)) / h2;
    if (t2 == 1) return ((k2 == k1 + 1) ? 1.0 : ((k2 == k1) ? -1.0 : 0.0)) / h;
    if (t2 == 2) return (k1 == k2) ? -1.0 : 0.0;
    return (k1 + 1 <= k2) ? -h : ((k1 == k2) ? -0.5 * h : 0.0);
  }
  if (t1 == 1) {
    if (t2 == 1) return (k1 == k2) ? 1.0 : 0.0;
    if (t2 == 2) return (k1 <= k2) ? h : 0.0;
    return (k1 <= k2) ? h2 * ((double)(k2 - k1) + 0.5) : 0.0;
  }
  const double n = (double)((k1 < k2 ? k1 : k2) + 1);
  if (t1 == 2) {
    if (t2 == 2) return h2 * n;
    return h2 * h * (n * ((double)k2 + 0.5) - 0.5 * n * (n - 1.0));
  }
  const double a = (double)k1 + 0.5, b = (double)k2 + 0.5;
  return h2 * h2 * (n * a * b - (a + b) * 0.5 * n * (n - 1.0) + (n - 1.0) * n * (2.0 * n - 1.0) / 6.0);
}

struct PGeom { double h; int K; double i00, i01, i11; };   // (CC')^-1

SCP_DEV PGeom make_pgeom(double h, int K) {
  PGeom g; g.h = h; g.K = K;
  double h00 = lib_dot(2, K - 1, 2, K - 1, h), h01 = lib_dot(2, K - 1, 3, K - 1, h), h11 = lib_dot(3, K - 1, 3, K - 1, h);
  double det = h00 * h11 - h01 * h01;
  g.i00 = h11 / det; g.i01 = -h01 / det; g.i11 = h00 / det;
  return g;
}

// u' Pi u'  for library vectors (type 4 uses S_{k-1})
SCP_DEV double lib_proj(const PGeom& g, int t1, int k1, int t2, int k2) {
  const double c1v = lib_dot(t1, k1, 2, g.K - 1, g.h), c1p = lib_dot(t1, k1, 3, g.K - 1, g.h);
  const double c2v = lib_dot(t2, k2, 2, g.K - 1, g.h), c2p = lib_dot(t2, k2, 3, g.K - 1, g.h);
  return lib_dot(t1, k1, t2, k2, g.h) - (c1v * (g.i00 * c2v + g.i01 * c2p) + c1p * (g.i01 * c2v + g.i11 * c2p));
}

SCP_DEV double lib_projc(const PGeom& g, const PRow& a, int ta, int ka, const PRow& b, int tb, int kb) {
  return lib_dot(ta, ka, tb, kb, g.h) - (a.cv * (g.i00 * b.cv + g.i01 * b.cp) + a.cp * (g.i01 * b.cv + g.i11 * b.cp));
}

SCP_DEV double gram_entry(const PGeom& g, const PRow& a, const PRow& b) {
  if (a.type < 4 && b.type < 4) return (a.q == b.q) ? lib_projc(g, a, a.type, a.k, b, b.type, b.k) : 0.0;
  if (a.type == 4 && b.type == 4) {
    int sgn = (a.q == b.q) + (a.j == b.j) - (a.q == b.j) - (a.j == b.q);
    if (sgn == 0) return 0.0;
    return (double)sgn * (a.ex * b.ex + a.ey * b.ey) * lib_projc(g, a, 3, a.k - 1, b, 3, b.k - 1);
  }
  const PRow& d = a.type < 4 ? a : b;
  const PRow& cr = a.type < 4 ? b : a;
  const int ag = d.q >> 1;
  double coef = (ag == cr.q) ? 1.0 : ((ag == cr.j) ? -1.0 : 0.0);
  if (coef == 0.0) return 0.0;
  coef *= (d.q & 1) ? cr.ey : cr.ex;
  return coef * lib_projc(g, d, d.type, d.k, cr, 3, cr.k - 1);
}

SCP_DEV PRow load_prow(Ctx& c, int r) {
  PRow w;
  w.type = (c.wi + c.g->L.ptype)[r]; w.q = (c.wi + c.g->L.pq)[r]; w.j = (c.wi + c.g->L.pj2)[r];
  w.k = (c.wi + c.g->L.pk)[r]; w.ex = (c.wd + c.g->L.pex)[r]; w.ey = (c.wd + c.g->L.pey)[r];
  w.cv = (c.wd + c.g->L.pcv)[r]; w.cp = (c.wd + c.g->L.pcp)[r];
  return w;
}

// marks: dyn rows pmark[cls*QK + e] in {-1 lower, 0, +1 upper}; collision entries pcmark[s] in {0,1}
// (both owners' entries carry the mark; the i<j entry is the row's identity).
//
// The active rows live in a list (ptype/pq/pj2/pk/psgn/pb/pex/pey/pcv/pcp, n entries) together with
//   G0   = A_W Pi A_W'            (n x n, full symmetric storage, column stride ld)
//   Ainv = (G0 + delta I)^-1      (same storage)
// and the list position of every marked row (ppos / pcpos).  Rows enter and leave one at a time with
// O(n^2) bordered-inverse updates (add: Schur complement of the new row; drop: rank-one downdate, then
// the last row moves into the hole), so a round that changes c rows costs c n^2 instead of n^3.

// list slot `slot` <- box row (cls, e) active at its upper (m>0) / lower (m<0) bound
SCP_DEV void polish_fill_dyn(Ctx& c, int slot, int cls, int e, int m) {
  const int K = c.K, q = e / K, k = e - q * K;
  const double h = c.g->pb.time_step;
  (c.wi + c.g->L.ptype)[slot] = cls; (c.wi + c.g->L.pq)[slot] = q; (c.wi + c.g->L.pj2)[slot] = -1;
  (c.wi + c.g->L.pk)[slot] = k; (c.wi + c.g->L.psgn)[slot] = m;
  (c.wd + c.g->L.pex)[slot] = 0.0; (c.wd + c.g->L.pey)[slot] = 0.0;
  double b;
  if (cls == 0) b = m > 0 ? c.g->pb.jerk_limit : -c.g->pb.jerk_limit;
  else if (cls == 1) b = m > 0 ? c.g->pb.acc_limit : -c.g->pb.acc_limit;
  else if (cls == 2) b = (m > 0 ? c.g->pb.vel_limit : -c.g->pb.vel_limit) - c.v0[q];
  else b = (m > 0 ? c.g->pb.space[2 + (q & 1)] : c.g->pb.space[q & 1]) - (c.wd + c.g->L.off)[e];
  (c.wd + c.g->L.pb_)[slot] = b;
  (c.wd + c.g->L.pcv)[slot] = lib_dot(cls, k, 2, K - 1, h);
  (c.wd + c.g->L.pcp)[slot] = lib_dot(cls, k, 3, K - 1, h);
}

// list slot `slot` <- collision row of CSR entry sidx (owner i < partner j, state k)
SCP_DEV void polish_fill_col(Ctx& c, int slot, int sidx, int k, int i) {
  const int K = c.K;
  const double h = c.g->pb.time_step;
  const int j = (c.wi + c.g->L.c_j)[sidx];
  const double ex = (c.wd + c.g->L.c_eta)[2 * sidx], ey = (c.wd + c.g->L.c_eta)[2 * sidx + 1];
  const double* off = c.wd + c.g->L.off;
  (c.wi + c.g->L.ptype)[slot] = 4; (c.wi + c.g->L.pq)[slot] = i; (c.wi + c.g->L.pj2)[slot] = j;
  (c.wi + c.g->L.pk)[slot] = k; (c.wi + c.g->L.psgn)[slot] = -1;
  (c.wd + c.g->L.pex)[slot] = ex; (c.wd + c.g->L.pey)[slot] = ey;
  // eta.(p_i - p_j) >= bound with p = off[k-1] + S_{k-1} x
  (c.wd + c.g->L.pb_)[slot] = (c.wd + c.g->L.c_bound)[sidx] -
      (ex * (off[(2 * i) * K + k - 1] - off[(2 * j) * K + k - 1]) + ey * (off[(2 * i + 1) * K + k - 1] - off[(2 * j + 1) * K + k - 1]));
  (c.wd + c.g->L.pcv)[slot] = lib_dot(3, k - 1, 2, K - 1, h);
  (c.wd + c.g->L.pcp)[slot] = lib_dot(3, k - 1, 3, K - 1, h);
}

// rhs entry 2 (a_r . x_d - b_r) of list row r
SCP_DEV double polish_rhs_entry(Ctx& c, const PGeom& g, int r) {
  const PRow a = load_prow(c, r);
  const double* deq = c.wd + c.g->L.deq;
  double d0, d1;
  if (a.type < 4) { d0 = deq[2 * a.q]; d1 = deq[2 * a.q + 1]; }
  else {
    d0 = a.ex * (deq[2 * (2 * a.q)] - deq[2 * (2 * a.j)]) + a.ey * (deq[2 * (2 * a.q + 1)] - deq[2 * (2 * a.j + 1)]);
    d1 = a.ex * (deq[2 * (2 * a.q) + 1] - deq[2 * (2 * a.j) + 1]) + a.ey * (deq[2 * (2 * a.q + 1) + 1] - deq[2 * (2 * a.j + 1) + 1]);
  }
  const double axd = a.cv * (g.i00 * d0 + g.i01 * d1) + a.cp * (g.i01 * d0 + g.i11 * d1);
  return 2.0 * (axd - (c.wd + c.g->L.pb_)[r]);
}

// ---- active-set list maintenance and the dense solve (G0 + delta diag G0) y = rhs
// Kept per list: the exact Gram matrix G0 (full symmetric storage in the slot scratch, stride pcap; rows are closed
// forms and are recomputed in parallel for every list position that changed) and Ainv = (G0 + delta diag G0)^-1 as a
// PACKED symmetric matrix (row-major lower triangle: A(i,j), i >= j, at i(i+1)/2 + j -- appending a row appends
// storage).  During a polish attempt Ainv lives in SHARED memory on the GPU (the ADMM's hot arrays are parked in
// their global homes meanwhile); between attempts it is saved to the slot scratch.  A fresh list is inverted by
// symmetric Gauss-Jordan sweeps, later rows enter / leave with O(n^2) bordered-inverse updates / rank-one
// downdates -- the same algebra as round 1, which ran out of L2 (43 % of the solver's cycles).
SCP_DEV size_t sp_row(int i) { return ((size_t)i * (size_t)(i + 1)) / 2; }
SCP_DEV size_t sp_idx(int i, int j) { return i >= j ? sp_row(i) + j : sp_row(j) + i; }

// (i, j) of packed index e, branch free: float square root + two integer corrections (data-dependent loops cost
// ~30 cycles per trip on the GPU; measured: this loop body was 5x slower with an incremental (i, j) walk)
SCP_DEV void sp_decode(long long e, int& i, int& j) {
  int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
  r -= ((long long)sp_row(r) > e);
  r += ((long long)sp_row(r + 1) <= e);
  r -= ((long long)sp_row(r) > e);
  r += ((long long)sp_row(r + 1) <= e);
  i = r; j = (int)(e - (long long)sp_row(r));
}

// A_ij += s v_i v_j over the packed lower triangle of an n x n matrix, skipping row/column `skip` (-1: none)
#ifndef SCP_EMU
// One warp per row of the packed triangle (rows dealt cyclically to the warps, lanes along the row): no index decoding,
// conflict-free, and -- SH: A and v live in shared memory -- LDS/STS through the dynamic shared array instead of generic
// loads (ncu, round 2: the flat decode loop was 13 % of the solver's samples, 58 % of them long-scoreboard stalls).
template <bool SH>
__device__ __forceinline__ void sp_rank1_rows(double* A, int n, const double* v, double s, int skip) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int oa = SH ? (int)(A - smem) : 0, ov = SH ? (int)(v - smem) : 0;
  for (int i = threadIdx.x >> 5; i < n; i += nw) {
    if (i == skip) continue;
    const double svi = s * (SH ? smem[ov + i] : v[i]);
    const int r = (int)sp_row(i);
    for (int j = lane; j <= i; j += 32) {
      if (j == skip) continue;
      if (SH) smem[oa + r + j] += svi * smem[ov + j];
      else A[r + j] += svi * v[j];
    }
  }
  __syncthreads();
}
#endif
SCP_DEV void sp_rank1(Ctx& c, double* A, int n, const double* v, double s, int skip) {
#ifndef SCP_EMU
  if (c.team == 1) {
    if (__isShared(A) && __isShared(v)) sp_rank1_rows<true>(A, n, v, s, skip);
    else sp_rank1_rows<false>(A, n, v, s, skip);
    return;
  }
#endif
  const long long total = (long long)sp_row(n);
  SCP_PHASE(c) {
    for (long long e = tid; e < total; e += c.nthreads) {
      int i, j;
      sp_decode(e, i, j);
      if (i != skip && j != skip) A[e] += s * v[i] * v[j];
    }
  }
  SCP_SYNC(c);
}

// out = A x for the packed symmetric A.  GPU, one-CTA team: one warp per row, lanes stride over the columns and the
// partial sums meet in a shuffle reduction (a dependent fp64 FMA chain costs 45 cycles per link on B200).
SCP_DEV void sp_matvec(Ctx& c, const double* A, int n, const double* x, double* out, int accumulate) {
#ifndef SCP_EMU
  if (c.team == 1) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (__isShared(A) && __isShared(x)) {          // shared-memory operands: LDS with 32-bit addresses
      const int oa = (int)(A - smem), ox = (int)(x - smem);
      for (int i = threadIdx.x >> 5; i < n; i += nw) {
        const int r = oa + (int)sp_row(i);
        double a = 0.0;
        for (int j = lane; j <= i; j += 32) a += smem[r + j] * smem[ox + j];
        for (int j = i + 1 + lane; j < n; j += 32) a += smem[oa + (int)sp_row(j) + i] * smem[ox + j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) out[i] = (accumulate ? out[i] : 0.0) + a;
      }
      __syncthreads();
      return;
    }
    for (int i = threadIdx.x >> 5; i < n; i += nw) {
      const double* row = A + sp_row(i);
      double a = 0.0;
      for (int j = lane; j <= i; j += 32) a += row[j] * x[j];
      for (int j = i + 1 + lane; j < n; j += 32) a += A[sp_row(j) + i] * x[j];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
      if (lane == 0) out[i] = (accumulate ? out[i] : 0.0) + a;
    }
    __syncthreads();
    return;
  }
#endif
  SCP_PHASE(c) {
    for (int i = tid; i < n; i += c.nthreads) {
      const double* row = A + sp_row(i);
      double a0 = 0.0, a1 = 0.0;
      int j = 0;
      for (; j + 1 <= i; j += 2) { a0 += row[j] * x[j]; a1 += row[j + 1] * x[j + 1]; }
      if (j <= i) a0 += row[j] * x[j];
      for (j = i + 1; j < n; ++j) a1 += A[sp_row(j) + i] * x[j];
      out[i] = (accumulate ? out[i] : 0.0) + (a0 + a1);
    }
  }
  SCP_SYNC(c);
}

// Ainv of the first n list rows from scratch: A <- G0 (+ relative diagonal shift), then n symmetric Gauss-Jordan
// sweeps (sweep k: A_ij -= A_ik A_kj / A_kk, A_ik /= A_kk, A_kk = -1/A_kk) leave -inverse; negated at the end.
// Returns 0 on a non-positive pivot (dependent active rows).
SCP_DEV int polish_invert(Ctx& c, double* A, int n) {
  const int ld = c.g->L.pcap;
  const double* G0 = c.wd + c.g->L.pG;
  double* col = c.sh + 3 * (size_t)c.rs;          // reduction column 3 is free here (n <= pcap <= 1024 <= rs)
  SCP_PHASE(c) {
    for (int e = tid; e < n * n; e += c.nthreads) {
      const int i = e / n, j = e - i * n;
      if (j <= i) A[sp_row(i) + j] = (i == j) ? G0[(size_t)i * ld + j] * (1.0 + 1e-11) : G0[(size_t)j * ld + i];
    }
  }
  SCP_SYNC(c);
  for (int k = 0; k < n; ++k) {
    SCP_PHASE(c) { for (int i = tid; i < n; i += c.nthreads) col[i] = A[sp_idx(i, k)]; }
    SCP_SYNC(c);
    const double akk = col[k];
    if (!(akk > 0.0)) return 0;                   // uniform: every thread reads the same value
    const double d = 1.0 / akk;
    sp_rank1(c, A, n, col, -d, k);
    SCP_PHASE(c) {
      for (int i = tid; i < n; i += c.nthreads) A[sp_idx(i, k)] = (i == k) ? -d : col[i] * d;
    }
    SCP_SYNC(c);
  }
  // after all sweeps the pivots are negative definite: A = -(G)^-1 with the swept sign convention folded in
  SCP_PHASE(c) {
    const long long total = (long long)sp_row(n);
    for (long long e = tid; e < total; e += c.nthreads) A[e] = -A[e];
  }
  SCP_SYNC(c);
  return 1;
}

// Row in list slot n joins Ainv (bordered inverse, Schur complement of the new row).  0 on a non-positive complement.
SCP_DEV int polish_add(Ctx& c, double* A, int n) {
  const int ld = c.g->L.pcap;
  const double* G0 = c.wd + c.g->L.pG;
  double* gv = c.sh + 3 * (size_t)c.rs;           // n+1 Gram entries of the new row
  double* u = c.sh + 2 * (size_t)c.rs;            // n
  double* red = c.sh;
  SCP_PHASE(c) { for (int i = tid; i <= n; i += c.nthreads) gv[i] = G0[(size_t)n * ld + i]; }
  SCP_SYNC(c);
  sp_matvec(c, A, n, gv, u, 0);
  SCP_PHASE(c) {
    double part = 0.0;
    for (int i = tid; i < n; i += c.nthreads) part += u[i] * gv[i];
    red[tid] = part;
  }
  SCP_SYNC(c);
  const double gtu = reduce_finish(c, 0, 1);
  const double sch = gv[n] * (1.0 + 1e-11) - gtu;
  if (!(sch > 0.0)) return 0;
  const double is = 1.0 / sch;
  sp_rank1(c, A, n, u, is, -1);
  SCP_PHASE(c) {
    double* row = A + sp_row(n);
    for (int i = tid; i < n; i += c.nthreads) row[i] = -u[i] * is;
    if (tid == 0) row[n] = is;
  }
  SCP_SYNC(c);
  return 1;
}

// Row at list position p leaves Ainv (rank-one downdate); the last row (n-1) moves into the hole.
SCP_DEV void polish_drop(Ctx& c, double* A, int p, int n) {
  double* cpv = c.sh + 3 * (size_t)c.rs;
  double* lastrow = c.sh + 2 * (size_t)c.rs;
  SCP_PHASE(c) { for (int i = tid; i < n; i += c.nthreads) cpv[i] = A[sp_idx(p, i)]; }
  SCP_SYNC(c);
  sp_rank1(c, A, n, cpv, -1.0 / cpv[p], p);
  const int last = n - 1;
  if (p != last) {
    SCP_PHASE(c) { for (int i = tid; i < n; i += c.nthreads) lastrow[i] = A[sp_row(last) + i]; }
    SCP_SYNC(c);
    SCP_PHASE(c) {
      for (int i = tid; i < last; i += c.nthreads) if (i != p) A[sp_idx(p, i)] = lastrow[i];
      if (tid == 0) A[sp_row(p) + p] = lastrow[last];
    }
    SCP_SYNC(c);
  }
}

// y = Ainv rhs, refined against the exact G0 until the residual is at rounding level.  The incrementally updated
// inverse drifts (and is ruined by a nearly dependent row), so the residual is the health check of the list:
// returns 0 when it cannot be brought down, and the caller rebuilds the list from scratch on its next attempt.
SCP_DEV int polish_solve(Ctx& c, const double* A, int n) {
  const int ld = c.g->L.pcap;
  const double* G0 = c.wd + c.g->L.pG;
  const double* rhs = c.wd + c.g->L.prhs;
  double* y = c.wd + c.g->L.py;
  double* z = c.sh + 3 * (size_t)c.rs;
  double* ys = c.sh + 2 * (size_t)c.rs;           // y while it is being refined (shared memory on the GPU)
  double* red = c.sh;
  double bnorm = 0.0;
  int ok = -1;
  for (int sweep = 0; sweep < 6 && ok < 0; ++sweep) {
#ifndef SCP_EMU
    if (c.team == 1 && sweep > 0) {
      // z = rhs - G0 ys, one warp per row (G0 is symmetric: row r is read along its contiguous column)
      const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
      for (int r = threadIdx.x >> 5; r < n; r += nw) {
        const double* g0 = G0 + (size_t)r * ld;
        double a = 0.0;
        for (int q2 = lane; q2 < n; q2 += 32) a += g0[q2] * ys[q2];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) z[r] = -a;
      }
      __syncthreads();
    }
#endif
    SCP_PHASE(c) {
      double zm = 0.0, bm = 0.0;
      for (int r = tid; r < n; r += c.nthreads) {
        double acc = rhs[r];
#ifndef SCP_EMU
        if (c.team == 1) { if (sweep > 0) acc += z[r]; }
        else
#endif
        if (sweep > 0) for (int q2 = 0; q2 < n; ++q2) acc -= G0[(size_t)q2 * ld + r] * ys[q2];
        z[r] = acc;
        zm = SCP_FMAX(zm, fabs(acc)); bm = SCP_FMAX(bm, fabs(rhs[r]));
      }
      red[tid] = zm; red[c.rs + tid] = bm;
    }
    SCP_SYNC(c);
    const double znorm = reduce_finish(c, 0, 0);
    if (sweep == 0) bnorm = reduce_finish(c, 1, 0);
    if (!(znorm == znorm)) ok = 0;
    else if (sweep > 0 && znorm <= 1e-11 * (1.0 + bnorm)) ok = 1;
    else if (sweep == 5) ok = znorm <= 1e-9 * (1.0 + bnorm);
    else sp_matvec(c, A, n, z, ys, sweep > 0);
  }
  if (ok == 1) {
    SCP_PHASE(c) { for (int r = tid; r < n; r += c.nthreads) y[r] = ys[r]; }
    SCP_SYNC(c);
  }
  return ok == 1;
}

// Applies the pending mark changes (pdec vs pmark / pcdec vs pcmark, score >= threshold): rows leave the list (the
// last row moves into the hole) and join it (appended); the Gram rows/columns and right-hand-side entries of every
// list position that changed are recomputed in parallel; then Ainv follows, one row at a time (or, for a fresh list,
// by one inversion).  Returns the new n, or -1 when the set outgrows its storage or an update breaks down.
SCP_DEV int polish_apply(Ctx& c, const PGeom& g, double* A, int ncap, int n, int use_col, double thr_add, double thr_drop) {
  const int K = c.K, N = c.N, QK = c.Q * K;
  int* pmark = c.wi + c.g->L.pmark; int* pcmark = c.wi + c.g->L.pcmark;
  const int* pdec = c.wi + c.g->L.pdec; const int* pcdec = c.wi + c.g->L.pcdec;
  const double* pscore = c.wd + c.g->L.pscore; const double* pcscore = c.wd + c.g->L.pcscore;
  int* ppos = c.wi + c.g->L.ppos; int* pcpos = c.wi + c.g->L.pcpos; int* pid = c.wi + c.g->L.pid;
  int* chg = c.wi + c.g->L.chg;
  int* pslot = c.wi + c.g->L.pslot;               // list slot a change fills (-1: the change only removes a row)
  int* pdirty = c.wi + c.g->L.pdirty;             // list positions whose Gram row must be recomputed
  int* pdl = c.wi + c.g->L.pdl;                   // [0, nd): dirty positions; [pcap/2, ...): positions dropped, in order
  const int* pcown = c.wi + c.g->L.pcown;
  const int* coff = c.wi + c.g->L.coff; const int* cj = c.wi + c.g->L.c_j;
  const int total = 4 * QK + (use_col ? c.ncand : 0);
  const int np = c.np;
  const int per = (total + np - 1) / np;
  int* part = (int*)(c.sh + 2 * (size_t)c.rs);
  int* misc = (int*)(c.sh + 4 * (size_t)c.rs + 8);     // [1] new n, [2] dirty count, [3] drops
  int* drops = pdl + c.g->L.pcap / 2;
  SCP_PHASE(c) {
    int cnt = 0;
    if (tid < np) for (int u = tid * per; u < total && u < (tid + 1) * per; ++u) {
      if (u < 4 * QK) { const int nm = pdec[u]; cnt += (nm != pmark[u] && pscore[u] >= (nm != 0 ? thr_add : thr_drop)); }
      else { const int sidx = u - 4 * QK, nm = pcdec[sidx]; cnt += (pcown[sidx] >= 0 && nm != pcmark[sidx] && pcscore[sidx] >= (nm != 0 ? thr_add : thr_drop)); }
    }
    if (tid < np) part[tid] = cnt;
    for (int e = tid; e < c.g->L.pcap; e += c.nthreads) pdirty[e] = 0;
  }
  SCP_SYNC(c);
  int nchg = 0;
  for (int e = 0; e < np; ++e) nchg += part[e];
  const int chg_cap = c.g->L.pcap / 2;
  SCP_PHASE(c) {
    int base = 0;
    if (tid < np) for (int e = 0; e < tid; ++e) base += part[e];
    if (tid < np) for (int u = tid * per; u < total && u < (tid + 1) * per; ++u) {
      int hit;
      if (u < 4 * QK) { const int nm = pdec[u]; hit = (nm != pmark[u] && pscore[u] >= (nm != 0 ? thr_add : thr_drop)); }
      else { const int sidx = u - 4 * QK, nm = pcdec[sidx]; hit = (pcown[sidx] >= 0 && nm != pcmark[sidx] && pcscore[sidx] >= (nm != 0 ? thr_add : thr_drop)); }
      if (hit) { if (base < chg_cap) chg[base] = u; ++base; }
    }
  }
  SCP_SYNC(c);
  if (nchg > chg_cap) nchg = chg_cap;
  if (nchg == 0) return n;
  // (1) one thread: removals (list metadata only, the positions are replayed on Ainv below) and the slots of the
  //     rows that join
  SCP_PHASE(c) {
    if (tid == 0) {
      int* ptype = c.wi + c.g->L.ptype; int* pq = c.wi + c.g->L.pq; int* pj2 = c.wi + c.g->L.pj2;
      int* pk = c.wi + c.g->L.pk; int* psgn = c.wi + c.g->L.psgn;
      const size_t offs[6] = {c.g->L.pb_, c.g->L.pex, c.g->L.pey, c.g->L.pcv, c.g->L.pcp, c.g->L.prhs};
      int nn = n, ndrop = 0;
      for (int ci = 0; ci < nchg; ++ci) {
        const int id = chg[ci];
        const int is_col = id >= 4 * QK, sidx = is_col ? id - 4 * QK : 0;
        const int m = is_col ? pcmark[sidx] : pmark[id];
        if (m == 0) continue;
        const int p = is_col ? pcpos[sidx] : ppos[id], last = nn - 1;
        if (is_col) pcpos[sidx] = -1; else ppos[id] = -1;
        if (p != last) {
          ptype[p] = ptype[last]; pq[p] = pq[last]; pj2[p] = pj2[last]; pk[p] = pk[last]; psgn[p] = psgn[last]; pid[p] = pid[last];
          for (int f = 0; f < 6; ++f) (c.wd + offs[f])[p] = (c.wd + offs[f])[last];
          const int mid = pid[p];
          if (mid < 4 * QK) ppos[mid] = p; else pcpos[mid - 4 * QK] = p;
          pdirty[p] = 1;
        }
        drops[ndrop++] = p;
        nn = last;
      }
      int fail = 0;
      for (int ci = 0; ci < nchg; ++ci) {
        const int id = chg[ci];
        const int is_col = id >= 4 * QK, sidx = is_col ? id - 4 * QK : 0;
        const int joins = is_col ? (pcmark[sidx] == 0) : (pdec[id] != 0);
        if (joins && nn < ncap) { pslot[ci] = nn; pdirty[nn] = 1; ++nn; }
        else { pslot[ci] = -1; if (joins) fail = 1; }
      }
      misc[1] = fail ? -1 : nn; misc[3] = ndrop;
    }
  }
  SCP_SYNC(c);
  const int n_new = misc[1], ndrop = misc[3];
  if (n_new < 0) return -1;
  // (2) one thread per change: fill the slot of a joining row, update the marks
  SCP_PHASE(c) {
    for (int ci = tid; ci < nchg; ci += c.nthreads) {
      const int id = chg[ci], slot = pslot[ci];
      if (id < 4 * QK) {
        const int nm = pdec[id];
        if (slot >= 0) { polish_fill_dyn(c, slot, id / QK, id % QK, nm); pid[slot] = id; ppos[id] = slot; }
        pmark[id] = nm;
      } else {
        const int sidx = id - 4 * QK, nm = pcdec[sidx];
        const int t = pcown[sidx], ck = t / N, ci_ = t - ck * N, cjj = cj[sidx];
        if (slot >= 0) { polish_fill_col(c, slot, sidx, ck, ci_); pid[slot] = id; pcpos[sidx] = slot; }
        pcmark[sidx] = nm;
        for (int s2 = coff[ck * N + cjj]; s2 < coff[ck * N + cjj + 1]; ++s2) if (cj[s2] == ci_) pcmark[s2] = nm;   // mirror entry
      }
    }
    if (tid == 0) {
      int nd = 0;
      for (int p = 0; p < n_new; ++p) if (pdirty[p]) pdl[nd++] = p;
      misc[2] = nd;
    }
  }
  SCP_SYNC(c);
  // (3) Gram rows / columns and right-hand sides of the changed positions (final list)
  const int nd = misc[2];
  {
    const int ld = c.g->L.pcap;
    double* G0 = c.wd + c.g->L.pG;
    double* prhs = c.wd + c.g->L.prhs;
    SCP_PHASE(c) {
      for (int e = tid; e < nd * n_new; e += c.nthreads) {
        const int di = e / n_new, i = e - di * n_new, p = pdl[di];
        const PRow a = load_prow(c, p), b = load_prow(c, i);
        const double v = gram_entry(g, a, b);
        G0[(size_t)p * ld + i] = v; G0[(size_t)i * ld + p] = v;
        if (i == 0) prhs[p] = polish_rhs_entry(c, g, p);
      }
    }
    SCP_SYNC(c);
  }
  // (4) Ainv: replay the removals, then the rows that joined (slots n0 .. n_new-1 in order)
  int nn = n;
  for (int di = 0; di < ndrop; ++di) { polish_drop(c, A, drops[di], nn); --nn; }
  if (nn == 0 && n_new > 8) return polish_invert(c, A, n_new) ? n_new : -1;
  for (; nn < n_new; ++nn)
    if (!polish_add(c, A, nn)) return -1;
  return n_new;
}

SCP_DEV void polish_set_hot(Ctx& c, int parked) {
  double* p[7];
  for (int a = 0; a < 7; ++a) p[a] = (!parked && c.hot_s[a]) ? c.hot_s[a] : c.hot_g[a];
  c.a_P = p[0]; c.a_F = p[1]; c.a_x = p[2]; c.a_vp = p[3]; c.a_vv = p[4]; c.a_vj = p[5]; c.a_va = p[6];
}
// hot arrays -> global homes (park) / back (unpark); no-ops when nothing is parked
SCP_DEV void polish_park(Ctx& c, int unpark) {
  if (!c.pol_smem) return;
  const int QK = c.Q * c.K;
  SCP_PHASE(c) {
    for (int a = 0; a < 7; ++a) {
      if (!c.hot_s[a]) continue;
      double* src = unpark ? c.hot_g[a] : c.hot_s[a];
      double* dst = unpark ? c.hot_s[a] : c.hot_g[a];
      for (int e = tid; e < QK; e += c.nthreads) dst[e] = src[e];
    }
  }
  SCP_SYNC(c);
  polish_set_hot(c, !unpark);
}

// One full polish.  Returns 1 when the active set is stable (KKT certificate on the carried
// rows) and the ADMM state has been replaced by the exact solution; 0 leaves everything untouched.
SCP_DEV int polish(Ctx& c, int with_collisions, int max_rounds) {
  const int K = c.K, N = c.N, QK = c.Q * K, nch = (K + CH - 1) / CH;
  const double h = c.g->pb.time_step, ih = 1.0 / h;
  const PGeom g = make_pgeom(h, K);
  int* pmark = c.wi + c.g->L.pmark;
  int* pcmark = c.wi + c.g->L.pcmark;
  int* pdec = c.wi + c.g->L.pdec;
  int* pcdec = c.wi + c.g->L.pcdec;
  double* pscore = c.wd + c.g->L.pscore;
  double* pcscore = c.wd + c.g->L.pcscore;
  double* sc3 = c.sh + 3 * (size_t)c.rs;     // fourth reduction column
  const int full_rounds = 6;
  int* ppos = c.wi + c.g->L.ppos;
  int* pcpos = c.wi + c.g->L.pcpos;
  int* pcown = c.wi + c.g->L.pcown;
  const int* coff = c.wi + c.g->L.coff;
  const int* cj = c.wi + c.g->L.c_j;
  const double* ceta = c.wd + c.g->L.c_eta;
  const double* cb = c.wd + c.g->L.c_bound;
  double* lam = c.wd + c.g->L.lam;
  double* plam = c.wd + c.g->L.plam;
  double *vj = c.a_vj, *va = c.a_va, *vv = c.a_vv, *vp = c.a_vp;
  double *yj = c.wd + c.g->L.yj, *ya = c.wd + c.g->L.ya, *yv = c.wd + c.g->L.yv, *yp = c.wd + c.g->L.yp;
  double *off = c.wd + c.g->L.off, *deq = c.wd + c.g->L.deq;
  double *xt = c.wd + c.g->L.xt, *Pt = c.wd + c.g->L.Pt, *w = c.a_rhs, *FY = c.wd + c.g->L.FY;
  double* red = c.sh;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double lo[2] = {c.g->pb.space[0], c.g->pb.space[1]}, hi[2] = {c.g->pb.space[2], c.g->pb.space[3]};
  const int use_col = with_collisions && c.ncand > 0;
  const double ptol = 1e-9, dtol = 1e-9, etol = 1e-7;

  // initial guess of W from the ADMM state: y != 0  <=>  v outside its box; lam > 0.  The list, its inverse and the
  // marks survive between attempts on the same candidate set, so only the difference to the new guess is applied.
  const int fresh = !(c.pol_valid && c.pol_use_col == use_col && !c.pol_col_stale);
  const int col_fresh = fresh;
  long long tq = SCP_CLOCK();
  SCP_PHASE(c) {
    for (int e = tid; e < QK; e += c.nthreads) {
      int q = e / K, k = e - q * K;
      double v = va[e];
      const int ma = v > al ? 1 : (v < -al ? -1 : 0);
      int mj = 0, mv = 0, mp = 0;
      if (k < K - 1) {
        v = vj[e]; mj = v > jl ? 1 : (v < -jl ? -1 : 0);
        double v0q = c.v0[q];
        v = vv[e]; mv = v > vl - v0q ? 1 : (v < -vl - v0q ? -1 : 0);
        int a2 = q & 1;
        v = vp[e]; mp = v > hi[a2] - off[e] ? 1 : (v < lo[a2] - off[e] ? -1 : 0);
      }
      pdec[e] = mj; pdec[QK + e] = ma; pdec[2 * QK + e] = mv; pdec[3 * QK + e] = mp;
      for (int cls = 0; cls < 4; ++cls) {
        pscore[cls * QK + e] = 1.0;
        if (fresh) { pmark[cls * QK + e] = 0; ppos[cls * QK + e] = -1; }
      }
    }
    if (use_col)
      for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
        int k = 1 + t / N, i = t - (k - 1) * N;
        for (int sidx = coff[k * N + i]; sidx < coff[k * N + i + 1]; ++sidx) {
          pcscore[sidx] = 1.0;
          pcdec[sidx] = lam[((size_t)k * N + i) * N + cj[sidx]] > 0.0;
          if (col_fresh) {
            pcmark[sidx] = 0; pcpos[sidx] = -1;
            pcown[sidx] = cj[sidx] > i ? k * N + i : -1;     // the i<j entry is the row's identity
          }
        }
      }
  }
  SCP_SYNC(c);
  int n = fresh ? 0 : c.pol_n;
  double ta = 0.0, td = 0.0;               // score thresholds of the pending decisions (0: apply all of them)
  // where Ainv lives during this attempt: the parked shared-memory region when the set is expected to fit (count of
  // marked rows + 12; a set that outgrows the region fails the attempt and later attempts use the scratch), else the slot scratch; a kept list comes back from the slot scratch
  double* Asave = c.wd + c.g->L.pL;
  double* A = Asave;
  int ncap = c.g->L.pcap;
  if (c.pol_smem) {
    SCP_PHASE(c) {
      double cnt = 0.0;
      for (int e = tid; e < 4 * QK; e += c.nthreads) cnt += (pdec[e] != 0);
      if (use_col) for (int sidx = tid; sidx < c.ncand; sidx += c.nthreads) cnt += (pcown[sidx] >= 0 && pcdec[sidx] != 0);
      red[tid] = cnt;
    }
    SCP_SYNC(c);
    const int guess = (int)reduce_finish(c, 0, 1);
    int cap_n = (int)((sqrt(8.0 * (double)c.pol_smem_doubles + 1.0) - 1.0) * 0.5);
    while (sp_row(cap_n) + cap_n + 1 > c.pol_smem_doubles) --cap_n;
    if (cap_n > c.g->L.pcap) cap_n = c.g->L.pcap;
    const int want = (guess > n ? guess : n);
    if (!c.pol_no_smem && want + 12 <= cap_n) {
      A = c.pol_smem; ncap = cap_n;
      if (!fresh && n > 0) {
        const long long total = (long long)sp_row(n);
        SCP_PHASE(c) { for (long long e = tid; e < total; e += c.nthreads) A[e] = Asave[e]; }
        SCP_SYNC(c);
      }
    }
  }
  c.pol_A = A;
  for (int round = 0;; ++round) {
    // pending mark changes -> list / inverse: the initial guess before round 0, the decisions of round-1 afterwards
    // (ONE call site: polish_apply and the O(n^2) updates it calls exist once in the kernel)
    if (round > 0) tq = SCP_CLOCK();
    n = polish_apply(c, g, A, ncap, n, use_col, ta, td);
    if (round == 0) { c.t_pbuild += SCP_CLOCK() - tq; c.pol_use_col = use_col; c.pol_col_stale = 0; }
    else c.t_papply += SCP_CLOCK() - tq;
    c.pol_n = n; c.pol_valid = n >= 0;
    if (n < 0) { if (A != Asave) c.pol_no_smem = 1; return 0; }
    if (round >= max_rounds) break;
    c.polish_rounds++;
    tq = SCP_CLOCK();
    if (n > 0 && !polish_solve(c, A, n)) { c.pol_valid = 0; return 0; }
    c.t_psolve += SCP_CLOCK() - tq;
    tq = SCP_CLOCK();
    // scatter multipliers to the dense arrays
    const double* y = c.wd + c.g->L.py;
    SCP_PHASE(c) {
      for (int e = tid; e < QK; e += c.nthreads) { yj[e] = 0; ya[e] = 0; yv[e] = 0; yp[e] = 0; FY[e] = 0; }
      if (use_col) for (int sidx = tid; sidx < c.ncand; sidx += c.nthreads) plam[sidx] = 0.0;
    }
    SCP_SYNC(c);
    SCP_PHASE(c) {
      for (int r = tid; r < n; r += c.nthreads) {
        PRow a = load_prow(c, r);
        if (a.type == 0) yj[a.q * K + a.k] = y[r];
        else if (a.type == 1) ya[a.q * K + a.k] = y[r];
        else if (a.type == 2) yv[a.q * K + a.k] = y[r];
        else if (a.type == 3) yp[a.q * K + a.k] = y[r];
        else {
          // collision row (lower bound): lam = -y on both owners' entries
          for (int side = 0; side < 2; ++side) {
            int own = side ? a.j : a.q, oth = side ? a.q : a.j;
            for (int sidx = coff[a.k * N + own]; sidx < coff[a.k * N + own + 1]; ++sidx)
              if (cj[sidx] == oth) plam[sidx] = -y[r];
          }
        }
      }
    }
    SCP_SYNC(c);
    if (use_col) {
      SCP_PHASE(c) {
        for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
          int k = 1 + t / N, i = t - (k - 1) * N;
          double fx = 0, fy = 0;
          for (int sidx = coff[k * N + i]; sidx < coff[k * N + i + 1]; ++sidx) { fx += plam[sidx] * ceta[2 * sidx]; fy += plam[sidx] * ceta[2 * sidx + 1]; }
          FY[(2 * i) * K + k] = fx; FY[(2 * i + 1) * K + k] = fy;
        }
      }
      SCP_SYNC(c);
    }
    transpose_rows(c, 2);                   // w = A'y
    // x = x_d - 1/2 Pi w ; per agent-axis: cw = C w (two dots), then the update
    double* cw = c.wd + c.g->L.scr;         // 2 per q
    SCP_PHASE(c) {
      for (int q = tid; q < c.Q; q += c.nthreads) {
        double a0 = 0, a1 = 0;
        for (int k = 0; k < K; ++k) { double wv = w[q * K + k]; a0 += h * wv; a1 += h * h * ((double)(K - 1 - k) + 0.5) * wv; }
        // coefficients of C' : H^-1 (d + 1/2 cw)
        double r0 = deq[2 * q] + 0.5 * a0, r1 = deq[2 * q + 1] + 0.5 * a1;
        cw[2 * q] = g.i00 * r0 + g.i01 * r1; cw[2 * q + 1] = g.i01 * r0 + g.i11 * r1;
      }
    }
    SCP_SYNC(c);
    SCP_PHASE(c) {
      for (int e = tid; e < QK; e += c.nthreads) {
        int q = e / K, k = e - q * K;
        xt[e] = -0.5 * w[e] + h * cw[2 * q] + h * h * ((double)(K - 1 - k) + 0.5) * cw[2 * q + 1];
      }
    }
    SCP_SYNC(c);
    // rows at xt: check violated / wrong-sign, update marks; positions Pt
    double* s1 = c.wd + c.g->L.scr + 2 * c.Q;
    double* s2 = s1 + (size_t)c.Q * nch;
    SCP_PHASE(c) {
      for (int t = tid; t < c.Q * nch; t += c.nthreads) {
        int q = t / nch, ch = t - q * nch;
        int k0 = ch * CH, k1 = k0 + CH < K ? k0 + CH : K;
        double a1 = 0, a2 = 0;
        for (int k = k0; k < k1; ++k) { a1 += xt[q * K + k]; a2 += a1; }
        s1[t] = a1; s2[t] = a2;
      }
    }
    SCP_SYNC(c);
    SCP_PHASE(c) {
      double changes = 0.0, broken = 0.0, madd = 0.0, mdrop = 0.0;
      for (int t = tid; t < c.Q * nch; t += c.nthreads) {
        int q = t / nch, ch = t - q * nch;
        int k0 = ch * CH, k1 = k0 + CH < K ? k0 + CH : K;
        double c1 = 0, c2 = 0;
        for (int cc = 0; cc < ch; ++cc) { c2 += s2[q * nch + cc] + (double)CH * c1; c1 += s1[q * nch + cc]; }
        const double v0q = c.v0[q];
        const int ax = q & 1;
        for (int k = k0; k < k1; ++k) {
          const int e = q * K + k;
          const double xk = xt[e];
          c1 += xk; c2 += c1;
          const double rv_ = h * c1, rp_ = h * h * (c2 - 0.5 * c1);
          if (k + 1 < K) Pt[q * K + k + 1] = off[e] + rp_;
          // class, value, lower, upper, multiplier
          for (int cls = 0; cls < 4; ++cls) {
            if (cls != 1 && k >= K - 1) continue;
            double val, lw, up, ym;
            if (cls == 0) { val = (xt[e + 1] - xk) * ih; lw = -jl; up = jl; ym = yj[e]; }
            else if (cls == 1) { val = xk; lw = -al; up = al; ym = ya[e]; }
            else if (cls == 2) { val = rv_; lw = -vl - v0q; up = vl - v0q; ym = yv[e]; }
            else { val = rp_; lw = lo[ax] - off[e]; up = hi[ax] - off[e]; ym = yp[e]; }
            int m = pmark[cls * QK + e], nm = m;
            const double sc = 1.0 + SCP_FMAX(fabs(lw), fabs(up));
            double score = 0.0;
            if (m == 0) {
              if (val > up + ptol * sc) { nm = 1; score = (val - up) / sc; }
              else if (val < lw - ptol * sc) { nm = -1; score = (lw - val) / sc; }
            } else if (m > 0) { if (ym < -dtol) { nm = 0; score = -ym; } if (fabs(val - up) > etol * sc) broken += 1.0; }
            else { if (ym > dtol) { nm = 0; score = ym; } if (fabs(val - lw) > etol * sc) broken += 1.0; }
            pdec[cls * QK + e] = nm; pscore[cls * QK + e] = score;
            if (nm != m) { changes += 1.0; if (nm != 0) madd = SCP_FMAX(madd, score); else mdrop = SCP_FMAX(mdrop, score); }
          }
        }
        if (ch == 0) Pt[q * K] = c.p0[q];
      }
      red[tid] = changes; red[c.rs + tid] = broken; red[2 * (size_t)c.rs + tid] = madd;
      sc3[tid] = mdrop;
    }
    SCP_SYNC(c);
    double changes = reduce_finish(c, 0, 1);
    double broken = reduce_finish(c, 1, 1);
    double madd = reduce_finish(c, 2, 0);
    SCP_PHASE(c) { red[tid] = sc3[tid]; }
    SCP_SYNC(c);
    double mdrop = reduce_finish(c, 0, 0);
    if (use_col) {
      SCP_PHASE(c) {
        double ch2 = 0.0, br2 = 0.0, ma2 = 0.0, md2 = 0.0;
        for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
          int k = 1 + t / N, i = t - (k - 1) * N;
          const double pix = Pt[(2 * i) * K + k], piy = Pt[(2 * i + 1) * K + k];
          for (int sidx = coff[k * N + i]; sidx < coff[k * N + i + 1]; ++sidx) {
            const int j = cj[sidx];
            const double gval = ceta[2 * sidx] * (pix - Pt[(2 * j) * K + k]) + ceta[2 * sidx + 1] * (piy - Pt[(2 * j + 1) * K + k]);
            int m = pcmark[sidx], nm = m;
            double score = 0.0;
            if (!m) { if (gval < cb[sidx] - ptol) { nm = 1; score = cb[sidx] - gval; } }
            else { if (plam[sidx] < -dtol) { nm = 0; score = -plam[sidx]; } if (fabs(gval - cb[sidx]) > etol) br2 += 1.0; }
            pcdec[sidx] = nm; pcscore[sidx] = score;
            if (nm != m) { if (j > i) ch2 += 1.0; if (nm) ma2 = SCP_FMAX(ma2, score); else md2 = SCP_FMAX(md2, score); }
          }
        }
        red[tid] = ch2; red[c.rs + tid] = br2; red[2 * (size_t)c.rs + tid] = ma2; sc3[tid] = md2;
      }
      SCP_SYNC(c);
      changes += reduce_finish(c, 0, 1);
      broken += reduce_finish(c, 1, 1);
      madd = SCP_FMAX(madd, reduce_finish(c, 2, 0));
      SCP_PHASE(c) { red[tid] = sc3[tid]; }
      SCP_SYNC(c);
      mdrop = SCP_FMAX(mdrop, reduce_finish(c, 0, 0));
    }
#ifdef SCP_EMU_DEBUG
    fprintf(stderr, "   round %d n=%d changes=%g broken=%g\n", round, n, changes, broken);
    if (broken > 0.0) fprintf(stderr, "  POLISH BROKEN at round %d n=%d\n", round, n);
#endif
    c.t_peval += SCP_CLOCK() - tq;
    if (broken > 0.0) return 0;             // active rows not met as equalities: inconsistent (infeasible) set
    // thrashing: from the third round on a convergent primal-dual active-set iteration has few decisions left
    // (measured on config 2: <= 20 % of the set; the bar is 30 %); the attempts that keep changing a third of a 250-row set every
    // round are the ones on infeasible subproblems -- 40 rounds of O(n^2) updates each, the tail of a batch
    if (round >= 2 && changes > 0.3 * (double)n + 8.0) return 0;
    {
      // the decisions are applied at the top of the next round: all of them in the first rounds (primal-dual active
      // set step); afterwards only the worst violated row and the worst wrong-sign multiplier per round, which
      // breaks the cycles the full step can enter
      const int careful = round >= full_rounds;
      ta = careful ? madd * (1.0 - 1e-12) : 0.0; td = careful ? mdrop * (1.0 - 1e-12) : 0.0;
    }
#ifdef SCP_EMU_DEBUG
    if (changes == 0.0) fprintf(stderr, "  POLISH OK after %d rounds n=%d\n", round + 1, n);
    else if (round == max_rounds - 1) fprintf(stderr, "  POLISH FAIL after %d rounds n=%d changes=%g\n", round + 1, n, changes);
#endif
    if (changes == 0.0) {
      // accept: x, P, and an ADMM state consistent with (x, y): v = bound + y/(rho r) on active rows,
      // v = row value elsewhere; lam = plam; F = FY (2 lam' - lam with lam' = lam)
      double* x = c.a_x;
      double* P = c.a_P;
      double* F = c.a_F;
      for (int e2 = 0; e2 < 1; ++e2) {
        SCP_PHASE(c) { for (int e = tid; e < QK; e += c.nthreads) { x[e] = xt[e]; P[e] = Pt[e]; F[e] = FY[e]; } }
        SCP_SYNC(c);
      }
      forward_rows(c, 0);                   // v := A x, posrow/velrow/P
      const double rho = c.rho;
      SCP_PHASE(c) {
        for (int r = tid; r < n; r += c.nthreads) {
          PRow a = load_prow(c, r);
          if (a.type == 4) continue;
          const int e = a.q * K + a.k;
          if (a.type == 0) vj[e] += y[r] / (rho * c.g->tb.rj[a.k]);
          else if (a.type == 1) va[e] += y[r] / (rho * c.g->tb.ra[a.k]);
          else if (a.type == 2) vv[e] += y[r] / (rho * c.g->tb.rv[a.k]);
          else vp[e] += y[r] / (rho * c.g->tb.rp[a.k]);
        }
        if (use_col)
          for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
            int k = 1 + t / N, i = t - (k - 1) * N;
            for (int sidx = coff[k * N + i]; sidx < coff[k * N + i + 1]; ++sidx)
              lam[((size_t)k * N + i) * N + cj[sidx]] = SCP_FMAX(plam[sidx], 0.0);
          }
      }
      SCP_SYNC(c);
      return 1;
    }
  }
  return 0;
}

// ------------------------------------------------------------------ primal infeasibility
// OSQP's certificate (paper section 3.4) for {C x = d, l <= A x <= u}: the last multiplier step
// dy (box rows: rho r (A x - z); collision rows: -dlam) together with dmu = -(CC')^-1 C A'dy
// certifies infeasibility when  ||A'dy + C'dmu||_inf <= eps ||dy||_inf  and
// u'(dy)+ + l'(dy)- + d'dmu <= -eps ||dy||_inf.  Called at check iterations (rows stored, FD/dlam fresh).
SCP_DEV int primal_infeasible(Ctx& c, int with_collisions) {
  const int K = c.K, N = c.N, QK = c.Q * K;
  const double h = c.g->pb.time_step, ih = 1.0 / h, rho = c.rho;
  const double eps = 1e-3;   // OSQP's eps_prim_inf default is 1e-4 on its scaled problem; rows here are unscaled
  const PGeom g = make_pgeom(h, K);
  double* FD = c.wd + c.g->L.Pt;
  const int use_col = with_collisions && c.ncand > 0;
  if (!use_col) { SCP_PHASE(c) { for (int e = tid; e < QK; e += c.nthreads) FD[e] = 0.0; } SCP_SYNC(c); }
  transpose_rows(c, 3);                       // rhs <- A'dy
  double* w = c.a_rhs;
  double* dmu = c.wd + c.g->L.scr;            // 2 per q
  const double* deq = c.wd + c.g->L.deq;
  double* red = c.sh;
  SCP_PHASE(c) {
    for (int q = tid; q < c.Q; q += c.nthreads) {
      double a0 = 0, a1 = 0;
      for (int k = 0; k < K; ++k) { double wv = w[q * K + k]; a0 += h * wv; a1 += h * h * ((double)(K - 1 - k) + 0.5) * wv; }
      dmu[2 * q] = -(g.i00 * a0 + g.i01 * a1); dmu[2 * q + 1] = -(g.i01 * a0 + g.i11 * a1);
    }
  }
  SCP_SYNC(c);
  const double *x = c.a_x, *vj = c.a_vj, *va = c.a_va, *vv = c.a_vv, *vp = c.a_vp;
  const double *velrow = c.wd + c.g->L.velrow, *posrow = c.wd + c.g->L.posrow;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double lo[2] = {c.g->pb.space[0], c.g->pb.space[1]}, hi[2] = {c.g->pb.space[2], c.g->pb.space[3]};
  SCP_PHASE(c) {
    double gn = 0.0, yn = 0.0, sup = 0.0;
    for (int e = tid; e < QK; e += c.nthreads) {
      int q = e / K, k = e - q * K;
      gn = SCP_FMAX(gn, fabs(w[e] + h * dmu[2 * q] + h * h * ((double)(K - 1 - k) + 0.5) * dmu[2 * q + 1]));
      double z = clampd(va[e], -al, al), dy = rho * c.g->tb.ra[k] * (x[e] - z);
      yn = SCP_FMAX(yn, fabs(dy)); sup += al * fabs(dy);                   // u dy+ + l dy- with u = -l = al
      if (k < K - 1) {
        z = clampd(vj[e], -jl, jl); dy = rho * c.g->tb.rj[k] * ((x[e + 1] - x[e]) * ih - z);
        yn = SCP_FMAX(yn, fabs(dy)); sup += jl * fabs(dy);
        const double v0q = c.v0[q], offe = c.p0[q] + h * (double)(k + 1) * v0q;
        z = clampd(vv[e], -vl - v0q, vl - v0q); dy = rho * c.g->tb.rv[k] * (velrow[e] - z);
        yn = SCP_FMAX(yn, fabs(dy)); sup += dy > 0 ? (vl - v0q) * dy : (-vl - v0q) * dy;
        const int a2 = q & 1;
        z = clampd(vp[e], lo[a2] - offe, hi[a2] - offe); dy = rho * c.g->tb.rp[k] * (posrow[e] - z);
        yn = SCP_FMAX(yn, fabs(dy)); sup += dy > 0 ? (hi[a2] - offe) * dy : (lo[a2] - offe) * dy;
      }
      if (k == 0) sup += deq[2 * q] * dmu[2 * q] + deq[2 * q + 1] * dmu[2 * q + 1];
    }
    red[tid] = gn; red[c.rs + tid] = yn; red[2 * (size_t)c.rs + tid] = sup;
  }
  SCP_SYNC(c);
  double gn = reduce_finish(c, 0, 0), yn = reduce_finish(c, 1, 0), sup = reduce_finish(c, 2, 1);
  if (use_col) {
    const int* coff = c.wi + c.g->L.coff;
    const int* cj = c.wi + c.g->L.c_j;
    const double* ceta = c.wd + c.g->L.c_eta;
    const double* cb = c.wd + c.g->L.c_bound;
    const double* dlam = c.wd + c.g->L.plam;
    SCP_PHASE(c) {
      double yn2 = 0.0, sup2 = 0.0;
      for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
        int k = 1 + t / N, i = t - (k - 1) * N;
        for (int sidx = coff[k * N + i]; sidx < coff[k * N + i + 1]; ++sidx) {
          const int j = cj[sidx];
          if (j < i) continue;                       // each row once
          const double dl = dlam[sidx];              // dy = -dl ; u = +inf: only dy <= 0 (dl >= 0) can carry the certificate
          if (dl <= 0.0) continue;
          const double offi = ceta[2 * sidx] * ((c.p0[2 * i] + h * k * c.v0[2 * i]) - (c.p0[2 * j] + h * k * c.v0[2 * j])) +
                              ceta[2 * sidx + 1] * ((c.p0[2 * i + 1] + h * k * c.v0[2 * i + 1]) - (c.p0[2 * j + 1] + h * k * c.v0[2 * j + 1]));
          yn2 = SCP_FMAX(yn2, dl);
          sup2 += (cb[sidx] - offi) * (-dl);         // l dy-  (row coordinates)
        }
      }
      red[tid] = yn2; red[c.rs + tid] = sup2;
    }
    SCP_SYNC(c);
    yn = SCP_FMAX(yn, reduce_finish(c, 0, 0));
    sup += reduce_finish(c, 1, 1);
  }
#ifdef SCP_EMU_DEBUG
  { static int cnt = 0; if (++cnt % 40 == 0) fprintf(stderr, "   pinf gn=%.3e yn=%.3e sup=%.3e\n", gn, yn, sup); }
#endif
  if (!(yn > 1e-10)) return 0;
  return (gn <= eps * yn) && (sup <= -eps * yn);
}

// Signature (count, weighted id sum) of the rows the ADMM state currently marks active: box rows with
// v outside the box (y != 0), collision rows with lam > 0.  Two equal signatures at consecutive
// residual checks = the active set has settled -> worth a polish attempt.
SCP_DEV void active_signature(Ctx& c, int with_collisions, double* count, double* idsum) {
  const int K = c.K, N = c.N, QK = c.Q * K;
  const double *vj = c.a_vj, *va = c.a_va, *vv = c.a_vv, *vp = c.a_vp;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double lo[2] = {c.g->pb.space[0], c.g->pb.space[1]}, hi[2] = {c.g->pb.space[2], c.g->pb.space[3]};
  const double h = c.g->pb.time_step;
  double* red = c.sh;
  SCP_PHASE(c) {
    double n = 0.0, sidsum = 0.0;
    for (int e = tid; e < QK; e += c.nthreads) {
      int q = e / K, k = e - q * K;
      double v = va[e];
      int m = (v > al) ? 2 : ((v < -al) ? 1 : 0);
      if (m) { n += 1.0; sidsum += (double)((4 * e + 0) * 3 + m); }
      if (k < K - 1) {
        v = vj[e]; m = (v > jl) ? 2 : ((v < -jl) ? 1 : 0);
        if (m) { n += 1.0; sidsum += (double)((4 * e + 1) * 3 + m); }
        const double v0q = c.v0[q], offe = c.p0[q] + h * (double)(k + 1) * v0q;
        v = vv[e]; m = (v > vl - v0q) ? 2 : ((v < -vl - v0q) ? 1 : 0);
        if (m) { n += 1.0; sidsum += (double)((4 * e + 2) * 3 + m); }
        const int a2 = q & 1;
        v = vp[e]; m = (v > hi[a2] - offe) ? 2 : ((v < lo[a2] - offe) ? 1 : 0);
        if (m) { n += 1.0; sidsum += (double)((4 * e + 3) * 3 + m); }
      }
    }
    if (with_collisions && c.ncand > 0) {
      const int* coff = c.wi + c.g->L.coff;
      const int* cj = c.wi + c.g->L.c_j;
      const double* lam = c.wd + c.g->L.lam;
      for (int t = tid; t < (K - 1) * N; t += c.nthreads) {
        int k = 1 + t / N, i = t - (k - 1) * N;
        for (int sidx = coff[k * N + i]; sidx < coff[k * N + i + 1]; ++sidx) {
          const int j = cj[sidx];
          if (j > i && lam[((size_t)k * N + i) * N + j] > 0.0) { n += 1.0; sidsum += (double)(12 * QK + (k * N + i) * N + j); }
        }
      }
    }
    red[tid] = n; red[c.rs + tid] = sidsum;
  }
  SCP_SYNC(c);
  *count = reduce_finish(c, 0, 1);
  *idsum = reduce_finish(c, 1, 1);
}

// ------------------------------------------------------------------ ADMM
struct AdmmOut { int iters; int solved; int certified; int infeasible; int polish_attempts; double pri, dua, npri, ndua; };
// a subproblem whose active set keeps cycling stops asking for the polish after pb.polish_max_failed failed attempts and
// ends on the ADMM residual test instead (a few scenarios spent > 50 attempts x 40 rounds: the tail of a batch)

SCP_DEV AdmmOut admm_run(Ctx& c, int with_collisions, int keep_state, double eps_abs, double eps_rel, int maxit) {
  AdmmOut o; o.iters = 0; o.solved = 0; o.certified = 0; o.infeasible = 0; o.polish_attempts = 0; o.pri = o.dua = INFINITY; o.npri = o.ndua = 0.0;
  const int K = c.K;
  double* red = c.sh;
  double* x = c.a_x;
  const double vl = c.g->pb.vel_limit, al = c.g->pb.acc_limit, jl = c.g->pb.jerk_limit;
  const double lo[2] = {c.g->pb.space[0], c.g->pb.space[1]}, hi[2] = {c.g->pb.space[2], c.g->pb.space[3]};
  double *vj = c.a_vj, *va = c.a_va, *vv = c.a_vv, *vp = c.a_vp;
  double *posrow = c.wd + c.g->L.posrow, *velrow = c.wd + c.g->L.velrow, *off = c.wd + c.g->L.off;
  const double ih = 1.0 / c.g->pb.time_step;
  if (!keep_state) forward_rows(c, 0);
  const int check = c.g->pb.check_every;
  double prev_sc = -1.0, prev_ss = -1.0, fail_sc = -2.0, fail_ss = -2.0;
  int it_mark = 0; double pri_mark = INFINITY;
  // polish_first: "iteration 0" is a polish attempt on the active set the state already implies (previous
  // subproblem's multipliers with warm duals, nothing otherwise) before iterating at all.  It shares the ONE polish
  // call site below, so that the polish (and everything it inlines) exists once in the kernel.
  const int pre_polish = (c.g->pb.polish && c.g->pb.polish_first) ? 1 : 0;
  for (int it = pre_polish ? 0 : 1; it <= maxit; ++it) {
    int want_polish = 0, chk = 0;          // rounds of the polish attempt this pass makes (0: none); residual check
#ifdef SCP_PROFILE_SPLIT     // diagnostic build (make PROFILE_SPLIT=1): rel_step[29..31] of the record := cycles in fused
    long long ts0 = 0;       // iterations, in their collision rows, in check iterations
#endif
    double sc = 0.0, ss = 0.0, pri = 0.0, npri = 0.0, dua = 0.0, ndua = 0.0;
    if (it == 0) want_polish = c.g->pb.polish_first;
    else {
      chk = (it % check == 0) || it == maxit;
#ifdef SCP_PROFILE_SPLIT
      ts0 = SCP_CLOCK();
#endif
#ifndef SCP_EMU
      if (!chk && c.fused_epl > 0) {
        if (c.g->pb.team_mode == 3 || !c.mma_ok) admm_iter_fused<2>(c, c.fused_rows);   // warp-fused iteration (also: A/B timing)
        else if (c.all_hot) admm_iter_mma<true>(c);
        else admm_iter_mma<false>(c);
        __syncthreads();
      } else
#endif
      {
        transpose_rows(c, 0);
        x_update(c, chk);
        forward_rows(c, 1, chk);
      }
#ifdef SCP_PROFILE_SPLIT
      const long long ts1 = SCP_CLOCK();
      if (!chk) c.t_fused += ts1 - ts0;
#endif
      double pri_col = 0.0;
      if (with_collisions && c.ncand > 0) {
#ifndef SCP_EMU
        if (!chk && c.fused_epl > 0 && c.g->pb.team_mode != 3) {
          if (c.all_hot) collision_rows_fast<true>(c); else collision_rows_fast<false>(c);
        } else
#endif
        {
          collision_rows(c, chk);
          if (chk) pri_col = reduce_finish(c, 0, 0);
        }
      }
#ifdef SCP_PROFILE_SPLIT
      if (!chk) c.t_colx += SCP_CLOCK() - ts1;
#endif
      o.iters = it;
      if (chk) {
        // ---- residuals in reference units (OSQP termination test, unscaled)
        SCP_PHASE(c) {
          double pr = 0.0, nr = 0.0;
          for (int e = tid; e < c.Q * K; e += c.nthreads) {
            int q = e / K, k = e - q * K;
            double ax = x[e];
            pr = SCP_FMAX(pr, fabs(ax - clampd(va[e], -al, al)));
            nr = SCP_FMAX(nr, fabs(ax));
            if (k < K - 1) {
              double aj = (x[e + 1] - ax) * ih;
              pr = SCP_FMAX(pr, fabs(aj - clampd(vj[e], -jl, jl)));
              double v0q = c.v0[q];
              pr = SCP_FMAX(pr, fabs(velrow[e] - clampd(vv[e], -vl - v0q, vl - v0q)));
              int a2 = q & 1;
              pr = SCP_FMAX(pr, fabs(posrow[e] - clampd(vp[e], lo[a2] - off[e], hi[a2] - off[e])));
              nr = SCP_FMAX(nr, SCP_FMAX(fabs(aj), SCP_FMAX(fabs(velrow[e]), fabs(posrow[e]))));
            }
          }
          red[tid] = pr; red[c.rs + tid] = nr;
        }
        SCP_SYNC(c);
        pri = reduce_finish(c, 0, 0);
        npri = reduce_finish(c, 1, 0);
        pri = SCP_FMAX(pri, pri_col);
        transpose_rows(c, 1);      // rhs <- 2x + A'y + C'mu
        const double* dres = c.a_rhs;
        SCP_PHASE(c) {
          double du = 0.0, nd = 0.0;
          for (int e = tid; e < c.Q * K; e += c.nthreads) {
            du = SCP_FMAX(du, fabs(dres[e]));
            nd = SCP_FMAX(nd, SCP_FMAX(fabs(2.0 * x[e]), fabs(dres[e] - 2.0 * x[e])));
          }
          red[tid] = du; red[c.rs + tid] = nd;
        }
        SCP_SYNC(c);
        dua = reduce_finish(c, 0, 0);
        ndua = reduce_finish(c, 1, 0);
        o.pri = pri; o.dua = dua; o.npri = npri; o.ndua = ndua;
        if (pri <= eps_abs + eps_rel * npri && dua <= eps_abs + eps_rel * ndua) { o.solved = 1; break; }
        if (!(pri == pri) || !(dua == dua)) break;   // NaN guard
        if (it >= 4 * check && primal_infeasible(c, with_collisions)) { o.infeasible = 1; break; }
        // stalled: the primal residual is still far from feasibility and has not dropped by 20 % over the last
        // `stall_window` iterations -- the signature of an infeasible subproblem long before the certificate's
        // direction test converges.  The iterate is kept, as for an iteration-limit exit (scp.py:446-449).
        if (c.g->pb.stall_window > 0 && it - it_mark >= c.g->pb.stall_window) {
          if (pri > 1e-3 * (1.0 + npri) && pri > 0.8 * pri_mark) { o.infeasible = 2; break; }
          it_mark = it; pri_mark = pri;
        }
        if (c.g->pb.polish) {
          // polish when the active set has not changed between two consecutive checks (and differs from the
          // last set that failed), once the iterate is inside a loose residual gate
          active_signature(c, with_collisions, &sc, &ss);
          const double gate = c.g->pb.polish_first_eps;
          const int settled = (sc == prev_sc && ss == prev_ss) && !(sc == fail_sc && ss == fail_ss);
          prev_sc = sc; prev_ss = ss;
          if (settled && o.polish_attempts < c.pol_max_failed &&
              pri <= gate * (1.0 + npri) && dua <= gate * (1.0 + ndua))
            want_polish = c.g->pb.polish_rounds;
        }
      }
    }
#ifdef SCP_PROFILE_SPLIT
    if (chk) c.t_chk += SCP_CLOCK() - ts0;
#endif
    if (want_polish) {
      const long long t0 = SCP_CLOCK();
      polish_park(c, 0);
      const int pol = polish(c, with_collisions, want_polish);
      if (c.pol_valid && c.pol_A != c.wd + c.g->L.pL) {      // keep Ainv for the next attempt on this candidate set
        const long long total = (long long)sp_row(c.pol_n);
        double* dst = c.wd + c.g->L.pL;
        const double* src = c.pol_A;
        SCP_PHASE(c) { for (long long e = tid; e < total; e += c.nthreads) dst[e] = src[e]; }
        SCP_SYNC(c);
      }
      polish_park(c, 1);
      c.t_polish += SCP_CLOCK() - t0;
      o.polish_attempts++;
      if (pol) { o.solved = 1; o.certified = 1; o.pri = o.dua = 0.0; break; }
      if (it > 0) { fail_sc = sc; fail_ss = ss; }
    }
    if (chk && c.g->pb.adapt_every > 0 && it % c.g->pb.adapt_every == 0 && it < maxit) {
      double est = sqrt((pri / SCP_FMAX(npri, 1e-12)) / SCP_FMAX(dua / SCP_FMAX(ndua, 1e-12), 1e-12));
      if (est > 5.0 || est < 0.2) {
        est = clampd(est, 1e-2, 1e2);
        double nrho = clampd(c.rho * est, 1e-6, 1e6);
        est = nrho / c.rho;
        // keep y: v = z + y/rho  ->  v = z + (v - z)/est
        SCP_PHASE(c) {
          for (int e = tid; e < c.Q * K; e += c.nthreads) {
            int q = e / K, k = e - q * K;
            double v = va[e], z = clampd(v, -al, al); va[e] = z + (v - z) / est;
            if (k < K - 1) {
              v = vj[e]; z = clampd(v, -jl, jl); vj[e] = z + (v - z) / est;
              double v0q = c.v0[q];
              v = vv[e]; z = clampd(v, -vl - v0q, vl - v0q); vv[e] = z + (v - z) / est;
              int a2 = q & 1;
              v = vp[e]; z = clampd(v, lo[a2] - off[e], hi[a2] - off[e]); vp[e] = z + (v - z) / est;
            }
          }
        }
        SCP_SYNC(c);
        c.rho = nrho;
        factor_operator(c);
      }
    }
  }
  return o;
}


// One subproblem: a single ADMM run to the final tolerance; the polish is attempted from inside the
// run whenever the active set has settled, and ends it with the exact minimiser when it certifies.
SCP_DEV AdmmOut solve_qp(Ctx& c, int with_collisions, int keep_state, int cap_hits = 0) {
  const long long t0 = SCP_CLOCK();
  const long long p0 = c.t_polish;
  // every subproblem of this scenario that ran into the iteration cap halves the cap of the next ones (floor 500):
  // from its first unsolved subproblem on the scenario's iterates are solver dependent anyway (scp.py:446-449)
  int cap = c.g->pb.max_admm_iter;
  if (!with_collisions && c.g->pb.max_admm_iter_qp0 > 0) cap = c.g->pb.max_admm_iter_qp0;   // OSQP default max_iter, scp.py:360
  if (c.g->pb.cap_halving) {
    cap >>= (cap_hits < 3 ? cap_hits : 3);
    if (cap < 500) cap = c.g->pb.max_admm_iter < 500 ? c.g->pb.max_admm_iter : 500;
  }
  AdmmOut a = admm_run(c, with_collisions, keep_state, c.g->pb.eps_abs, c.g->pb.eps_rel, cap);
  c.t_admm += (SCP_CLOCK() - t0) - (c.t_polish - p0);
  return a;
}

// ------------------------------------------------------------------ outputs
// scp.py:168-175: positions/velocities for k = 0..K-1 (state k), reference layout (N,K,2).
SCP_DEV void write_outputs(Ctx& c) {
  const int K = c.K;
  const double* x = c.a_x;
  const double* P = c.a_P;
  const double* velrow = c.wd + c.g->L.velrow;
  SCP_PHASE(c) {
    for (int e = tid; e < c.Q * K; e += c.nthreads) {
      int q = e / K, k = e - q * K, i = q >> 1, ax = q & 1;
      size_t o = ((size_t)i * K + k) * 2 + ax;
      c.acc[o] = x[e];
      c.pos[o] = P[e];
      c.vel[o] = (k == 0) ? c.v0[q] : c.v0[q] + velrow[e - 1];
    }
  }
  SCP_SYNC(c);
}

// ------------------------------------------------------------------ the SCP loop
// resumable = 0: the whole loop scp.py:131-180 in one call.
// resumable = 1 (later quanta) / 2 (first quantum): one quantum -- the initial QP + gate + first SCP iteration on the first call, one SCP iteration on
// every later call; between quanta the scenario lives in its record (stage in `reserved2`) and in its `acc` output.
// A suspended subproblem restarts from its accelerations only, which is all the reference carries between
// iterations too (scp.py:155-166).  Returns 1 when the scenario is finished.
SCP_DEV int solve_scenario(Ctx& c, int resumable) {
  const int K = c.K, N = c.N;
  double* red = c.sh;
  scp_b200_record r;
  int stage = 0;
  if (resumable == 1) { r = *c.rec; stage = r.reserved2; }    // resumable == 2: first quantum, the record is not initialised yet
  if (stage != 1) {
    stage = 0;
    r.status = SCP_B200_STATUS_OK; r.scp_iterations = 0; r.converged = 0; r.initial_feasible = 0;
    r.admm_iterations = 0; r.qp_unsolved = 0; r.qp_infeasible = 0; r.polish_attempts = 0;
    r.cycles_total = r.cycles_admm = r.cycles_polish = 0; r.polish_rounds = 0; r.reserved2 = 0;
    r.cycles_pbuild = r.cycles_psolve = r.cycles_peval = r.cycles_papply = 0; r.rebuilds = 0; r.max_copies = 0; r.polish_ok = 0;
    r.first_violation[0] = r.first_violation[1] = r.first_violation[2] = -1;
    r.first_violation_dist = 0; r.min_separation = INFINITY; r.objective = 0; r.pri_res = r.dua_res = 0;
    r.cand_row_iters = 0; r.device_ns = 0;
    for (int e = 0; e < SCP_B200_MAX_SCP_ITER; ++e) r.rel_step[e] = 0.0;
  }

  setup_scenario(c);
  c.rho = c.g->pb.rho0; c.copies = 0; c.ncand = 0; c.t_admm = 0; c.t_polish = 0; c.polish_rounds = 0; c.pol_valid = 0; c.pol_n = 0; c.pol_use_col = 0; c.pol_col_stale = 0; c.pol_no_smem = 0; c.pol_A = nullptr;
  c.t_pbuild = c.t_psolve = c.t_peval = c.t_papply = 0; c.t_fused = c.t_colx = c.t_chk = 0;
  const long long t_begin = SCP_CLOCK(), ns_begin = SCP_NANOS();
  double minsep; long long frow; double fdist;
  int feasible = 0, it = 0, converged = 0;
  double* x = c.a_x;
  double* xprev = c.wd + c.g->L.xprev;
  double* P = c.a_P;
  double* Pb = c.wd + c.g->L.Pbar;
  int qp0 = (stage == 0);                 // the first pass of the loop below is the initial QP (scp.py:138)
  if (!qp0) {
    // resume: accelerations of the last iterate -> x, positions
    SCP_PHASE(c) {
      for (int e = tid; e < c.Q * K; e += c.nthreads) {
        int q = e / K, k = e - q * K;
        x[e] = c.acc[((size_t)(q >> 1) * K + k) * 2 + (q & 1)];
      }
    }
    SCP_SYNC(c);
    forward_rows(c, 0);
    it = r.scp_iterations;
  }
  // One loop for QP #0 and the avoidance QPs, so that solve_qp -> admm_run -> polish is inlined at ONE call site
  // (the fully inlined kernel with two sites was 88 k instructions and took 7 minutes to compile).
  for (;;) {
    int warm = 0;
    if (!qp0) {
      if (!(r.status != SCP_B200_STATUS_INITIAL_QP_FAILED && it < c.g->pb.max_scp_iter && !converged && !feasible)) break;
      SCP_PHASE(c) {
        for (int e = tid; e < c.Q * K; e += c.nthreads) { Pb[e] = P[e]; xprev[e] = x[e]; }
      }
      SCP_SYNC(c);
      warm = c.g->pb.warm_duals && it > 0;               // OSQP restarts y = 0 every SCP iteration (scp.py:441-443);
      mark_near_rows(c, c.g->pb.cand_margin, warm);      // keeping the duals changes the path, not the minimiser
      if (!warm) c.rho = c.g->pb.rho0;
    }
    AdmmOut a; a.iters = 0; a.solved = 0; a.certified = 0; a.infeasible = 0; a.polish_attempts = 0; a.pri = a.dua = 0;
    int have_state = 0;
    for (int attempt = 0;; ++attempt) {
      int need_factor = 1;
      if (!qp0) {
        const int old_copies = c.copies;
        build_candidates(c);
        if (c.copies > r.max_copies) r.max_copies = c.copies;
        need_factor = !have_state || c.copies != old_copies;
      }
      if (need_factor) factor_operator(c);
      // a scenario that already has unsolved subproblems is solver dependent from there on (scp.py:446-449) and its later
      // subproblems are usually infeasible as well: each of them gets half the failed-polish budget of the one before
      // (a failed attempt costs as much as ~170 ADMM iterations; one scenario of the 8192 of the 8-GPU bench spent 1.7 s there)
      c.pol_max_failed = c.g->pb.polish_max_failed;
      if (c.g->pb.cap_halving && !qp0) {
        c.pol_max_failed >>= (r.qp_unsolved < 3 ? r.qp_unsolved : 3);
        if (c.pol_max_failed < 1) c.pol_max_failed = 1;
      }
      a = solve_qp(c, !qp0, !qp0 && (have_state || warm), qp0 ? 0 : r.qp_unsolved - r.qp_infeasible);   // QP #0 scp.py:138, QP #t scp.py:155
      have_state = 1;
      r.admm_iterations += a.iters;
      if (qp0) break;
      r.cand_row_iters += 0.5 * (double)c.ncand * (double)a.iters;
      if (!a.solved) break;                            // unsolved (e.g. infeasible) subproblem: keep x, like scp.py:446-449
      int bad = (c.ncand < N * (N - 1) * (K - 1)) ? verify_rows(c, c.g->pb.verify_tol) : 0;
      if (bad == 0 || attempt >= 20) break;
      r.rebuilds++;
    }
    r.polish_ok += a.certified;
    r.polish_attempts += a.polish_attempts;
    r.pri_res = a.pri; r.dua_res = a.dua;
    if (qp0) {
      // scp.py:363-365 raises unless OSQP reports "solved" or "solved inaccurate" (status_val 1 or 2): an initial QP that
      // ends at its iteration cap within 10x of OSQP's default tolerances (eps 1e-3) is accepted like status 2
      const int inaccurate_ok = !a.infeasible && a.pri <= 10.0 * (1e-3 + 1e-3 * a.npri) && a.dua <= 10.0 * (1e-3 + 1e-3 * a.ndua);
      if (!a.solved) { r.qp_unsolved++; if (!inaccurate_ok) r.status = SCP_B200_STATUS_INITIAL_QP_FAILED; }
      forward_rows(c, 0);                                 // positions of the initial guess, scp.py:140
      gate_and_minsep(c, &minsep, &frow, &fdist);         // scp.py:144
      feasible = frow < 0;
      r.initial_feasible = feasible;
      if (!feasible) {
        long long npairs = (long long)N * (N - 1) / 2;
        int k = (int)(frow / npairs); long long p = frow - (long long)k * npairs;
        int i = 0; while (p >= N - 1 - i) { p -= N - 1 - i; ++i; }
        r.first_violation[0] = k; r.first_violation[1] = i; r.first_violation[2] = i + 1 + (int)p;
        r.first_violation_dist = fdist;
        if (k == 0 && r.status == SCP_B200_STATUS_OK) r.status = SCP_B200_STATUS_START_TOO_CLOSE;
      }
      qp0 = 0;
      continue;
    }
    if (!a.solved) r.qp_unsolved++;
    r.qp_infeasible += (a.infeasible != 0);
    // rel step on accelerations, scp.py:157-163
    SCP_PHASE(c) {
      double s0 = 0, s1 = 0;
      for (int e = tid; e < c.Q * K; e += c.nthreads) { double d = x[e] - xprev[e]; s0 += d * d; s1 += xprev[e] * xprev[e]; }
      red[tid] = s0; red[c.rs + tid] = s1;
    }
    SCP_SYNC(c);
    double dn = reduce_finish(c, 0, 1), pn = reduce_finish(c, 1, 1);
    double rel = sqrt(dn) / sqrt(pn);
    if (it < SCP_B200_MAX_SCP_ITER) r.rel_step[it] = rel;
    if (rel <= c.g->pb.scp_tolerance) converged = 1;
    ++it;
    if (resumable) break;
  }
  r.scp_iterations = it; r.converged = converged;
  r.cycles_total += SCP_CLOCK() - t_begin; r.cycles_admm += c.t_admm; r.cycles_polish += c.t_polish; r.polish_rounds += c.polish_rounds;
  r.cycles_pbuild += c.t_pbuild; r.cycles_psolve += c.t_psolve; r.cycles_peval += c.t_peval; r.cycles_papply += c.t_papply;
  r.device_ns += SCP_NANOS() - ns_begin;
#ifdef SCP_PROFILE_SPLIT
  r.rel_step[29] += (double)c.t_fused; r.rel_step[30] += (double)c.t_colx; r.rel_step[31] += (double)c.t_chk;
#endif
  if (resumable && r.status != SCP_B200_STATUS_INITIAL_QP_FAILED && it < c.g->pb.max_scp_iter && !converged && !feasible) {
    // suspend: the iterate goes to the scenario's acc output, the counters to its record
    r.reserved2 = 1;
    SCP_PHASE(c) {
      for (int e = tid; e < c.Q * K; e += c.nthreads) {
        int q = e / K, k = e - q * K;
        c.acc[((size_t)(q >> 1) * K + k) * 2 + (q & 1)] = x[e];
      }
      if (tid == 0) *c.rec = r;
    }
    SCP_SYNC(c);
    return 0;
  }
  r.reserved2 = 2;
  forward_rows(c, 0);
  gate_and_minsep(c, &minsep, &frow, &fdist);
  r.min_separation = minsep;
  SCP_PHASE(c) {
    double s = 0;
    for (int e = tid; e < c.Q * K; e += c.nthreads) s += x[e] * x[e];
    red[tid] = s;
  }
  SCP_SYNC(c);
  r.objective = reduce_finish(c, 0, 1);
  write_outputs(c);
  SCP_PHASE(c) { if (tid == 0) *c.rec = r; }
  SCP_SYNC(c);
  return 1;
}

}  // namespace scp
#endif
