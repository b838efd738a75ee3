// scp_stream.cu -- streaming SCP solver (sm_100a): the loop of scp.py:131-180 as a sequence of
// HBM/L2-streaming kernels over ALL scenarios and agents of a batch, with the per-scenario control
// flow (ADMM termination, rho adaptation, candidate verification, SCP convergence) decided on the
// device by small control kernels, so the host only replays a fixed launch sequence and polls one
// counter.  It is the path for scenarios too large for one CTA's shared memory (config 3: 200 agents,
// config 5: batches of 50-200 agents) and for ONE scenario whose agents are sharded over the GPUs of a
// node (config 4): rank g owns a block of agents, every ADMM iteration ends with an NCCL all-gather of
// the positions (N*K*2 doubles), the residual scalars are all-gathered at check iterations.
//
// Same QP splitting as scp_device.inl: box rows in the v = z + y/rho form,
// terminal equalities exact in the x-update, collision rows as contact-force iterations on per-agent
// copies, one K x K operator per scenario.  No polish here: subproblems end on the ADMM residual test.
//
// Reference code replaced (src/path_planning/solvers/scp.py):
//   k_iter          one ADMM iteration, fused:        OSQP iteration inside problem.solve() :362,:445 with the rows of
//                   collision rows + x-update + A x   _add_collision_constraints :453-557 (matrix free)
//   k_build         candidate rows + linearisation    _add_collision_constraints           :487-549
//   k_scan          gate / min separation / verify    _fast_check_avoidance_constraints    :597-615
//   k_control*      termination + SCP loop            generate_trajectories                :152-166
//   k_output        result dict                       :168-175

#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: every NCCL function is resolved with dlsym at run time

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/scp_b200.h"
#include "scp_defaults.h"
#include "scp_tables.h"

extern int scp_b200_set_error(int code, const char* msg);   // scp_b200.cu

namespace ss {

constexpr int NRED = 24;
enum { R_PRI = 0, R_NPRI, R_DUA, R_NDUA, R_PRICOL, R_DN, R_PN, R_OBJ, R_MINSEP, R_FIRST, R_BAD, R_MAXD, R_COPIES, R_NCAND, R_OVER, R_SPARE,
       R_VIOLJ, R_VIOLA, R_VIOLV, R_VIOLP };   // box-row classes outside the ADMM: max violation at the iterate
enum { FL_SCAN = 1, FL_GATE = 2, FL_VERIFY = 4, FL_SNAPSHOT = 8, FL_BUILD = 16, FL_RESCALE = 64, FL_RESET = 128,
       FL_FACTOR = 256, FL_FINISH = 512 };
constexpr int MAXC_MAX = 48;
constexpr int AX_THREADS = 256;

struct State {
  int phase;        // 0: initial QP, 1: QP with collision rows, 2: finished
  int flags;        // work requested from the predicated kernels of this macro step
  int qp_it, scp_it, copies, attempt, have_state, qp_solved, stalled, it_mark;
  int on_mask, reset_mask;   // box-row classes carried by the ADMM (1 jerk, 2 acc, 4 vel, 8 pos); classes whose v := A x on FL_RESET
  double rho, est, margin, pri_mark, ncand;
  double pri, dua, dn, pn, obj, minsep;
};

struct Dev {
  scp_b200_problem pb;
  int B, N, K, Q, Qs;          // Qs = row stride of one scenario in the per-agent-axis arrays (>= Q, padded for the all-gather)
  int a_lo, a_hi;              // agents owned by this rank
  int maxc, G, rank;
  const double *rj, *ra, *rv, *rp, *rc, *B2;
  const double* Bc[4];         // K x K unit-rho operators of the box-row classes: D'RjD, Ra, V'RvV, S'RpS
  double *x, *xprev, *va, *vj, *vv, *vp, *P, *P1, *Pbar, *F, *FY, *mu, *qsum;   // P / P1: position ping-pong (P is current between check periods)
  double *Nmat, *N0, *Qm, *gg;
  const double *p0, *v0, *pf, *vf;
  int *cnt, *cj;
  double *cex, *cey, *cb, *lam, *lam1, *lamt;   // lam / lam1: multiplier ping-pong (lam is current between check periods)
  double *slab, *gath;
  State* st;
  scp_b200_record* rec;
  double *acc, *pos, *vel;     // (B, Npad, K, 2)
  int Npad;
  int* done;
  // peer-memory exchange of the position slices (agent sharding without a per-iteration NCCL call)
  int p2p;
  double* peerP[8];            // P of every rank (own entry = local pointer)
  double* peerP1[8];
  unsigned* peerFlag[8];       // flag arrays of every rank: flag[g] = k_iter launches rank g has completed and published
  unsigned *flag, *iter_no, *cta_done, *err;
  long long* t0ns;             // %globaltimer at the start of the solve (k_init)
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
__device__ __forceinline__ long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
__device__ __forceinline__ void atomic_max_pos(double* addr, double v) {
  atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_min_pos(double* addr, double v) {
  atomicMin((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ double warp_max_nan(double v) {   // NaN-propagating max of non-negative values (bit order)
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { unsigned long long o = __shfl_xor_sync(0xffffffffu, b, d); b = o > b ? o : b; }
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

template <int EPL>
__device__ __forceinline__ void suffix_sum(double (&v)[EPL], int lane) {
  double carry = 0.0;
#pragma unroll
  for (int e = EPL - 1; e >= 0; --e) {
    double s = v[e];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { double o = __shfl_down_sync(0xffffffffu, s, d); if (lane + d < 32) s += o; }
    s += carry; v[e] = s;
    carry = __shfl_sync(0xffffffffu, s, 0);
  }
}
template <int EPL>
__device__ __forceinline__ void prefix_sum(double (&v)[EPL], int lane) {
  double carry = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    double s = v[e];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { double o = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += o; }
    s += carry; v[e] = s;
    carry = __shfl_sync(0xffffffffu, s, 31);
  }
}

// ---------------------------------------------------------------------------------- init
__global__ void k_init(const __grid_constant__ Dev d) {
  const int b = blockIdx.y, K = d.K;
  const double h = d.pb.time_step;
  const size_t base = (size_t)b * d.Qs * K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.Qs * K; e += gridDim.x * blockDim.x) {
    const int q = e / K, k = e - q * K;
    double p = 0.0;
    if (q < d.Q) { const double p0 = d.p0[(size_t)b * d.Q + q], v0 = d.v0[(size_t)b * d.Q + q]; p = p0 + h * (double)k * v0; }
    d.x[base + e] = 0.0; d.xprev[base + e] = 0.0; d.va[base + e] = 0.0; d.vj[base + e] = 0.0; d.vv[base + e] = 0.0; d.vp[base + e] = 0.0;
    d.P[base + e] = p; d.P1[base + e] = p; d.Pbar[base + e] = p; d.F[base + e] = 0.0; d.FY[base + e] = 0.0;
  }
  const int Nown = d.a_hi - d.a_lo;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < Nown * K; t += gridDim.x * blockDim.x) d.cnt[(size_t)b * Nown * K + t] = 0;
  if (blockIdx.x == 0) {
    for (int t = threadIdx.x; t < 2 * d.Qs; t += blockDim.x) d.mu[(size_t)b * 2 * d.Qs + t] = 0.0;
    for (int t = threadIdx.x; t < NRED; t += blockDim.x)
      d.slab[(size_t)b * NRED + t] = (t == R_MINSEP || t == R_FIRST) ? INFINITY : 0.0;
    if (threadIdx.x == 0) {
      State s;
      s.phase = 0; s.flags = FL_FACTOR; s.qp_it = 0; s.scp_it = 0; s.copies = 0; s.attempt = 0; s.have_state = 0; s.qp_solved = 0;
      s.stalled = 0; s.it_mark = 0; s.on_mask = d.pb.lazy_rows ? 0 : 15; s.reset_mask = 0; s.rho = d.pb.rho0; s.est = 1.0; s.margin = d.pb.cand_margin; s.pri_mark = INFINITY; s.ncand = 0.0;
      s.pri = s.dua = INFINITY; s.dn = s.pn = s.obj = 0.0; s.minsep = INFINITY;
      d.st[b] = s;
      scp_b200_record r;
      memset(&r, 0, sizeof(r));
      r.status = SCP_B200_STATUS_OK;
      r.first_violation[0] = r.first_violation[1] = r.first_violation[2] = -1;
      r.min_separation = INFINITY;
      d.rec[b] = r;
      if (b == 0) { *d.done = 0; *d.t0ns = globaltimer_ns(); }
    }
  }
}

// ---------------------------------------------------------------------------------- operator
// One CTA per scenario that asked for it: M(rho, copies) inverted in shared memory by Gauss-Jordan (SPD, no
// pivoting), then the equality-constrained solution operator (scp_device.inl factor_operator):
//   x = Nmat r + N0 d,  mu = Qm r - G d.
__global__ void __launch_bounds__(512) k_factor(const __grid_constant__ Dev d) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, K = d.K, tid = threadIdx.x, nt = blockDim.x;
  const State& S = d.st[b];
  if (!(S.flags & FL_FACTOR)) return;
  double* M = sm;
  double* colb = sm + (size_t)K * K;
  double* rowb = colb + K;
  double* mc = rowb + K;            // 2K
  __shared__ double G3[3];
  const double rho = S.rho, sig = d.pb.sigma, cp = (double)S.copies, h = d.pb.time_step;
  const int on = S.on_mask;
  for (int e = tid; e < K * K; e += nt) {
    const int r = e / K, c = e - r * K;
    double v = cp * d.B2[e];
    if (on & 1) v += d.Bc[0][e];
    if (on & 2) v += d.Bc[1][e];
    if (on & 4) v += d.Bc[2][e];
    if (on & 8) v += d.Bc[3][e];
    M[e] = rho * v + (r == c ? 2.0 + sig : 0.0);
  }
  __syncthreads();
  for (int p = 0; p < K; ++p) {
    for (int e = tid; e < K; e += nt) { colb[e] = M[e * K + p]; rowb[e] = M[p * K + e]; }
    __syncthreads();
    const double ip = 1.0 / colb[p];
    for (int e = tid; e < K * K; e += nt) {
      const int r = e / K, c = e - r * K;
      double v;
      if (r == p) v = (c == p) ? ip : rowb[c] * ip;
      else if (c == p) v = -colb[r] * ip;
      else v = M[e] - colb[r] * rowb[c] * ip;
      M[e] = v;
    }
    __syncthreads();
  }
  for (int k = tid; k < K; k += nt) {
    double a0 = 0.0, a1 = 0.0;
    for (int j = 0; j < K; ++j) { const double m = M[k * K + j]; a0 += m * h; a1 += m * (h * h * ((double)(K - 1 - j) + 0.5)); }
    mc[2 * k] = a0; mc[2 * k + 1] = a1;
  }
  __syncthreads();
  if (tid == 0) {
    double h00 = 0, h01 = 0, h11 = 0;
    for (int k = 0; k < K; ++k) {
      const double cv = h, cpk = h * h * ((double)(K - 1 - k) + 0.5);
      h00 += cv * mc[2 * k]; h01 += cv * mc[2 * k + 1]; h11 += cpk * mc[2 * k + 1];
    }
    const double det = h00 * h11 - h01 * h01;
    G3[0] = h11 / det; G3[1] = -h01 / det; G3[2] = h00 / det;
    d.gg[(size_t)b * 4 + 0] = G3[0]; d.gg[(size_t)b * 4 + 1] = G3[1]; d.gg[(size_t)b * 4 + 2] = G3[2];
  }
  __syncthreads();
  const double g00 = G3[0], g01 = G3[1], g11 = G3[2];
  double* N0 = d.N0 + (size_t)b * 2 * K;
  double* Qm = d.Qm + (size_t)b * 2 * K;
  for (int k = tid; k < K; k += nt) {
    const double n0 = mc[2 * k] * g00 + mc[2 * k + 1] * g01, n1 = mc[2 * k] * g01 + mc[2 * k + 1] * g11;
    N0[2 * k] = n0; N0[2 * k + 1] = n1;
    Qm[k] = n0; Qm[K + k] = n1;
    colb[k] = n0; rowb[k] = n1;
  }
  __syncthreads();
  double* out = d.Nmat + (size_t)b * K * K;
  for (int e = tid; e < K * K; e += nt) {
    const int r = e / K, c = e - r * K;
    out[e] = M[e] - (mc[2 * r] * colb[c] + mc[2 * r + 1] * rowb[c]);
  }
}

// ---------------------------------------------------------------------------------- agent-axis kernel
// (MODE 0/1 are the agent-axis half of the first, two-kernel form of the iteration, kept for reference and for
// debugging against k_iter; only MODE 2 is launched.)
// One warp per (scenario, agent-axis): lanes hold the steps k = lane + 32 e.  MODE 0: one ADMM iteration of the
// box rows and the x-update (scp_device.inl admm_iter_fused, state streamed from HBM/L2 instead of shared memory);
// MODE 1: the same plus the primal residual terms and the equality multipliers (check iteration);
// MODE 2: v := A x (multipliers reset at the start of a subproblem: OSQP warm start with x only, scp.py:443).
template <int EPL, int MODE>
__global__ void __launch_bounds__(AX_THREADS) k_axis(const __grid_constant__ Dev d, int qpc) {
  extern __shared__ double sm[];
  const int b = blockIdx.y;
  const State& S = d.st[b];
  if (MODE <= 1) { if (S.phase >= 2) return; } else { if (!(S.flags & FL_RESET)) return; }
  const int K = d.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int q_hi = 2 * d.a_hi;
  const int qa = 2 * d.a_lo + blockIdx.x * qpc, qb = min(qa + qpc, q_hi);
  if (qa >= qb) return;
  double* Nm = sm;
  double* myrhs = sm + (MODE <= 1 ? (size_t)K * K : 0) + (size_t)warp * K;
  if (MODE <= 1) {
    const double* src = d.Nmat + (size_t)b * K * K;
    for (int e = threadIdx.x; e < K * K; e += blockDim.x) Nm[e] = src[e];
    __syncthreads();
  }
  const double h = d.pb.time_step, ih = 1.0 / h, rho = S.rho, sig = d.pb.sigma;
  const double vl = d.pb.vel_limit, al = d.pb.acc_limit, jl = d.pb.jerk_limit;
  const double cpr = (double)S.copies * rho;
  // over-relaxation of the box rows (OSQP's alpha): v' = alpha A x' + (1 - alpha) z + (v - z)
  const double alpha = (MODE <= 1 && d.pb.relax_pct > 0) ? 0.01 * (double)d.pb.relax_pct : 1.0;
  const int on = (MODE <= 1) ? S.on_mask : 0;
  const double* N0 = d.N0 + (size_t)b * 2 * K;
  const double* Qm = d.Qm + (size_t)b * 2 * K;
  double trj[EPL], tra[EPL], trv[EPL], trp[EPL], trc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int k = lane + 32 * e;
    const bool in = k < K;
    trj[e] = (in && (on & 1)) ? rho * d.rj[k] : 0.0; tra[e] = (in && (on & 2)) ? rho * d.ra[k] : 0.0;
    trv[e] = (in && (on & 4)) ? rho * d.rv[k] : 0.0; trp[e] = (in && (on & 8)) ? rho * d.rp[k] : 0.0;
    trc[e] = in ? cpr * d.rc[k] : 0.0;
  }
  // a class outside the ADMM keeps v = A x (no multiplier): s = 0 and no relaxation
  const double alj = (on & 1) ? alpha : 1.0, ala = (on & 2) ? alpha : 1.0, alv = (on & 4) ? alpha : 1.0, alp = (on & 8) ? alpha : 1.0;
  double pr = 0.0, nr = 0.0, voj = 0.0, voa = 0.0, vov = 0.0, vop = 0.0;
  for (int q = qa + warp; q < qb; q += nw) {
    const size_t row = ((size_t)b * d.Qs + q) * K;
    const size_t q2 = (size_t)b * d.Q + q;
    const double v0q = d.v0[q2], p0q = d.p0[q2];
    const double lv = -vl - v0q, uv = vl - v0q;
    const double plo = d.pb.space[q & 1], phi = d.pb.space[2 + (q & 1)];
    double *x = d.x + row, *vj = d.vj + row, *va = d.va + row, *vv = d.vv + row, *vp = d.vp + row, *P = d.P + row;
    const double* F = d.F + row;
    double xn[EPL], sj[EPL], sa[EPL], sv[EPL], sp[EPL], off[EPL];
    if (MODE <= 1) {
      double xo[EPL], wj[EPL], wv[EPL], wp[EPL], wa[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = lane + 32 * e;
        xo[e] = 0; sj[e] = sa[e] = sv[e] = sp[e] = 0; wj[e] = wv[e] = wp[e] = wa[e] = 0; off[e] = 0;
        if (k < K) {
          xo[e] = x[k];
          double v = va[k], z = clampd(v, -al, al);
          sa[e] = (on & 2) ? v - alpha * z : 0.0; wa[e] = tra[e] * (2 * z - v);
          if (k < K - 1) {
            v = vj[k]; z = clampd(v, -jl, jl); sj[e] = (on & 1) ? v - alpha * z : 0.0; wj[e] = trj[e] * (2 * z - v);
            v = vv[k]; z = clampd(v, lv, uv); sv[e] = (on & 4) ? v - alpha * z : 0.0; wv[e] = trv[e] * (2 * z - v);
            off[e] = p0q + h * (double)(k + 1) * v0q;
            v = vp[k]; z = clampd(v, plo - off[e], phi - off[e]); sp[e] = (on & 8) ? v - alpha * z : 0.0;
            wp[e] = trp[e] * (2 * z - v) + trc[e] * (P[k + 1] - off[e]) + F[k + 1];
          }
        }
      }
      // transpose: D'wj + wa + V'wv + S'wp
      double r1v[EPL], r1p[EPL], r2p[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) { r1v[e] = wv[e]; r1p[e] = wp[e]; }
      suffix_sum<EPL>(r1v, lane);
      suffix_sum<EPL>(r1p, lane);
#pragma unroll
      for (int e = 0; e < EPL; ++e) r2p[e] = r1p[e];
      suffix_sum<EPL>(r2p, lane);
      double last = 0.0;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = lane + 32 * e;
        double prev = __shfl_up_sync(0xffffffffu, wj[e], 1);
        if (lane == 0) prev = last;
        last = __shfl_sync(0xffffffffu, wj[e], 31);
        if (k < K) myrhs[k] = sig * xo[e] + (prev - wj[e]) * ih + wa[e] + h * r1v[e] + h * h * (r2p[e] - 0.5 * r1p[e]);
      }
      __syncwarp();
      // x = Nmat rhs + N0 d   (Nmat symmetric: lane walks its columns, rhs broadcast from the warp's shared row)
      const double d0 = d.vf[q2] - v0q, d1 = d.pf[q2] - (p0q + h * (double)K * v0q);
      int kc[EPL];
      double a0[EPL], a1[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) { const int k = lane + 32 * e; kc[e] = k < K ? k : K - 1; a0[e] = 0.0; a1[e] = 0.0; }
      int j = 0;
      for (; j + 1 < K; j += 2) {
        const double r0 = myrhs[j], r1 = myrhs[j + 1];
        const double* n0 = Nm + (size_t)j * K;
#pragma unroll
        for (int e = 0; e < EPL; ++e) { a0[e] += n0[kc[e]] * r0; a1[e] += n0[K + kc[e]] * r1; }
      }
      if (j < K) {
        const double r0 = myrhs[j];
        const double* n0 = Nm + (size_t)j * K;
#pragma unroll
        for (int e = 0; e < EPL; ++e) a0[e] += n0[kc[e]] * r0;
      }
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = lane + 32 * e;
        xn[e] = (k < K) ? N0[2 * k] * d0 + N0[2 * k + 1] * d1 + (a0[e] + a1[e]) : 0.0;
      }
      if (MODE == 1) {
        // mu = Qm rhs - G d
        double m0 = 0.0, m1 = 0.0;
        for (int jj = lane; jj < K; jj += 32) { const double r = myrhs[jj]; m0 += Qm[jj] * r; m1 += Qm[K + jj] * r; }
        m0 = warp_sum(m0); m1 = warp_sum(m1);
        if (lane == 0) {
          const double* gg = d.gg + (size_t)b * 4;
          d.mu[((size_t)b * d.Qs + q) * 2] = m0 - (gg[0] * d0 + gg[1] * d1);
          d.mu[((size_t)b * d.Qs + q) * 2 + 1] = m1 - (gg[1] * d0 + gg[2] * d1);
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = lane + 32 * e;
        xn[e] = (k < K) ? x[k] : 0.0;
        sj[e] = sa[e] = sv[e] = sp[e] = 0.0;
        off[e] = (k < K - 1) ? p0q + h * (double)(k + 1) * v0q : 0.0;
      }
    }
    // forward rows and v update
    double c1[EPL], c2[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) c1[e] = xn[e];
    prefix_sum<EPL>(c1, lane);
#pragma unroll
    for (int e = 0; e < EPL; ++e) c2[e] = c1[e];
    prefix_sum<EPL>(c2, lane);
    double first_next = 0.0;
#pragma unroll
    for (int e = EPL - 1; e >= 0; --e) {
      const int k = lane + 32 * e;
      double nxt = __shfl_down_sync(0xffffffffu, xn[e], 1);
      if (lane == 31) nxt = first_next;
      first_next = __shfl_sync(0xffffffffu, xn[e], 0);
      if (k < K) {
        if (MODE <= 1) x[k] = xn[e];
        const double nva = ala * xn[e] + sa[e];
        if (MODE <= 1 || (S.reset_mask & 2)) va[k] = nva;
        if (MODE == 1) {
          const double dv = fabs(xn[e] - clampd(nva, -al, al));
          if (on & 2) pr = fmax(pr, dv); else voa = fmax(voa, dv);
          nr = fmax(nr, fabs(xn[e]));
        }
        if (k < K - 1) {
          const double rv_ = h * c1[e], rp_ = h * h * (c2[e] - 0.5 * c1[e]);
          const double aj = (nxt - xn[e]) * ih;
          const double nvj = alj * aj + sj[e], nvv = alv * rv_ + sv[e], nvp = alp * rp_ + sp[e];
          if (MODE <= 1 || (S.reset_mask & 1)) vj[k] = nvj;
          if (MODE <= 1 || (S.reset_mask & 4)) vv[k] = nvv;
          if (MODE <= 1 || (S.reset_mask & 8)) vp[k] = nvp;
          if (MODE <= 1) P[k + 1] = off[e] + rp_;
          if (MODE == 1) {
            const double dj = fabs(aj - clampd(nvj, -jl, jl)), dv = fabs(rv_ - clampd(nvv, lv, uv));
            const double dp = fabs(rp_ - clampd(nvp, plo - off[e], phi - off[e]));
            if (on & 1) pr = fmax(pr, dj); else voj = fmax(voj, dj);
            if (on & 4) pr = fmax(pr, dv); else vov = fmax(vov, dv);
            if (on & 8) pr = fmax(pr, dp); else vop = fmax(vop, dp);
            nr = fmax(nr, fmax(fabs(aj), fmax(fabs(rv_), fabs(rp_))));
          }
        }
      }
    }
  }
  if (MODE == 1) {
    pr = warp_max_nan(fabs(pr)); nr = warp_max_nan(fabs(nr));
    if (lane == 0) { atomic_max_pos(d.slab + (size_t)b * NRED + R_PRI, pr); atomic_max_pos(d.slab + (size_t)b * NRED + R_NPRI, nr); }
    if (on != 15) {
      voj = warp_max(voj); voa = warp_max(voa); vov = warp_max(vov); vop = warp_max(vop);
      if (lane == 0) {
        double* sl = d.slab + (size_t)b * NRED;
        if (voj > 0.0) atomic_max_pos(sl + R_VIOLJ, voj);
        if (voa > 0.0) atomic_max_pos(sl + R_VIOLA, voa);
        if (vov > 0.0) atomic_max_pos(sl + R_VIOLV, vov);
        if (vop > 0.0) atomic_max_pos(sl + R_VIOLP, vop);
      }
    }
  }
}


// ---------------------------------------------------------------------------------- fused ADMM iteration
// ONE kernel per ADMM iteration.  One warp per (scenario, agent); lane L owns the EPL consecutive steps
// k = L*EPL .. L*EPL+EPL-1 ("blocked": a scan over k is EPL local adds + ONE warp scan, rows move as 16-byte vectors).
//   1. collision rows of the agent at its steps (candidate walk, multiplier update, force on p_i[k]) from the
//      positions of the PREVIOUS iterate (buffer `cur`) -- both owners of a row see the same numbers, so their
//      copies of the multiplier stay identical (scp_device.inl collision_rows);
//   2. for each axis: box rows -> right-hand side (transposed scans), x = Nmat rhs + N0 d with the K x K operator
//      staged in shared memory by cp.async while step 1 runs, forward scans -> rows of A x, new positions into
//      buffer `cur ^ 1` (scp_device.inl admm_iter_fused).
// Box-row classes outside the ADMM (lazy rows) cost no memory traffic: only their violation is reduced at checks.
// CHK = 1 adds the primal residual terms, the equality multipliers and sum_j lam eta (FY) for the dual residual.
constexpr int IT_THREADS = 256;   // 8 warps = 4 agents x 2 axes

template <int EPL>
__device__ __forceinline__ void load_row(const double* __restrict__ p, int k0, int K, bool vec, double (&v)[EPL]) {
  if ((EPL & 1) == 0 && vec) {
#pragma unroll
    for (int e = 0; e < EPL; e += 2) {
      double2 t = make_double2(0.0, 0.0);
      if (k0 + e < K) t = *reinterpret_cast<const double2*>(p + k0 + e);
      v[e] = t.x; v[e + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPL; ++e) v[e] = (k0 + e < K) ? p[k0 + e] : 0.0;
  }
}
// no bounds check: the caller guarantees k0 .. k0+EPL-1 are readable (the staged operator is padded by EPL doubles)
template <int EPL>
__device__ __forceinline__ void load_row_nc(const double* __restrict__ p, int k0, bool vec, double (&v)[EPL]) {
  if ((EPL & 1) == 0 && vec) {
#pragma unroll
    for (int e = 0; e < EPL; e += 2) { const double2 t = *reinterpret_cast<const double2*>(p + k0 + e); v[e] = t.x; v[e + 1] = t.y; }
  } else {
#pragma unroll
    for (int e = 0; e < EPL; ++e) v[e] = p[k0 + e];
  }
}
template <int EPL>
__device__ __forceinline__ void store_row(double* __restrict__ p, int k0, int K, bool vec, const double (&v)[EPL]) {
  if ((EPL & 1) == 0 && vec) {
#pragma unroll
    for (int e = 0; e < EPL; e += 2)
      if (k0 + e < K) *reinterpret_cast<double2*>(p + k0 + e) = make_double2(v[e], v[e + 1]);
  } else {
#pragma unroll
    for (int e = 0; e < EPL; ++e) if (k0 + e < K) p[k0 + e] = v[e];
  }
}

// inclusive suffix sums over k of a (first order) and of (b, then once more: c = suffix of suffix of b)
template <int EPL>
__device__ __forceinline__ void suffix3(double (&a)[EPL], double (&b)[EPL], double (&c)[EPL], int lane) {
  double ta = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
  for (int e = EPL - 1; e >= 0; --e) { ta += a[e]; a[e] = ta; t1 += b[e]; b[e] = t1; t2 += t1; c[e] = t2; }
  double ia = ta, i1 = t1, i2 = t2;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const double oa = __shfl_down_sync(0xffffffffu, ia, dd), o1 = __shfl_down_sync(0xffffffffu, i1, dd), o2 = __shfl_down_sync(0xffffffffu, i2, dd);
    if (lane + dd < 32) { ia += oa; i2 += o2 + (double)(dd * EPL) * o1; i1 += o1; }
  }
  double ca = __shfl_down_sync(0xffffffffu, ia, 1), c1 = __shfl_down_sync(0xffffffffu, i1, 1), c2 = __shfl_down_sync(0xffffffffu, i2, 1);
  if (lane == 31) { ca = 0.0; c1 = 0.0; c2 = 0.0; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) { a[e] += ca; c[e] += c2 + (double)(EPL - e) * c1; b[e] += c1; }
}
// inclusive prefix sums: b = prefix of a's values given in b, c = prefix of prefix
template <int EPL>
__device__ __forceinline__ void prefix2(double (&b)[EPL], double (&c)[EPL], int lane) {
  double t1 = 0.0, t2 = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) { t1 += b[e]; b[e] = t1; t2 += t1; c[e] = t2; }
  double i1 = t1, i2 = t2;
#pragma unroll
  for (int dd = 1; dd < 32; dd <<= 1) {
    const double o1 = __shfl_up_sync(0xffffffffu, i1, dd), o2 = __shfl_up_sync(0xffffffffu, i2, dd);
    if (lane >= dd) { i2 += o2 + (double)(dd * EPL) * o1; i1 += o1; }
  }
  double c1 = __shfl_up_sync(0xffffffffu, i1, 1), c2 = __shfl_up_sync(0xffffffffu, i2, 1);
  if (lane == 0) { c1 = 0.0; c2 = 0.0; }
#pragma unroll
  for (int e = 0; e < EPL; ++e) { c[e] += c2 + (double)(e + 1) * c1; b[e] += c1; }
}

constexpr long long P2P_SPIN_LIMIT = 6000000000LL;   // ~3 s: a lost peer must not hang the box

// positions of launch n (buffer `cur`) are complete, and every peer has finished reading the buffer this launch
// overwrites, once every peer has published launch n
__device__ __forceinline__ void p2p_wait(const Dev& d) {
  const unsigned n = *(volatile unsigned*)d.iter_no;
  const long long t0 = clock64();
  for (int g = 0; g < d.G; ++g) {
    if (g == d.rank) continue;
    while (*(volatile unsigned*)(d.flag + g) < n) {
      if (clock64() - t0 > P2P_SPIN_LIMIT) { *d.err = 1u; break; }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

// BOX = 0: no box-row class is carried by this scenario's ADMM (the usual state with lazy rows) -- all box code
// compiles away and only x and the positions move through memory.
template <int EPL, int CHK, int BOX, int VEC>
__device__ __forceinline__ void iter_body(const Dev& d, const State& S, int apc, int cur, double* sm) {
  const int b = blockIdx.y;
  const int K = d.K, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, npair = blockDim.x >> 6;
  const int ia = d.a_lo + blockIdx.x * apc, ib = min(ia + apc, d.a_hi);
  const int ax = warp & 1;                                    // two warps per agent: one per axis
  double* Nm = sm;
  const int Kr = K + (K & 1);
  double* rhs_rows = sm + (size_t)K * K + 4;                                // 4 doubles of padding behind the operator
  double* myrhs = rhs_rows + (size_t)warp * Kr;
  double* xout = rhs_rows + (size_t)(blockDim.x >> 5) * Kr;                 // [8 right-hand sides][Kr]: N rhs, written by the tensor-pipe product
  constexpr bool vec = VEC != 0;      // K even: rows are 16-byte aligned and move as double2 (compile time: no dual code paths)
  if (threadIdx.x < 4) sm[(size_t)K * K + threadIdx.x] = 0.0;
  // operator -> shared memory: ONE TMA bulk copy (cp.async.bulk, completion on an mbarrier) issued by thread 0 when
  // the rows are 16-byte aligned; it lands while the collision rows are walked
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(&mbar);
  {
    const double* src = d.Nmat + (size_t)b * K * K;
    if (vec) {
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_a), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned bytes = (unsigned)((size_t)K * K * sizeof(double));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(Nm)),
                     "l"(src), "r"(bytes), "r"(mbar_a)
                     : "memory");
      }
    } else {
      for (int e = threadIdx.x; e < K * K; e += blockDim.x) Nm[e] = src[e];
    }
  }
  if (d.p2p) {
    if (threadIdx.x == 0) p2p_wait(d);
    __syncthreads();
  }
  const double h = d.pb.time_step, ih = 1.0 / h, rho = S.rho, sig = d.pb.sigma;
  const double vl = d.pb.vel_limit, al = d.pb.acc_limit, jl = d.pb.jerk_limit;
  const double cpr = (double)S.copies * rho;
  const double alpha = d.pb.relax_pct > 0 ? 0.01 * (double)d.pb.relax_pct : 1.0;
  const int on = BOX ? S.on_mask : 0;
  // heavy-ball extrapolation of the collision state (positions in the prox term, forces, multipliers); only without box rows
  const double beta = (!BOX && S.phase == 1 && d.pb.momentum_pct > 0) ? 0.01 * (double)d.pb.momentum_pct : 0.0;
  const double* N0 = d.N0 + (size_t)b * 2 * K;
  const double* Qm = d.Qm + (size_t)b * 2 * K;
  const double* Pc = (cur ? d.P1 : d.P) + (size_t)b * d.Qs * K;
  double* Pn = (cur ? d.P : d.P1) + (size_t)b * d.Qs * K;
  const double* lamc = cur ? d.lam1 : d.lam;
  double* lamn = cur ? d.lam : d.lam1;
  const int k0 = lane * EPL;
  const int Nown = d.a_hi - d.a_lo;
  const size_t T = (size_t)d.B * Nown * K;
  double trj[EPL], tra[EPL], trv[EPL], trp[EPL], trc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int k = k0 + e;
    const bool in = k < K;
    trj[e] = (in && (on & 1)) ? rho * d.rj[k] : 0.0; tra[e] = (in && (on & 2)) ? rho * d.ra[k] : 0.0;
    trv[e] = (in && (on & 4)) ? rho * d.rv[k] : 0.0; trp[e] = (in && (on & 8)) ? rho * d.rp[k] : 0.0;
    trc[e] = in ? cpr * d.rc[k] : 0.0;
  }
  double pr = 0.0, nr = 0.0, voj = 0.0, voa = 0.0, vov = 0.0, vop = 0.0, worst = 0.0;
  bool staged = false;
  // Rounds are uniform over the CTA (the operator product of a round is a block-wide tensor-pipe GEMM over the CTA's
  // eight agent-axes): a warp without an agent of its own in a round recomputes the CTA's first agent and stores nothing.
  const int rounds = (ib - ia + npair - 1) / npair;
  for (int rnd = 0; rnd < rounds; ++rnd) {
    const int i_raw = ia + (warp >> 1) + rnd * npair;
    const bool live = i_raw < ib;
    const int i = live ? i_raw : ia;
    const double pr_s = pr, nr_s = nr, voj_s = voj, voa_s = voa, vov_s = vov, vop_s = vop;   // a recomputed agent adds nothing
    // ---- 1. collision rows of agent i: this axis' component of the force on state k (fz[e] <-> state k0+e).
    // Both axis warps walk the rows and compute the same multipliers; the x warp stores them (other buffer).
    double fz[EPL], fyo[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) { fz[e] = 0.0; fyo[e] = 0.0; }
    if (live && S.phase == 1) {
      const int il = i - d.a_lo;
      int n[EPL], nmax = 0;
      double pix[EPL], piy[EPL], rce[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = k0 + e;
        n[e] = (k >= 1 && k < K) ? d.cnt[((size_t)b * Nown + il) * K + k] : 0;
      }
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = k0 + e;
        nmax = n[e] > nmax ? n[e] : nmax;
        pix[e] = piy[e] = 0.0; rce[e] = 1.0;
        if (n[e] > 0) { pix[e] = Pc[(size_t)(2 * i) * K + k]; piy[e] = Pc[(size_t)(2 * i + 1) * K + k]; rce[e] = rho * d.rc[k - 1]; }
      }
      // candidate walk, software pipelined: the slot data of candidate s2+1 is in flight while the partner positions
      // of candidate s2 (addresses depend on cj) are fetched; loads of the EPL steps are independent, stores come last
      const int* __restrict__ cjp = d.cj;
      const double* __restrict__ cexp = d.cex;
      const double* __restrict__ ceyp = d.cey;
      const double* __restrict__ cbp = d.cb;
      const size_t tb = ((size_t)b * Nown + il) * K + k0;
      int jn[EPL];
      double exn[EPL], eyn[EPL], cbn[EPL], l0n[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        jn[e] = 0; exn[e] = eyn[e] = cbn[e] = l0n[e] = 0.0;
        if (0 < n[e]) { const size_t o = tb + e; jn[e] = cjp[o]; exn[e] = cexp[o]; eyn[e] = ceyp[o]; cbn[e] = cbp[o]; l0n[e] = lamc[o]; }
      }
      for (int s2 = 0; s2 < nmax; ++s2) {
        int jc[EPL];
        double exc[EPL], eyc[EPL], cbc[EPL], l0c[EPL], pjx[EPL], pjy[EPL], l1v[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) { jc[e] = jn[e]; exc[e] = exn[e]; eyc[e] = eyn[e]; cbc[e] = cbn[e]; l0c[e] = l0n[e]; }
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          pjx[e] = pjy[e] = 0.0;
          if (s2 < n[e]) { const int k = k0 + e; pjx[e] = Pc[(size_t)(2 * jc[e]) * K + k]; pjy[e] = Pc[(size_t)(2 * jc[e] + 1) * K + k]; }
        }
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          if (s2 + 1 < n[e]) { const size_t o = (size_t)(s2 + 1) * T + tb + e; jn[e] = cjp[o]; exn[e] = cexp[o]; eyn[e] = ceyp[o]; cbn[e] = cbp[o]; l0n[e] = lamc[o]; }
        }
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          l1v[e] = 0.0;
          if (s2 < n[e]) {
            const double g = exc[e] * (pix[e] - pjx[e]) + eyc[e] * (piy[e] - pjy[e]);
            const double l1 = fmax(0.0, l0c[e] + 0.5 * rce[e] * (cbc[e] - g));
            l1v[e] = l1;
            const double ea = ax ? eyc[e] : exc[e];
            fz[e] += (2.0 * l1 - l0c[e]) * ea;
            if (CHK) { fyo[e] += l1 * ea; worst = fmax(worst, fabs(l1 - l0c[e]) / rce[e]); }
          }
        }
        if (ax == 0) {
#pragma unroll
          for (int e = 0; e < EPL; ++e)
            if (s2 < n[e]) {
              const size_t o = (size_t)s2 * T + tb + e;
              double lh = l1v[e];
              if (beta > 0.0) { lh = fmax(0.0, l1v[e] + beta * (l1v[e] - d.lamt[o])); d.lamt[o] = l1v[e]; }
              lamn[o] = lh;
            }
        }
      }
    }
    if (!staged) {
      __syncthreads();              // mbarrier initialised (thread 0) / plain staging stores done
      if (vec) {
        unsigned ok = 0;
        do {
          asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                       : "=r"(ok) : "r"(mbar_a), "r"(0u) : "memory");
        } while (!ok);
      }
      staged = true;
    }
    // ---- 2. this warp's axis of agent i
    {
      const int q = 2 * i + ax;
      const size_t row = ((size_t)b * d.Qs + q) * K;
      const size_t q2 = (size_t)b * d.Q + q;
      const double v0q = d.v0[q2], p0q = d.p0[q2];
      const double lv = -vl - v0q, uv = vl - v0q;
      const double plo = d.pb.space[ax], phi = d.pb.space[2 + ax];
      double xo[EPL], va[EPL], vj[EPL], vv[EPL], vp[EPL], pc1[EPL], off[EPL];
      load_row<EPL>(d.x + row, k0, K, vec, xo);
      if (BOX) {
        if (on & 2) load_row<EPL>(d.va + row, k0, K, vec, va);
        if (on & 1) load_row<EPL>(d.vj + row, k0, K, vec, vj);
        if (on & 4) load_row<EPL>(d.vv + row, k0, K, vec, vv);
        if (on & 8) load_row<EPL>(d.vp + row, k0, K, vec, vp);
      }
      // previous position of state k+1 (prox term of the padded copies): pc1[e] = Pc[q][k+1]
#pragma unroll
      for (int e = 0; e < EPL; ++e) { const int k = k0 + e; pc1[e] = (k < K - 1) ? Pc[(size_t)q * K + k + 1] : 0.0; }
      if (beta > 0.0) {
        // own row of the other buffer still holds the positions of two iterations ago; forces of the last one are in F
        double fold[EPL];
        load_row<EPL>(d.F + row, k0, K, vec, fold);
        if (live) store_row<EPL>(d.F + row, k0, K, vec, fz);
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          const int k = k0 + e;
          if (k < K - 1) pc1[e] += beta * (pc1[e] - Pn[(size_t)q * K + k + 1]);
          fz[e] += beta * (fz[e] - fold[e]);
        }
      }
      double sj[EPL], sa[EPL], sv[EPL], sp[EPL], wj[EPL], wa[EPL], r1v[EPL], r1p[EPL], r2p[EPL];
      // the force on state k+1 enters row k: shift by one step
      double fnext = __shfl_down_sync(0xffffffffu, fz[0], 1);
      if (lane == 31) fnext = 0.0;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = k0 + e;
        sj[e] = sa[e] = sv[e] = sp[e] = 0.0; wj[e] = wa[e] = r1v[e] = r1p[e] = 0.0; off[e] = 0.0;
        if (k < K) {
          if (BOX && (on & 2)) { const double v = va[e], z = clampd(v, -al, al); sa[e] = v - alpha * z; wa[e] = tra[e] * (2 * z - v); }
          if (k < K - 1) {
            if (BOX && (on & 1)) { const double v = vj[e], z = clampd(v, -jl, jl); sj[e] = v - alpha * z; wj[e] = trj[e] * (2 * z - v); }
            if (BOX && (on & 4)) { const double v = vv[e], z = clampd(v, lv, uv); sv[e] = v - alpha * z; r1v[e] = trv[e] * (2 * z - v); }
            off[e] = p0q + h * (double)(k + 1) * v0q;
            double wp = trc[e] * (pc1[e] - off[e]) + ((e + 1 < EPL) ? fz[(e + 1 < EPL) ? e + 1 : e] : fnext);
            if (BOX && (on & 8)) { const double v = vp[e], z = clampd(v, plo - off[e], phi - off[e]); sp[e] = v - alpha * z; wp += trp[e] * (2 * z - v); }
            r1p[e] = wp;
          }
        }
      }
      suffix3<EPL>(r1v, r1p, r2p, lane);
      double wprev = 0.0;
      if (BOX) { wprev = __shfl_up_sync(0xffffffffu, wj[EPL - 1], 1); if (lane == 0) wprev = 0.0; }
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = k0 + e;
        const double prev = e == 0 ? wprev : wj[e == 0 ? 0 : e - 1];
        if (k < K) myrhs[k] = sig * xo[e] + (prev - wj[e]) * ih + wa[e] + h * r1v[e] + h * h * (r2p[e] - 0.5 * r1p[e]);
      }
      // x = Nmat rhs + N0 d for the CTA's eight agent-axes at once on the fp64 tensor pipe (mma.sync m8n8k4, SASS DMMA):
      // M = K steps in tiles of 8 (tiles `warp` and `warp + 8` of this warp, two accumulator chains in flight), N = the 8
      // right-hand-side rows, K-dimension in steps of 4.  Operand rows K..Kr-1 / steps beyond K are masked to zero.
      const double d0 = d.vf[q2] - v0q, d1 = d.pf[q2] - (p0q + h * (double)K * v0q);
      __syncthreads();
      {
        const int g = lane >> 2, t = lane & 3, nw8 = blockDim.x >> 5;
        const int MT = (K + 7) >> 3;
        const int m0 = warp * 8 + g, m1 = (warp + nw8) * 8 + g;
        const bool has1 = warp + nw8 < MT;
        const double* a0p = Nm + (size_t)(m0 < K ? m0 : K - 1) * K;
        const double* a1p = Nm + (size_t)(m1 < K ? m1 : K - 1) * K;
        const double* bp = rhs_rows + (size_t)g * Kr;
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
        if (warp < MT) {
          for (int kk = 0; kk < K; kk += 4) {
            const int ka = kk + t;
            const bool in = ka < K;
            const int kc = in ? ka : K - 1;
            const double bv = in ? bp[kc] : 0.0;
            const double av0 = in ? a0p[kc] : 0.0;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c00), "+d"(c01) : "d"(av0), "d"(bv));
            if (has1) {
              const double av1 = in ? a1p[kc] : 0.0;
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c10), "+d"(c11) : "d"(av1), "d"(bv));
            }
          }
          if (m0 < K) { xout[(size_t)(2 * t) * Kr + m0] = c00; xout[(size_t)(2 * t + 1) * Kr + m0] = c01; }
          if (has1 && m1 < K) { xout[(size_t)(2 * t) * Kr + m1] = c10; xout[(size_t)(2 * t + 1) * Kr + m1] = c11; }
        }
      }
      __syncthreads();
      double a0[EPL], a1[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) { const int k = k0 + e; a0[e] = (k < K) ? xout[(size_t)warp * Kr + k] : 0.0; a1[e] = 0.0; }
      double xn[EPL], c1[EPL], c2[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = k0 + e;
        xn[e] = (k < K) ? N0[2 * k] * d0 + N0[2 * k + 1] * d1 + (a0[e] + a1[e]) : 0.0;
        c1[e] = xn[e];
      }
      if (CHK) {
        double m0 = 0.0, m1 = 0.0;
        for (int jj = lane; jj < K; jj += 32) { const double r = myrhs[jj]; m0 += Qm[jj] * r; m1 += Qm[K + jj] * r; }
        m0 = warp_sum(m0); m1 = warp_sum(m1);
        if (lane == 0) {
          const double* gg = d.gg + (size_t)b * 4;
          if (live) {
            d.mu[((size_t)b * d.Qs + q) * 2] = m0 - (gg[0] * d0 + gg[1] * d1);
            d.mu[((size_t)b * d.Qs + q) * 2 + 1] = m1 - (gg[1] * d0 + gg[2] * d1);
          }
        }
      }
      __syncwarp();
      prefix2<EPL>(c1, c2, lane);
      double xnext = __shfl_down_sync(0xffffffffu, xn[0], 1);
      if (lane == 31) xnext = 0.0;
      double nva[EPL], nvj[EPL], nvv[EPL], nvp[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int k = k0 + e;
        nva[e] = nvj[e] = nvv[e] = nvp[e] = 0.0;
        if (k < K) {
          nva[e] = ((on & 2) ? alpha : 1.0) * xn[e] + sa[e];
          if (CHK) {
            const double dv = fabs(xn[e] - clampd(nva[e], -al, al));
            if (on & 2) pr = fmax(pr, dv); else voa = fmax(voa, dv);
            nr = fmax(nr, fabs(xn[e]));
          }
          if (k < K - 1) {
            const double rv_ = h * c1[e], rp_ = h * h * (c2[e] - 0.5 * c1[e]);
            const double aj = (((e + 1 < EPL) ? xn[(e + 1 < EPL) ? e + 1 : e] : xnext) - xn[e]) * ih;
            nvj[e] = ((on & 1) ? alpha : 1.0) * aj + sj[e];
            nvv[e] = ((on & 4) ? alpha : 1.0) * rv_ + sv[e];
            nvp[e] = ((on & 8) ? alpha : 1.0) * rp_ + sp[e];
            if (live) Pn[(size_t)q * K + k + 1] = off[e] + rp_;
            if (d.p2p && live) {   // the own slice goes straight into every peer's buffer over NVLink (B = 1 when sharded)
              for (int g = 0; g < d.G; ++g)
                if (g != d.rank) (cur ? d.peerP[g] : d.peerP1[g])[(size_t)q * K + k + 1] = off[e] + rp_;
            }
            if (CHK) {
              const double dj = fabs(aj - clampd(nvj[e], -jl, jl)), dv = fabs(rv_ - clampd(nvv[e], lv, uv));
              const double dp = fabs(rp_ - clampd(nvp[e], plo - off[e], phi - off[e]));
              if (on & 1) pr = fmax(pr, dj); else voj = fmax(voj, dj);
              if (on & 4) pr = fmax(pr, dv); else vov = fmax(vov, dv);
              if (on & 8) pr = fmax(pr, dp); else vop = fmax(vop, dp);
              nr = fmax(nr, fmax(fabs(aj), fmax(fabs(rv_), fabs(rp_))));
            }
          }
        }
      }
      if (live) {
        store_row<EPL>(d.x + row, k0, K, vec, xn);
        if (BOX) {
          if (on & 2) store_row<EPL>(d.va + row, k0, K, vec, nva);
          if (on & 1) store_row<EPL>(d.vj + row, k0, K, vec, nvj);
          if (on & 4) store_row<EPL>(d.vv + row, k0, K, vec, nvv);
          if (on & 8) store_row<EPL>(d.vp + row, k0, K, vec, nvp);
        }
        if (CHK) store_row<EPL>(d.FY + row, k0, K, vec, fyo);
      } else if (CHK) {
        pr = pr_s; nr = nr_s; voj = voj_s; voa = voa_s; vov = vov_s; vop = vop_s;
      }
      __syncwarp();
    }
  }
  if (CHK) {
    pr = warp_max_nan(fabs(pr)); nr = warp_max_nan(fabs(nr)); worst = warp_max_nan(fabs(worst));
    voj = warp_max(voj); voa = warp_max(voa); vov = warp_max(vov); vop = warp_max(vop);
    if (lane == 0) {
      double* sl = d.slab + (size_t)b * NRED;
      atomic_max_pos(sl + R_PRI, pr); atomic_max_pos(sl + R_NPRI, nr);
      if (worst > 0.0) atomic_max_pos(sl + R_PRICOL, worst);
      if (voj > 0.0) atomic_max_pos(sl + R_VIOLJ, voj);
      if (voa > 0.0) atomic_max_pos(sl + R_VIOLA, voa);
      if (vov > 0.0) atomic_max_pos(sl + R_VIOLV, vov);
      if (vop > 0.0) atomic_max_pos(sl + R_VIOLP, vop);
    }
  }
  if (d.p2p) {
    // publish: all stores of this CTA (local and remote) are ordered before the flag; the last CTA of the launch
    // bumps the launch counter and writes it into every peer's flag array
    __syncthreads();                 // the CTA's stores happen-before thread 0's system fence (cumulativity)
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned fin = atomicAdd(d.cta_done, 1u);
      if (fin == gridDim.x - 1) {
        *(volatile unsigned*)d.cta_done = 0u;
        const unsigned n = *(volatile unsigned*)d.iter_no + 1u;
        *(volatile unsigned*)d.iter_no = n;
        __threadfence_system();
        for (int g = 0; g < d.G; ++g)
          if (g != d.rank) *(volatile unsigned*)(d.peerFlag[g] + d.rank) = n;
      }
    }
  }
}

template <int EPL, int CHK>
__global__ void __launch_bounds__(IT_THREADS, 2) k_iter(const __grid_constant__ Dev d, int apc, int cur) {
  extern __shared__ double sm[];
  const State& S = d.st[blockIdx.y];
  if (S.phase >= 2) return;
  if (d.a_lo + (int)blockIdx.x * apc >= d.a_hi) return;
  const bool even = (d.K & 1) == 0;
  if (S.on_mask == 0) { if (even) iter_body<EPL, CHK, 0, 1>(d, S, apc, cur, sm); else iter_body<EPL, CHK, 0, 0>(d, S, apc, cur, sm); }
  else { if (even) iter_body<EPL, CHK, 1, 1>(d, S, apc, cur, sm); else iter_body<EPL, CHK, 1, 0>(d, S, apc, cur, sm); }
}

// the kernels after a check iteration read every agent's positions: wait until all peers have published it
__global__ void k_wait(const __grid_constant__ Dev d) {
  if (d.p2p && threadIdx.x == 0 && blockIdx.x == 0 && d.st[0].phase < 2) p2p_wait(d);
}

// ---------------------------------------------------------------------------------- dual residual
// rhs <- 2x + A'y + C'mu per agent-axis (scp_device.inl transpose_rows mode 1) reduced to max|.|, plus the
// per-agent-axis sums ||x - xprev||^2, ||xprev||^2 (scp.py:157-160) and ||x||^2.
template <int EPL>
__global__ void __launch_bounds__(AX_THREADS) k_dual(const __grid_constant__ Dev d) {
  const int b = blockIdx.y;
  const State& S = d.st[b];
  if (S.phase >= 2) return;
  const int K = d.K, lane = threadIdx.x & 31;
  const int q = 2 * d.a_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (q >= 2 * d.a_hi) return;
  const double h = d.pb.time_step, ih = 1.0 / h, rho = S.rho;
  const double vl = d.pb.vel_limit, al = d.pb.acc_limit, jl = d.pb.jerk_limit;
  const size_t row = ((size_t)b * d.Qs + q) * K;
  const size_t q2 = (size_t)b * d.Q + q;
  const double v0q = d.v0[q2], p0q = d.p0[q2];
  const double lv = -vl - v0q, uv = vl - v0q;
  const double plo = d.pb.space[q & 1], phi = d.pb.space[2 + (q & 1)];
  const int on = S.on_mask;
  const double *x = d.x + row, *xp = d.xprev + row, *vj = d.vj + row, *va = d.va + row, *vv = d.vv + row, *vp = d.vp + row, *FY = d.FY + row;
  double xo[EPL], wj[EPL], wa[EPL], r1v[EPL], r1p[EPL], r2p[EPL];
  double dn = 0.0, pn = 0.0, ob = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int k = lane + 32 * e;
    xo[e] = 0; wj[e] = wa[e] = r1v[e] = r1p[e] = 0;
    if (k < K) {
      xo[e] = x[k];
      const double xq = xp[k];
      dn += (xo[e] - xq) * (xo[e] - xq); pn += xq * xq; ob += xo[e] * xo[e];
      double v = va[k], z = clampd(v, -al, al);
      wa[e] = (on & 2) ? rho * d.ra[k] * (v - z) : 0.0;
      if (k < K - 1) {
        v = vj[k]; z = clampd(v, -jl, jl); wj[e] = (on & 1) ? rho * d.rj[k] * (v - z) : 0.0;
        v = vv[k]; z = clampd(v, lv, uv); r1v[e] = (on & 4) ? rho * d.rv[k] * (v - z) : 0.0;
        const double off = p0q + h * (double)(k + 1) * v0q;
        v = vp[k]; z = clampd(v, plo - off, phi - off); r1p[e] = ((on & 8) ? rho * d.rp[k] * (v - z) : 0.0) - FY[k + 1];
      }
    }
  }
  suffix_sum<EPL>(r1v, lane);
  suffix_sum<EPL>(r1p, lane);
#pragma unroll
  for (int e = 0; e < EPL; ++e) r2p[e] = r1p[e];
  suffix_sum<EPL>(r2p, lane);
  const double mu0 = d.mu[((size_t)b * d.Qs + q) * 2], mu1 = d.mu[((size_t)b * d.Qs + q) * 2 + 1];
  double du = 0.0, nd = 0.0, last = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int k = lane + 32 * e;
    double prev = __shfl_up_sync(0xffffffffu, wj[e], 1);
    if (lane == 0) prev = last;
    last = __shfl_sync(0xffffffffu, wj[e], 31);
    if (k < K) {
      const double o = wa[e] + (prev - wj[e]) * ih + h * r1v[e] + h * h * (r2p[e] - 0.5 * r1p[e]) + 2.0 * xo[e] + h * mu0 +
                       h * h * ((double)(K - 1 - k) + 0.5) * mu1;
      du = fmax(du, fabs(o));
      nd = fmax(nd, fmax(fabs(2.0 * xo[e]), fabs(o - 2.0 * xo[e])));
    }
  }
  du = warp_max_nan(fabs(du)); nd = warp_max_nan(fabs(nd));
  dn = warp_sum(dn); pn = warp_sum(pn); ob = warp_sum(ob);
  if (lane == 0) {
    atomic_max_pos(d.slab + (size_t)b * NRED + R_DUA, du); atomic_max_pos(d.slab + (size_t)b * NRED + R_NDUA, nd);
    double* qs = d.qsum + ((size_t)b * d.Qs + q) * 3;
    qs[0] = dn; qs[1] = pn; qs[2] = ob;
  }
}

// deterministic per-rank sums of the per-agent-axis partials: one warp per scenario
__global__ void k_local_sums(const __grid_constant__ Dev d) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= d.B) return;
  if (d.st[b].phase >= 2) return;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int q = 2 * d.a_lo + lane; q < 2 * d.a_hi; q += 32) {
    const double* qs = d.qsum + ((size_t)b * d.Qs + q) * 3;
    s0 += qs[0]; s1 += qs[1]; s2 += qs[2];
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { double* sl = d.slab + (size_t)b * NRED; sl[R_DN] = s0; sl[R_PN] = s1; sl[R_OBJ] = s2; }
}

// ---------------------------------------------------------------------------------- snapshot / candidates
__global__ void k_snapshot(const __grid_constant__ Dev d) {
  const int b = blockIdx.y;
  if (!(d.st[b].flags & FL_SNAPSHOT)) return;
  const size_t base = (size_t)b * d.Qs * d.K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.Qs * d.K; e += gridDim.x * blockDim.x) {
    d.Pbar[base + e] = d.P[base + e];
    d.xprev[base + e] = d.x[base + e];
  }
}

__device__ __forceinline__ void linearise_pair(double dx, double dy, int i, int j, double R, double& ex, double& ey, double& bound) {
  // scp.py:498-509, 547-549; the degenerate branch draws a random direction there, a fixed one here
  const double dist = hypot(dx, dy);
  if (dist < 1e-6) { ex = i < j ? 1.0 : -1.0; ey = 0.0; bound = R + (ex * dx + ey * dy - 1.0); }
  else { ex = dx / dist; ey = dy / dist; bound = R + ((ex * dx + ey * dy) - dist); }
}

// Candidate rows of (k, i): partners whose linearisation-point distance is below R + margin (multipliers start at
// zero: OSQP starts every subproblem from y = 0, scp.py:441-443).  Rows that a later verification finds violated
// are appended by k_scan.
__global__ void __launch_bounds__(256) k_build(const __grid_constant__ Dev d) {
  const int b = blockIdx.y;
  const State& S = d.st[b];
  if (!(S.flags & FL_BUILD)) return;
  const int K = d.K, N = d.N, Nown = d.a_hi - d.a_lo;
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  int mx = 0, over = 0;
  if (tl < Nown * K) {
    const int il = tl / K, k = tl - il * K, i = d.a_lo + il;
    if (k >= 1) {
      const size_t T = (size_t)d.B * Nown * K, t = ((size_t)b * Nown + il) * K + k;
      // warm duals: a row carried by the previous subproblem keeps its multiplier (same minimiser, fewer iterations;
      // the reference restarts OSQP from y = 0 every SCP iteration, scp.py:441-443)
      int oj[MAXC_MAX];
      double ol[MAXC_MAX];
      int on = 0;
      if (d.pb.warm_duals && S.scp_it > 0) {
        on = d.cnt[t];
        for (int s = 0; s < on; ++s) { oj[s] = d.cj[(size_t)s * T + t]; ol[s] = d.lam[(size_t)s * T + t]; }
      }
      const double* Pb = d.Pbar + (size_t)b * d.Qs * K;
      const double pix = Pb[(size_t)(2 * i) * K + k], piy = Pb[(size_t)(2 * i + 1) * K + k];
      const double R = d.pb.min_distance, r2 = (R + S.margin) * (R + S.margin);
      int n = 0;
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        const double dx = pix - Pb[(size_t)(2 * j) * K + k], dy = piy - Pb[(size_t)(2 * j + 1) * K + k];
        if (dx * dx + dy * dy < r2) {
          if (n < d.maxc) {
            double ex, ey, bound, l = 0.0;
            linearise_pair(dx, dy, i, j, R, ex, ey, bound);
            for (int s = 0; s < on; ++s) if (oj[s] == j) l = ol[s];
            const size_t o = (size_t)n * T + t;
            d.cj[o] = j; d.cex[o] = ex; d.cey[o] = ey; d.cb[o] = bound; d.lam[o] = l; d.lamt[o] = l;
            ++n;
          } else over = 1;
        }
      }
      d.cnt[t] = n; mx = n;
      const size_t r0 = ((size_t)b * d.Qs + 2 * i) * K + k;
      d.F[r0] = 0.0; d.F[r0 + K] = 0.0; d.FY[r0] = 0.0; d.FY[r0 + K] = 0.0;
    }
  }
  double nc = warp_sum((double)mx);
  double m = warp_max((double)mx), ov = warp_max((double)over);
  if ((threadIdx.x & 31) == 0) {
    double* sl = d.slab + (size_t)b * NRED;
    if (m > 0.0) atomic_max_pos(sl + R_COPIES, m);
    if (nc > 0.0) atomicAdd(sl + R_NCAND, nc);      // integer valued: exact in any order
    if (ov > 0.0) atomic_max_pos(sl + R_OVER, ov);
  }
}

// ---------------------------------------------------------------------------------- pair scan
// One thread per (scenario, own agent i, step k): all partners j.  Always: min separation of the current positions
// (scp.py:597-615 quantity).  FL_GATE: first row (k-major, i<j) below R - margin.  FL_VERIFY: every row of the full
// QP that is NOT carried is evaluated; a violated one joins the candidate rows of (k, i) with multiplier 0 (its other
// owner finds the same violation from the same numbers and appends its own copy), so the solve continues warm.
__global__ void __launch_bounds__(256) k_scan(const __grid_constant__ Dev d) {
  const int b = blockIdx.y;
  const State& S = d.st[b];
  if (!(S.flags & FL_SCAN)) return;
  const int gate = S.flags & FL_GATE, verify = S.flags & FL_VERIFY;
  const int K = d.K, N = d.N, Nown = d.a_hi - d.a_lo;
  const int tl = blockIdx.x * blockDim.x + threadIdx.x;
  double mn = INFINITY, fr = INFINITY, bad = 0.0, maxd = 0.0, over = 0.0;
  if (tl < Nown * K) {
    const int il = tl / K, k = tl - il * K, i = d.a_lo + il;
    const size_t T = (size_t)d.B * Nown * K, t = ((size_t)b * Nown + il) * K + k;
    int n = (verify && k >= 1) ? d.cnt[t] : 0;
    const int n0 = n;
    const double* Pc = d.P + (size_t)b * d.Qs * K;
    const double* Pb = d.Pbar + (size_t)b * d.Qs * K;
    const double px = Pc[(size_t)(2 * i) * K + k], py = Pc[(size_t)(2 * i + 1) * K + k];
    const double bx = Pb[(size_t)(2 * i) * K + k], by = Pb[(size_t)(2 * i + 1) * K + k];
    const double R = d.pb.min_distance, thr = R - d.pb.feas_margin, r2 = (R + S.margin) * (R + S.margin), tol = d.pb.verify_tol;
    const double npairs = (double)N * (double)(N - 1) * 0.5;
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const double cx = px - Pc[(size_t)(2 * j) * K + k], cy = py - Pc[(size_t)(2 * j + 1) * K + k];
      const double dd = sqrt(cx * cx + cy * cy);           // np.linalg.norm of a 2-vector
      mn = fmin(mn, dd);
      if (gate && j > i && dd < thr) {
        const double rowi = (double)k * npairs + (double)(((long long)i * (2 * N - i - 1)) / 2 + (j - i - 1));
        fr = fmin(fr, rowi);
      }
      if (verify && k >= 1) {
        const double dx = bx - Pb[(size_t)(2 * j) * K + k], dy = by - Pb[(size_t)(2 * j + 1) * K + k];
        const double d2 = dx * dx + dy * dy;
        if (!(d2 < r2)) {
          double ex, ey, bound;
          linearise_pair(dx, dy, i, j, R, ex, ey, bound);
          if (ex * cx + ey * cy < bound - tol) {
            int carried = 0;
            for (int s = 0; s < n; ++s) carried |= (d.cj[(size_t)s * T + t] == j);
            if (!carried) {
              if (n < d.maxc) {
                const size_t o = (size_t)n * T + t;
                d.cj[o] = j; d.cex[o] = ex; d.cey[o] = ey; d.cb[o] = bound; d.lam[o] = 0.0; d.lamt[o] = 0.0;
                ++n; bad += 1.0;
              } else over = 1.0;
            }
          }
        }
      }
    }
    if (n != n0) d.cnt[t] = n;
    maxd = (double)n;                                 // largest candidate count after the appends
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    fr = fmin(fr, __shfl_xor_sync(0xffffffffu, fr, s));
    bad += __shfl_xor_sync(0xffffffffu, bad, s);
    maxd = fmax(maxd, __shfl_xor_sync(0xffffffffu, maxd, s));
    over = fmax(over, __shfl_xor_sync(0xffffffffu, over, s));
  }
  if ((threadIdx.x & 31) == 0) {
    double* sl = d.slab + (size_t)b * NRED;
    if (mn < INFINITY) atomic_min_pos(sl + R_MINSEP, mn);
    if (fr < INFINITY) atomic_min_pos(sl + R_FIRST, fr);
    if (bad > 0.0) atomicAdd(sl + R_BAD, bad);        // rows appended (integer valued: exact in any order)
    if (maxd > 0.0) atomic_max_pos(sl + R_MAXD, maxd);
    if (over > 0.0) atomic_max_pos(sl + R_OVER, over);
  }
}

// ---------------------------------------------------------------------------------- rho rescale
// keep y when rho changes: v = z + y/rho  ->  v = z + (v - z)/est   (scp_device.inl admm_run)
__global__ void k_rescale(const __grid_constant__ Dev d) {
  const int b = blockIdx.y;
  const State& S = d.st[b];
  if (!(S.flags & FL_RESCALE)) return;
  const int K = d.K, nq = 2 * (d.a_hi - d.a_lo);
  const double est = S.est, h = d.pb.time_step;
  const double vl = d.pb.vel_limit, al = d.pb.acc_limit, jl = d.pb.jerk_limit;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nq * K; e += gridDim.x * blockDim.x) {
    const int ql = e / K, k = e - ql * K, q = 2 * d.a_lo + ql;
    const size_t o = ((size_t)b * d.Qs + q) * K + k;
    const int on = S.on_mask;
    double v = d.va[o], z = clampd(v, -al, al); if (on & 2) d.va[o] = z + (v - z) / est;
    if (k < K - 1) {
      const double v0q = d.v0[(size_t)b * d.Q + q], p0q = d.p0[(size_t)b * d.Q + q];
      const double off = p0q + h * (double)(k + 1) * v0q;
      v = d.vj[o]; z = clampd(v, -jl, jl); if (on & 1) d.vj[o] = z + (v - z) / est;
      v = d.vv[o]; z = clampd(v, -vl - v0q, vl - v0q); if (on & 4) d.vv[o] = z + (v - z) / est;
      v = d.vp[o]; z = clampd(v, d.pb.space[q & 1] - off, d.pb.space[2 + (q & 1)] - off); if (on & 8) d.vp[o] = z + (v - z) / est;
    }
  }
}

// ---------------------------------------------------------------------------------- outputs
// scp.py:168-175: accelerations, positions, velocities of states k = 0..K-1, reference layout (N, K, 2).
__global__ void k_output(const __grid_constant__ Dev d) {
  const int b = blockIdx.y;
  if (!(d.st[b].flags & FL_FINISH)) return;
  const int K = d.K;
  const int q = 2 * d.a_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * d.a_hi) return;
  const size_t row = ((size_t)b * d.Qs + q) * K;
  const double v0q = d.v0[(size_t)b * d.Q + q], h = d.pb.time_step;
  const int i = q >> 1, ax = q & 1;
  double c1 = 0.0;
  for (int k = 0; k < K; ++k) {
    const size_t o = (((size_t)b * d.Npad + i) * K + k) * 2 + ax;
    const double xk = d.x[row + k];
    d.acc[o] = xk;
    d.pos[o] = d.P[row + k];
    d.vel[o] = v0q + h * c1;
    c1 += xk;
  }
}

// ---------------------------------------------------------------------------------- control
__device__ double combine(const Dev& d, int b, int slot, int kind) {   // kind 0 max, 1 sum, 2 min; fixed rank order
  double r = kind == 2 ? INFINITY : 0.0;
  for (int g = 0; g < d.G; ++g) {
    const double v = d.gath[((size_t)g * d.B + b) * NRED + slot];
    if (kind == 0) { if (!(v <= r)) r = v; }            // NaN wins
    else if (kind == 1) r += v;
    else r = fmin(r, v);
  }
  return r;
}

__device__ void start_qp(State& S, const Dev& d) {
  S.qp_it = 0; S.it_mark = 0; S.pri_mark = INFINITY; S.stalled = 0; S.qp_solved = 0;
}

// after the check iteration: ADMM termination test of the running subproblem (OSQP's, in reference units)
__global__ void k_control1(const __grid_constant__ Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  State& S = d.st[b];
  if (S.phase >= 2) return;
  double pri = combine(d, b, R_PRI, 0);
  const double pc = combine(d, b, R_PRICOL, 0);
  if (!(pc <= pri)) pri = pc;
  const double npri = combine(d, b, R_NPRI, 0), dua = combine(d, b, R_DUA, 0), ndua = combine(d, b, R_NDUA, 0);
  S.dn = combine(d, b, R_DN, 1); S.pn = combine(d, b, R_PN, 1); S.obj = combine(d, b, R_OBJ, 1);
  double viol[4];
  for (int c = 0; c < 4; ++c) viol[c] = combine(d, b, R_VIOLJ + c, 0);
  double* sl = d.slab + (size_t)b * NRED;
  for (int s = 0; s <= R_OBJ; ++s) sl[s] = 0.0;
  for (int c = 0; c < 4; ++c) sl[R_VIOLJ + c] = 0.0;
  const int check = d.pb.check_every, maxit = d.pb.max_admm_iter;
  S.qp_it += check;
  scp_b200_record& r = d.rec[b];
  r.admm_iterations += check;
  if (S.phase == 1) r.cand_row_iters += 0.5 * S.ncand * (double)check;
  S.pri = pri; S.dua = dua;
  const double ea = d.pb.eps_abs, er = d.pb.eps_rel;
  const int nan = !(pri == pri) || !(dua == dua);
  int solved = !nan && pri <= ea + er * npri && dua <= ea + er * ndua;
  if (solved && S.on_mask != 15) {
    // the iterate solves the QP without the box-row classes that are outside the ADMM: a violated class joins
    // (its v already equals A x, i.e. multiplier 0) and the same subproblem continues warm
    int add = 0;
    for (int c = 0; c < 4; ++c) if (!(S.on_mask & (1 << c)) && viol[c] > ea + er * npri) add |= 1 << c;
    if (add) {
      S.on_mask |= add; S.flags |= FL_FACTOR | FL_RESET; S.reset_mask = add; solved = 0;
      S.it_mark = S.qp_it; S.pri_mark = INFINITY;
      r.rebuilds++;
    }
  }
  int stalled = 0;
  if (!solved && !nan && d.pb.stall_window > 0 && S.qp_it - S.it_mark >= d.pb.stall_window) {
    if (pri > 1e-3 * (1.0 + npri) && pri > 0.8 * S.pri_mark) stalled = 1;
    else { S.it_mark = S.qp_it; S.pri_mark = pri; }
  }
  if (solved || nan || stalled || S.qp_it >= maxit) {
    S.qp_solved = solved; S.stalled = stalled;
    const double full = (double)d.N * (double)(d.N - 1) * (double)(d.K - 1);
    S.flags |= FL_SCAN | (S.phase == 0 ? FL_GATE : ((solved && S.ncand < full) ? FL_VERIFY : 0));
  } else if (d.pb.adapt_every > 0 && S.qp_it % d.pb.adapt_every == 0) {
    double est = sqrt((pri / fmax(npri, 1e-12)) / fmax(dua / fmax(ndua, 1e-12), 1e-12));
    if (est > 5.0 || est < 0.2) {
      est = clampd(est, 1e-2, 1e2);
      const double nrho = clampd(S.rho * est, 1e-6, 1e6);
      S.est = nrho / S.rho; S.rho = nrho;
      S.flags |= FL_RESCALE | FL_FACTOR;
    }
  }
}

// after the pair scan: end of a subproblem -> gate (scp.py:144), candidate enlargement, or the SCP bookkeeping
// of scp.py:157-166
__global__ void k_control2(const __grid_constant__ Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  State& S = d.st[b];
  if (S.phase >= 2 || !(S.flags & FL_SCAN)) return;
  const double minsep = combine(d, b, R_MINSEP, 2), first = combine(d, b, R_FIRST, 2);
  const double bad = combine(d, b, R_BAD, 1), maxd = combine(d, b, R_MAXD, 0), over = combine(d, b, R_OVER, 0);
  double* sl = d.slab + (size_t)b * NRED;
  sl[R_MINSEP] = INFINITY; sl[R_FIRST] = INFINITY; sl[R_BAD] = 0.0; sl[R_MAXD] = 0.0; sl[R_OVER] = 0.0;
  scp_b200_record& r = d.rec[b];
  if (over > 0.0) {
    // the verification found a violated row that does not fit the per-(step, agent) candidate capacity: the iterate is
    // NOT the minimiser of the full QP (the reference carries every pair row, scp.py:487-552) -- reported, never silent
    r.reserved2 |= 4;
    if (bad == 0.0) S.qp_solved = 0;
  }
  S.minsep = minsep;
  r.pri_res = S.pri; r.dua_res = S.dua;
  int finish = 0, next_iter = 0;
  if (S.phase == 0) {
    if (!S.qp_solved) { r.status = SCP_B200_STATUS_INITIAL_QP_FAILED; r.qp_unsolved++; finish = 1; }
    const int feasible = !(first < INFINITY);
    r.initial_feasible = feasible;
    if (!feasible) {
      const int N = d.N, K = d.K;
      const long long np = (long long)N * (N - 1) / 2, fl = (long long)first;
      const int k = (int)(fl / np);
      long long p = fl - (long long)k * np;
      int i = 0;
      while (p >= N - 1 - i) { p -= N - 1 - i; ++i; }
      const int j = i + 1 + (int)p;
      r.first_violation[0] = k; r.first_violation[1] = i; r.first_violation[2] = j;
      const double* Pc = d.P + (size_t)b * d.Qs * K;
      const double dx = Pc[(size_t)(2 * i) * K + k] - Pc[(size_t)(2 * j) * K + k], dy = Pc[(size_t)(2 * i + 1) * K + k] - Pc[(size_t)(2 * j + 1) * K + k];
      r.first_violation_dist = sqrt(dx * dx + dy * dy);
      if (k == 0 && r.status == SCP_B200_STATUS_OK) r.status = SCP_B200_STATUS_START_TOO_CLOSE;
    }
    if (feasible || d.pb.max_scp_iter <= 0) finish = 1;
    if (!finish) next_iter = 1;
  } else {
    if (S.qp_solved && (S.flags & FL_VERIFY) && bad > 0.0 && S.attempt < 20) {
      // candidate rows were appended: same subproblem, warm state, operator rebuilt only if the copy count grew
      r.rebuilds++; S.attempt++;
      const int copies = (int)maxd > S.copies ? (int)maxd : S.copies;
      if (copies != S.copies) S.flags |= FL_FACTOR;
      S.copies = copies; S.ncand += bad;
      if (copies > r.max_copies) r.max_copies = copies;
      start_qp(S, d);
    } else {
      if (!S.qp_solved) { r.qp_unsolved++; if (S.stalled) r.qp_infeasible++; }
      const double rel = sqrt(S.dn) / sqrt(S.pn);
      if (S.scp_it < SCP_B200_MAX_SCP_ITER) r.rel_step[S.scp_it] = rel;
      const int conv = rel <= d.pb.scp_tolerance;
      S.scp_it++;
      r.scp_iterations = S.scp_it; r.converged = conv;
      if (conv || S.scp_it >= d.pb.max_scp_iter) finish = 1; else next_iter = 1;
    }
  }
  if (finish) {
    r.min_separation = minsep; r.objective = S.obj;
    S.flags |= FL_FINISH;
  } else if (next_iter) {
    S.phase = 1; S.rho = d.pb.rho0; S.margin = d.pb.cand_margin; S.attempt = 0; S.have_state = 0;
    S.flags |= FL_SNAPSHOT | FL_BUILD | FL_RESET; S.reset_mask = 15;
    start_qp(S, d);
  }
}

// after the candidate build: copies (the operator's collision weight) and whether the operator must be rebuilt
__global__ void k_control3(const __grid_constant__ Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  State& S = d.st[b];
  if (S.phase >= 2 || !(S.flags & FL_BUILD)) return;
  const int copies = (int)combine(d, b, R_COPIES, 0);
  const double ncand = combine(d, b, R_NCAND, 1), over = combine(d, b, R_OVER, 0);
  double* sl = d.slab + (size_t)b * NRED;
  sl[R_COPIES] = 0.0; sl[R_NCAND] = 0.0; sl[R_OVER] = 0.0;
  scp_b200_record& r = d.rec[b];
  if (over > 0.0) r.reserved2 |= 4;                 // candidate capacity exceeded: rows were dropped (reported by the host)
  if (!S.have_state || copies != S.copies) S.flags |= FL_FACTOR;
  S.copies = copies; S.ncand = ncand; S.have_state = 1;
  if (copies > r.max_copies) r.max_copies = copies;
}

__global__ void k_finalize(const __grid_constant__ Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B) return;
  State& S = d.st[b];
  if (S.phase < 2 && (S.flags & FL_FINISH)) { S.phase = 2; d.rec[b].device_ns = globaltimer_ns() - *d.t0ns; atomicAdd(d.done, 1); }
  S.flags = 0;
}

// ---------------------------------------------------------------------------------- host side
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi load_nccl_api() {
  NcclApi api;
  const char* env = getenv("SCP_B200_NCCL_LIB");
  void* h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);   // the copy PyTorch already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy) api.handle = h;
  return api;
}

NcclApi* nccl_api() {
  static NcclApi api = load_nccl_api();      // initialised once, thread safe (C++11 magic static)
  return api.handle ? &api : nullptr;
}

}  // namespace ss

struct scp_b200_stream {
  ss::Dev d;
  std::vector<void*> allocs;
  ncclComm_t comm = nullptr;
  int* h_done = nullptr;          // pinned, 2 entries
  cudaEvent_t ev[2] = {nullptr, nullptr}, t0 = nullptr, t1 = nullptr;
  void* tables = nullptr;
  std::vector<void*> ipc_opened;  // peer mappings (cudaIpcOpenMemHandle)
  cudaStream_t own = nullptr;     // the solver's stream (a captured graph cannot live on the legacy default stream)
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t gexec = nullptr;
  int graph_state = 0;            // 0: not tried, 1: one check period captured as a CUDA graph, -1: direct launches
  double* in4 = nullptr;          // the solver's own copy of p0, v0, pf, vf (kernel arguments stay constant -> graph reuse)
  scp_b200_record* rec_own = nullptr;
  void* io = nullptr;             // device staging of the host-buffer entry point
  size_t io_bytes = 0;
  int qpc = 8, nblk_q = 1, epl = 1, apc = 4, nblk_a = 1;
  size_t smem_iter = 0;
  size_t smem_axis = 0, smem_factor = 0;
  long long macro_steps = 0;
};

namespace {

#define SS_CUDA(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess) return scp_b200_set_error(100 + (int)_e, (std::string(#expr) + ": " + cudaGetErrorString(_e)).c_str()); \
  } while (0)
#define SS_NCCL(expr)                                                                                      \
  do {                                                                                                     \
    ncclResult_t _r = (expr);                                                                              \
    if (_r != ncclSuccess) return scp_b200_set_error(300 + (int)_r, (std::string(#expr) + ": " + (api->GetErrorString ? api->GetErrorString(_r) : "nccl error")).c_str()); \
  } while (0)

template <typename T>
int dev_alloc(scp_b200_stream* s, T** p, size_t n) {
  void* q = nullptr;
  SS_CUDA(cudaMalloc(&q, (n ? n : 1) * sizeof(T)));
  s->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}

template <int MODE>
void launch_axis(scp_b200_stream* s, cudaStream_t st) {
  const ss::Dev& d = s->d;
  dim3 grid(s->nblk_q, d.B);
  const size_t smem = (MODE <= 1 ? (size_t)d.K * d.K : 0) * sizeof(double) + (size_t)(ss::AX_THREADS / 32) * d.K * sizeof(double);
  switch (s->epl) {
    case 1: ss::k_axis<1, MODE><<<grid, ss::AX_THREADS, smem, st>>>(d, s->qpc); break;
    case 2: ss::k_axis<2, MODE><<<grid, ss::AX_THREADS, smem, st>>>(d, s->qpc); break;
    case 3: ss::k_axis<3, MODE><<<grid, ss::AX_THREADS, smem, st>>>(d, s->qpc); break;
    default: ss::k_axis<4, MODE><<<grid, ss::AX_THREADS, smem, st>>>(d, s->qpc); break;
  }
}

template <int CHK>
void launch_iter(scp_b200_stream* s, cudaStream_t st, int cur) {
  const ss::Dev& d = s->d;
  dim3 grid(s->nblk_a, d.B);
  switch (s->epl) {
    case 1: ss::k_iter<1, CHK><<<grid, ss::IT_THREADS, s->smem_iter, st>>>(d, s->apc, cur); break;
    case 2: ss::k_iter<2, CHK><<<grid, ss::IT_THREADS, s->smem_iter, st>>>(d, s->apc, cur); break;
    case 3: ss::k_iter<3, CHK><<<grid, ss::IT_THREADS, s->smem_iter, st>>>(d, s->apc, cur); break;
    default: ss::k_iter<4, CHK><<<grid, ss::IT_THREADS, s->smem_iter, st>>>(d, s->apc, cur); break;
  }
}

void launch_dual(scp_b200_stream* s, cudaStream_t st) {
  const ss::Dev& d = s->d;
  const int nq = 2 * (d.a_hi - d.a_lo);
  dim3 grid((nq * 32 + ss::AX_THREADS - 1) / ss::AX_THREADS, d.B);
  switch (s->epl) {
    case 1: ss::k_dual<1><<<grid, ss::AX_THREADS, 0, st>>>(d); break;
    case 2: ss::k_dual<2><<<grid, ss::AX_THREADS, 0, st>>>(d); break;
    case 3: ss::k_dual<3><<<grid, ss::AX_THREADS, 0, st>>>(d); break;
    default: ss::k_dual<4><<<grid, ss::AX_THREADS, 0, st>>>(d); break;
  }
}

template <typename F>
void set_axis_smem(F f, size_t smem) { cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }

}  // namespace

extern "C" {

// Defaults of the streaming solver: the reference's problem data (scp.py:32-74) with the ADMM settings that suit a
// solve without polish (fixed rho, OSQP's over-relaxation 1.6, residual tolerance 1e-4, a larger iteration cap).
void scp_b200_stream_default_problem(scp_b200_problem* prob, int n_agents, double time_horizon, double time_step,
                                     double min_distance) {
  scp_fill_default_problem(prob, n_agents, time_horizon, time_step, min_distance);
  prob->polish = 0;
  prob->adapt_every = 0;
  prob->relax_pct = 160;
  prob->eps_abs = 1e-4; prob->eps_rel = 1e-4;
  prob->max_admm_iter = 20000;
  prob->lazy_rows = 1;
  // measured (profiles/sweep_stream_r1.log): the best fixed rho falls with the horizon length, about (50/K)^2 --
  // K=50: 0.5..1, K=100: 0.2..0.3 (2.6-3x fewer iterations than rho = 1 on 100..1000 agents)
  const double kr = 50.0 / (double)(prob->n_steps > 1 ? prob->n_steps : 1);
  prob->rho0 = fmin(1.0, fmax(0.1, kr * kr));
  prob->stall_window = 1000;
  prob->check_every = 50;      // a check period costs ~15 small launches: fewer of them (a subproblem runs 10^3..10^4 iterations)
}

int scp_b200_nccl_unique_id(void* id128) {
  ss::NcclApi* api = ss::nccl_api();
  if (!api) return scp_b200_set_error(3, "libnccl.so.2 not found (set SCP_B200_NCCL_LIB)");
  ncclUniqueId id;
  SS_NCCL(api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

void scp_b200_stream_destroy(scp_b200_stream* s) {
  if (!s) return;
  for (void* p : s->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void* p : s->allocs) cudaFree(p);
  if (s->h_done) cudaFreeHost(s->h_done);
  if (s->io) cudaFree(s->io);
  if (s->gexec) cudaGraphExecDestroy(s->gexec);
  if (s->own) cudaStreamDestroy(s->own);
  for (auto e : {s->ev[0], s->ev[1], s->t0, s->t1, s->ev_in, s->ev_out}) if (e) cudaEventDestroy(e);
  if (s->comm) { ss::NcclApi* api = ss::nccl_api(); if (api) api->CommDestroy(s->comm); }
  delete s;
}

int scp_b200_stream_create(const scp_b200_problem* prob, int n_scenarios, int max_candidates, int rank, int world,
                           const void* nccl_id128, scp_b200_stream** out) {
  if (!prob || !out) return scp_b200_set_error(1, "null argument");
  *out = nullptr;
  const int N = prob->n_agents, K = prob->n_steps, B = n_scenarios;
  if (N < 1 || K < 2 || B < 1) return scp_b200_set_error(1, "bad sizes");
  if (K > 128) return scp_b200_set_error(1, "streaming solver: n_steps must be <= 128");
  if (B > 65535) return scp_b200_set_error(1, "streaming solver: at most 65535 scenarios per call");
  if (world < 1 || rank < 0 || rank >= world) return scp_b200_set_error(1, "bad rank/world");
  if (world > 1 && B != 1) return scp_b200_set_error(1, "agent sharding (world > 1) solves ONE scenario; shard batches by scenario instead");
  if (world > 1 && !nccl_id128) return scp_b200_set_error(1, "world > 1 needs an NCCL unique id");
  if (prob->check_every < 1 || prob->max_admm_iter < 1 || !(prob->time_step > 0)) return scp_b200_set_error(1, "bad ADMM settings");
  if (max_candidates <= 0) max_candidates = 16;
  if (max_candidates > ss::MAXC_MAX) max_candidates = ss::MAXC_MAX;
  if (max_candidates > N - 1) max_candidates = N > 1 ? N - 1 : 1;
  scp_b200_stream* s = new scp_b200_stream();
  ss::Dev& d = s->d;
  memset(&d, 0, sizeof(d));
  d.pb = *prob;
  d.pb.check_every = prob->check_every + (prob->check_every & 1);      // even: the position ping-pong ends each check period in buffer 0
  d.pb.max_admm_iter = ((prob->max_admm_iter + d.pb.check_every - 1) / d.pb.check_every) * d.pb.check_every;
  d.B = B; d.N = N; d.K = K; d.Q = 2 * N; d.G = world; d.rank = rank; d.maxc = max_candidates;
  const int nper = (N + world - 1) / world;
  d.Npad = nper * world; d.Qs = 2 * d.Npad;
  d.a_lo = rank * nper < N ? rank * nper : N;
  d.a_hi = (rank + 1) * nper < N ? (rank + 1) * nper : N;
  const int Nown = d.a_hi - d.a_lo;
  int rc = 0;
  auto fail = [&](int code) { scp_b200_stream_destroy(s); return code; };
  // tables: row weights and S'RcS as in scp_tables.h; the box-row operator is kept per class (it is linear in the
  // class weights) so that a scenario's operator can be assembled from the classes its ADMM carries
  {
    scp::HostTables t = scp::build_host_tables(d.pb);
    std::vector<double> blob(t.blob.begin() + (size_t)K * K, t.blob.end());          // [B2 | rj ra rv rp rc]
    for (int c = 0; c < 4; ++c) {
      scp_b200_problem pc = d.pb;
      if (c != 0) pc.w_jerk = 0.0;
      if (c != 1) pc.w_acc = 0.0;
      if (c != 2) pc.w_vel = 0.0;
      if (c != 3) pc.w_pos = 0.0;
      scp::HostTables tc = scp::build_host_tables(pc);
      blob.insert(blob.end(), tc.blob.begin(), tc.blob.begin() + (size_t)K * K);
    }
    double* tb = nullptr;
    if ((rc = dev_alloc(s, &tb, blob.size()))) return fail(rc);
    if (cudaMemcpy(tb, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) return fail(scp_b200_set_error(100, "table upload failed"));
    d.B2 = tb; d.rj = tb + (size_t)K * K; d.ra = d.rj + K; d.rv = d.ra + K; d.rp = d.rv + K; d.rc = d.rp + K;
    for (int c = 0; c < 4; ++c) d.Bc[c] = d.rc + K + (size_t)c * K * K;
  }
  const size_t QK = (size_t)B * d.Qs * K;
  double** arrs[] = {&d.x, &d.xprev, &d.va, &d.vj, &d.vv, &d.vp, &d.P, &d.P1, &d.Pbar, &d.F, &d.FY};
  for (double** a : arrs) if ((rc = dev_alloc(s, a, QK))) return fail(rc);
  if ((rc = dev_alloc(s, &d.mu, (size_t)B * d.Qs * 2))) return fail(rc);
  if ((rc = dev_alloc(s, &d.qsum, (size_t)B * d.Qs * 3))) return fail(rc);
  if ((rc = dev_alloc(s, &d.Nmat, (size_t)B * K * K))) return fail(rc);
  if ((rc = dev_alloc(s, &d.N0, (size_t)B * 2 * K))) return fail(rc);
  if ((rc = dev_alloc(s, &d.Qm, (size_t)B * 2 * K))) return fail(rc);
  if ((rc = dev_alloc(s, &d.gg, (size_t)B * 4))) return fail(rc);
  const size_t T = (size_t)B * (Nown > 0 ? Nown : 1) * K;
  if ((rc = dev_alloc(s, &d.cnt, T))) return fail(rc);
  if ((rc = dev_alloc(s, &d.cj, T * d.maxc))) return fail(rc);
  double** carr[] = {&d.cex, &d.cey, &d.cb, &d.lam, &d.lam1, &d.lamt};
  for (double** a : carr) if ((rc = dev_alloc(s, a, T * d.maxc))) return fail(rc);
  if ((rc = dev_alloc(s, &d.slab, (size_t)B * ss::NRED))) return fail(rc);
  if (world > 1) { if ((rc = dev_alloc(s, &d.gath, (size_t)world * B * ss::NRED))) return fail(rc); }
  else d.gath = d.slab;
  if ((rc = dev_alloc(s, &d.st, (size_t)B))) return fail(rc);
  if ((rc = dev_alloc(s, &d.done, 1))) return fail(rc);
  if ((rc = dev_alloc(s, &d.t0ns, 1))) return fail(rc);
  if ((rc = dev_alloc(s, &d.flag, 16))) return fail(rc);
  if (cudaMemset(d.flag, 0, 16 * sizeof(unsigned)) != cudaSuccess) return fail(scp_b200_set_error(100, "memset"));
  d.iter_no = d.flag + 8; d.cta_done = d.flag + 9; d.err = d.flag + 10;
  d.p2p = 0;
  const size_t OUT = (size_t)B * d.Npad * K * 2;
  if ((rc = dev_alloc(s, &d.acc, OUT))) return fail(rc);
  if ((rc = dev_alloc(s, &d.pos, OUT))) return fail(rc);
  if ((rc = dev_alloc(s, &d.vel, OUT))) return fail(rc);
  if ((rc = dev_alloc(s, &s->in4, (size_t)4 * B * d.Q))) return fail(rc);
  if ((rc = dev_alloc(s, &s->rec_own, (size_t)B))) return fail(rc);
  d.p0 = s->in4; d.v0 = s->in4 + (size_t)B * d.Q; d.pf = s->in4 + 2 * (size_t)B * d.Q; d.vf = s->in4 + 3 * (size_t)B * d.Q;
  d.rec = s->rec_own;
  if (cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking) != cudaSuccess) return fail(scp_b200_set_error(100, "stream"));
  for (cudaEvent_t* e : {&s->ev_in, &s->ev_out}) if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail(scp_b200_set_error(100, "event"));
  if (cudaMallocHost((void**)&s->h_done, 2 * sizeof(int)) != cudaSuccess) return fail(scp_b200_set_error(100, "pinned allocation failed"));
  for (cudaEvent_t* e : {&s->ev[0], &s->ev[1]}) if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail(scp_b200_set_error(100, "event"));
  for (cudaEvent_t* e : {&s->t0, &s->t1}) if (cudaEventCreate(e) != cudaSuccess) return fail(scp_b200_set_error(100, "event"));
  // launch geometry of the agent-axis kernel: about two CTAs per SM over the whole batch, >= 8 agent-axes per CTA
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nq = 2 * Nown;
  int nb = (2 * sms + B - 1) / B;
  const int nbmax = (nq + 7) / 8;
  if (nb > nbmax) nb = nbmax;
  if (nb < 1) nb = 1;
  int qpc = (nq + nb - 1) / nb;
  qpc = ((qpc + 7) / 8) * 8;
  if (qpc < 8) qpc = 8;
  s->qpc = qpc; s->nblk_q = nq > 0 ? (nq + qpc - 1) / qpc : 1;
  s->epl = (K + 31) / 32;
  s->smem_axis = ((size_t)K * K + (size_t)(ss::AX_THREADS / 32) * K) * sizeof(double);
  s->smem_factor = ((size_t)K * K + 4 * (size_t)K) * sizeof(double);
  set_axis_smem(ss::k_factor, s->smem_factor);
  {  // fused iteration kernel: one warp per agent, 4 warps per CTA, about two CTAs per SM over the whole batch
    const int nwarp = ss::IT_THREADS / 64;   // agents in flight per CTA (two warps each)
    int nba = (2 * sms + B - 1) / B;
    const int nbamax = (Nown + nwarp - 1) / nwarp;
    if (nba > nbamax) nba = nbamax;
    if (nba < 1) nba = 1;
    int apc = Nown > 0 ? (Nown + nba - 1) / nba : nwarp;
    apc = ((apc + nwarp - 1) / nwarp) * nwarp;
    s->apc = apc; s->nblk_a = Nown > 0 ? (Nown + apc - 1) / apc : 1;
    s->smem_iter = ((size_t)K * K + 4 + 2 * (size_t)(ss::IT_THREADS / 32) * (K + (K & 1))) * sizeof(double);   // operator, rhs rows, product rows
    set_axis_smem(ss::k_iter<1, 0>, s->smem_iter); set_axis_smem(ss::k_iter<1, 1>, s->smem_iter);
    set_axis_smem(ss::k_iter<2, 0>, s->smem_iter); set_axis_smem(ss::k_iter<2, 1>, s->smem_iter);
    set_axis_smem(ss::k_iter<3, 0>, s->smem_iter); set_axis_smem(ss::k_iter<3, 1>, s->smem_iter);
    set_axis_smem(ss::k_iter<4, 0>, s->smem_iter); set_axis_smem(ss::k_iter<4, 1>, s->smem_iter);
  }
  if (world > 1) {
    ss::NcclApi* api = ss::nccl_api();
    if (!api) return fail(scp_b200_set_error(3, "libnccl.so.2 not found (set SCP_B200_NCCL_LIB)"));
    ncclUniqueId id;
    memcpy(&id, nccl_id128, sizeof(id));
    ncclResult_t r = api->CommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) return fail(scp_b200_set_error(300 + (int)r, "ncclCommInitRank failed"));
  }
  *out = s;
  return 0;
}

// Peer-memory exchange for the agent-sharded solve: every rank exports IPC handles of its two position buffers and
// its flag array (3 x 64 bytes), the host all-gathers them, every rank maps its peers'.  After a successful connect
// k_iter stores the own position slice directly into every peer's buffer over NVLink and a flag per rank replaces
// the per-iteration ncclAllGather; the check-period slab exchange stays on NCCL.
int scp_b200_stream_ipc_handles(scp_b200_stream* s, void* out192) {
  if (!s || !out192) return scp_b200_set_error(1, "null argument");
  cudaIpcMemHandle_t h[3];
  SS_CUDA(cudaIpcGetMemHandle(&h[0], s->d.P));
  SS_CUDA(cudaIpcGetMemHandle(&h[1], s->d.P1));
  SS_CUDA(cudaIpcGetMemHandle(&h[2], s->d.flag));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(out192, h, sizeof(h));
  return 0;
}

int scp_b200_stream_ipc_connect(scp_b200_stream* s, const void* all_handles) {
  if (!s || !all_handles) return scp_b200_set_error(1, "null argument");
  ss::Dev& d = s->d;
  if (d.G < 2 || d.G > 8) return scp_b200_set_error(1, "peer exchange needs 2..8 ranks");
  if ((d.Npad / d.G) * (d.G - 1) >= d.N) return scp_b200_set_error(1, "peer exchange needs agents on every rank");
  const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
  for (int g = 0; g < d.G; ++g) {
    if (g == d.rank) { d.peerP[g] = d.P; d.peerP1[g] = d.P1; d.peerFlag[g] = d.flag; continue; }
    void* p[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; ++k) {
      cudaError_t e = cudaIpcOpenMemHandle(&p[k], h[3 * g + k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return scp_b200_set_error(100 + (int)e, (std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)).c_str());
      s->ipc_opened.push_back(p[k]);
    }
    d.peerP[g] = (double*)p[0]; d.peerP1[g] = (double*)p[1]; d.peerFlag[g] = (unsigned*)p[2];
  }
  d.p2p = 1;
  if (s->gexec) { cudaGraphExecDestroy(s->gexec); s->gexec = nullptr; }
  s->graph_state = 0;      // kernel arguments changed: capture again
  return 0;
}

// The whole of SCP.generate_trajectories (scp.py:131-180) for the solver's B scenarios (or for ONE scenario whose
// agents are sharded over `world` ranks: every rank calls this with the same inputs).  Blocks until finished.
int scp_b200_stream_solve(scp_b200_stream* s, const double* d_p0, const double* d_v0, const double* d_pf, const double* d_vf,
                          double* d_acc, double* d_pos, double* d_vel, scp_b200_record* d_records, void* stream,
                          float* device_ms, int64_t* macro_steps) {
  if (!s) return scp_b200_set_error(1, "null solver");
  ss::Dev& d = s->d;
  ss::NcclApi* api = ss::nccl_api();
  cudaStream_t user = (cudaStream_t)stream, st = s->own;
  // the solver works on its own stream, ordered after / before the caller's
  SS_CUDA(cudaEventRecord(s->ev_in, user));
  SS_CUDA(cudaStreamWaitEvent(st, s->ev_in, 0));
  {
    const size_t n2 = (size_t)d.B * d.Q * sizeof(double);
    SS_CUDA(cudaMemcpyAsync(s->in4, d_p0, n2, cudaMemcpyDeviceToDevice, st));
    SS_CUDA(cudaMemcpyAsync(s->in4 + (size_t)d.B * d.Q, d_v0, n2, cudaMemcpyDeviceToDevice, st));
    SS_CUDA(cudaMemcpyAsync(s->in4 + 2 * (size_t)d.B * d.Q, d_pf, n2, cudaMemcpyDeviceToDevice, st));
    SS_CUDA(cudaMemcpyAsync(s->in4 + 3 * (size_t)d.B * d.Q, d_vf, n2, cudaMemcpyDeviceToDevice, st));
  }
  const int B = d.B, K = d.K, Nown = d.a_hi - d.a_lo, G = d.G;
  const dim3 g_elem((d.Qs * K + 255) / 256 < 64 ? (d.Qs * K + 255) / 256 : 64, B);
  const dim3 g_ik((Nown * K + 255) / 256 > 0 ? (Nown * K + 255) / 256 : 1, B);
  const dim3 g_b((B + 127) / 128);
  const size_t pslice = (size_t)(d.Qs / G) * K;            // doubles per rank in the position all-gather (B == 1 when G > 1)
  auto gather_positions = [&](int buf) -> int {
    double* P = buf ? d.P1 : d.P;
    if (G > 1 && !d.p2p) SS_NCCL(api->AllGather(P + (size_t)d.rank * pslice, P, pslice, ncclFloat64, s->comm, st));
    return 0;
  };
  auto exchange = [&]() -> int {
    if (G > 1) SS_NCCL(api->AllGather(d.slab, d.gath, (size_t)B * ss::NRED, ncclFloat64, s->comm, st));
    return 0;
  };
  int rc0 = 0;
  SS_CUDA(cudaEventRecord(s->t0, st));
  ss::k_init<<<g_elem, 256, 0, st>>>(d);
  ss::k_factor<<<B, 512, s->smem_factor, st>>>(d);
  ss::k_finalize<<<g_b, 128, 0, st>>>(d);
  SS_CUDA(cudaGetLastError());
  // peer exchange: no rank may push positions into a peer's buffers before that peer's k_init has run
  if (d.p2p && (rc0 = exchange())) return rc0;
  const int check = d.pb.check_every;
  const long long max_macros = (long long)(d.pb.max_scp_iter + 2) * 22 * (d.pb.max_admm_iter / check + 2);
  s->h_done[0] = s->h_done[1] = 0;
  int rc = 0;
  auto body = [&]() -> int {
    // iteration it reads positions from buffer (it-1)&1 and writes buffer it&1; check is even, so the check
    // iteration leaves the current positions in buffer 0 (d.P), where every other kernel reads them
    for (int it = 1; it < check; ++it) {
      launch_iter<0>(s, st, (it - 1) & 1);
      if ((rc = gather_positions(it & 1))) return rc;
    }
    launch_iter<1>(s, st, (check - 1) & 1);
    if ((rc = gather_positions(check & 1))) return rc;
    if (d.p2p) ss::k_wait<<<1, 32, 0, st>>>(d);
    launch_dual(s, st);
    ss::k_local_sums<<<(B * 32 + 127) / 128, 128, 0, st>>>(d);
    if ((rc = exchange())) return rc;
    ss::k_control1<<<g_b, 128, 0, st>>>(d);
    ss::k_scan<<<g_ik, 256, 0, st>>>(d);
    if ((rc = exchange())) return rc;
    ss::k_control2<<<g_b, 128, 0, st>>>(d);
    ss::k_snapshot<<<g_elem, 256, 0, st>>>(d);
    ss::k_build<<<g_ik, 256, 0, st>>>(d);
    if ((rc = exchange())) return rc;
    ss::k_control3<<<g_b, 128, 0, st>>>(d);
    ss::k_rescale<<<g_elem, 256, 0, st>>>(d);
    launch_axis<2>(s, st);
    ss::k_factor<<<B, 512, s->smem_factor, st>>>(d);
    ss::k_output<<<dim3((2 * Nown + 127) / 128 > 0 ? (2 * Nown + 127) / 128 : 1, B), 128, 0, st>>>(d);
    ss::k_finalize<<<g_b, 128, 0, st>>>(d);
    return 0;
  };
  // one check period (check_every ADMM iterations + the control sequence) is captured once as a CUDA graph and
  // replayed: ~90 launches per period cost one graph launch; SCP_B200_STREAM_GRAPH=0 keeps direct launches
  if (s->graph_state == 0) {
    const char* env = getenv("SCP_B200_STREAM_GRAPH");
    s->graph_state = -1;
    if (!(env && env[0] == '0')) {
      if (G > 1) {   // NCCL sets up its connections on the first collective: keep that out of the capture
        if ((rc = gather_positions(0))) return rc;
        if ((rc = exchange())) return rc;
      }
      SS_CUDA(cudaStreamSynchronize(st));
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int brc = body();
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        if (brc == 0 && ce == cudaSuccess && graph && cudaGraphInstantiate(&s->gexec, graph, 0) == cudaSuccess) s->graph_state = 1;
        if (graph) cudaGraphDestroy(graph);
      }
      cudaGetLastError();
    }
  }
  auto macro = [&](long long m) -> int {
    if (s->graph_state == 1) SS_CUDA(cudaGraphLaunch(s->gexec, st));
    else if ((rc = body())) return rc;
    SS_CUDA(cudaMemcpyAsync(&s->h_done[m & 1], d.done, sizeof(int), cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaEventRecord(s->ev[m & 1], st));
    return 0;
  };
  long long m = 0;
  if ((rc = macro(0))) return rc;
  for (;;) {
    if ((rc = macro(m + 1))) return rc;
    SS_CUDA(cudaEventSynchronize(s->ev[m & 1]));
    if (s->h_done[m & 1] >= B) break;
    ++m;
    if (m > max_macros) return scp_b200_set_error(4, "streaming solver: macro step limit reached");
  }
  s->macro_steps = m + 1;
  // outputs: own agents were written by k_output; a sharded solve gathers the agent blocks (agent-major layout)
  const size_t oslice = (size_t)(d.Npad / G) * K * 2;
  if (G > 1) {
    for (double* o : {d.acc, d.pos, d.vel}) SS_NCCL(api->AllGather(o + (size_t)d.rank * oslice, o, oslice, ncclFloat64, s->comm, st));
  }
  const size_t row = (size_t)d.N * K * 2 * sizeof(double);
  if (d.Npad == d.N) {
    if (d_acc) SS_CUDA(cudaMemcpyAsync(d_acc, d.acc, row * B, cudaMemcpyDeviceToDevice, st));
    if (d_pos) SS_CUDA(cudaMemcpyAsync(d_pos, d.pos, row * B, cudaMemcpyDeviceToDevice, st));
    if (d_vel) SS_CUDA(cudaMemcpyAsync(d_vel, d.vel, row * B, cudaMemcpyDeviceToDevice, st));
  } else {
    const size_t prow = (size_t)d.Npad * K * 2 * sizeof(double);
    if (d_acc) SS_CUDA(cudaMemcpy2DAsync(d_acc, row, d.acc, prow, row, B, cudaMemcpyDeviceToDevice, st));
    if (d_pos) SS_CUDA(cudaMemcpy2DAsync(d_pos, row, d.pos, prow, row, B, cudaMemcpyDeviceToDevice, st));
    if (d_vel) SS_CUDA(cudaMemcpy2DAsync(d_vel, row, d.vel, prow, row, B, cudaMemcpyDeviceToDevice, st));
  }
  if (d_records) SS_CUDA(cudaMemcpyAsync(d_records, s->rec_own, (size_t)B * sizeof(scp_b200_record), cudaMemcpyDeviceToDevice, st));
  SS_CUDA(cudaEventRecord(s->t1, st));
  SS_CUDA(cudaEventRecord(s->ev_out, st));
  SS_CUDA(cudaStreamWaitEvent(user, s->ev_out, 0));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  if (d.p2p) {
    unsigned e = 0;
    SS_CUDA(cudaMemcpy(&e, d.err, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) return scp_b200_set_error(5, "peer exchange timed out waiting for a rank");
  }
  if (device_ms) SS_CUDA(cudaEventElapsedTime(device_ms, s->t0, s->t1));
  if (macro_steps) *macro_steps = s->macro_steps;
  return 0;
}

// Same with host buffers (what a non-PyTorch caller or the reference's SCP class binds): copies in, solves, copies out.
int scp_b200_stream_solve_host(scp_b200_stream* s, const double* h_p0, const double* h_v0, const double* h_pf,
                               const double* h_vf, double* h_acc, double* h_pos, double* h_vel,
                               scp_b200_record* h_records, float* device_ms, int64_t* macro_steps) {
  if (!s) return scp_b200_set_error(1, "null solver");
  const ss::Dev& d = s->d;
  const size_t n2 = (size_t)d.B * d.N * 2 * sizeof(double), n3 = (size_t)d.B * d.N * d.K * 2 * sizeof(double);
  const size_t nr = (size_t)d.B * sizeof(scp_b200_record);
  const size_t need = 4 * n2 + 3 * n3 + nr + 256;
  if (need > s->io_bytes) {
    if (s->io) cudaFree(s->io);
    s->io = nullptr; s->io_bytes = 0;
    SS_CUDA(cudaMalloc(&s->io, need));
    s->io_bytes = need;
  }
  char* io = (char*)s->io;
  double *p0 = (double*)io, *v0 = (double*)(io + n2), *pf = (double*)(io + 2 * n2), *vf = (double*)(io + 3 * n2);
  double *acc = (double*)(io + 4 * n2), *pos = (double*)(io + 4 * n2 + n3), *vel = (double*)(io + 4 * n2 + 2 * n3);
  scp_b200_record* rec = (scp_b200_record*)(io + 4 * n2 + 3 * n3);
  SS_CUDA(cudaMemcpy(p0, h_p0, n2, cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(v0, h_v0, n2, cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(pf, h_pf, n2, cudaMemcpyHostToDevice));
  SS_CUDA(cudaMemcpy(vf, h_vf, n2, cudaMemcpyHostToDevice));
  if (int rc = scp_b200_stream_solve(s, p0, v0, pf, vf, acc, pos, vel, rec, nullptr, device_ms, macro_steps)) return rc;
  if (h_acc) SS_CUDA(cudaMemcpy(h_acc, acc, n3, cudaMemcpyDeviceToHost));
  if (h_pos) SS_CUDA(cudaMemcpy(h_pos, pos, n3, cudaMemcpyDeviceToHost));
  if (h_vel) SS_CUDA(cudaMemcpy(h_vel, vel, n3, cudaMemcpyDeviceToHost));
  if (h_records) SS_CUDA(cudaMemcpy(h_records, rec, nr, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
