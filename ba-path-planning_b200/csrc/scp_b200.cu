// scp_b200.cu -- CUDA kernels (sm_100a) and the C ABI of include/scp_b200.h.
//
// Kernels
//   scp_solve_kernel      persistent, one CTA per resident scenario, scenarios
//                         handed out by an atomic counter; runs the whole SCP
//                         loop of scp.py:131-180 on device (scp_device.inl).
//   scp_reconstruct_kernel  scp.py:371-397 / 559-595 as warp-level prefix scans.
//   scp_linearize_kernel    scp.py:453-557 matrix free + scp.py:597-615 reduction.
// No CPU fallback: every entry point needs a CUDA device.

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/scp_b200.h"
#include "scp_defaults.h"
#include "scp_device.inl"
#include "scp_tables.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(100 + (int)_e, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
  } while (0)

constexpr int SOLVE_THREADS = 512;
constexpr size_t SMEM_BASE = (4 * scp::RED + scp::SH_EXTRA) * sizeof(double);   // scp::sh_doubles(RED)
constexpr int QUEUE_CAP = 1 << 20;                        // queue slots per launch: QLEVELS queues of QUEUE_CAP / QLEVELS
constexpr int QLEVELS = 4;
constexpr size_t HEADER_BYTES = 256 + (size_t)QUEUE_CAP * sizeof(int);
constexpr size_t TEAM_SCRATCH_BYTES = (size_t)24 << 20;   // team-wide reduction columns for the cooperative kernel
constexpr size_t SMEM_NMAT_LIMIT = 96 * 1024;
constexpr size_t SMEM_TOTAL_LIMIT = 227 * 1024;

size_t nmat_smem_bytes(int K) {
  size_t b = (size_t)K * K * sizeof(double);
  return b <= SMEM_NMAT_LIMIT ? b : 0;
}

// Shared-memory plan of scp_solve_kernel: base scratch, the K x K operator when it fits, the scratch rows of the fused
// iteration (K <= 64), then as many hot per-agent-axis arrays as fit, in order of reuse (P and F are also read by the
// collision rows, then x and the four v arrays, rhs last).  When all eight fit (config 2) the iteration runs its
// operator product on the tensor pipe (admm_iter_mma); otherwise the warp-fused iteration keeps the product's
// operands in registers (measured at 50 / 100 agents: faster than staging rhs and x through L2).
size_t fused_scratch_doubles(int N, int K) {
  size_t a = (size_t)(SOLVE_THREADS / 32) * K, b = (size_t)2 * K + (size_t)4 * N;   // old per-warp rows | N0 (2K) + d (4N)
  return ((a > b ? a : b) + 1) & ~(size_t)1;
}
size_t plan_smem(int N, int K, int* hot_mask) {
  size_t smem = SMEM_BASE + nmat_smem_bytes(K);
  const bool fused = nmat_smem_bytes(K) && K <= 64;
  if (fused) smem += fused_scratch_doubles(N, K) * sizeof(double);
  const size_t arr = (size_t)2 * N * K * sizeof(double);
  int mask = 0;
  for (int o = 0; o < 8; ++o) {
    const int a = o;                                     // array index in {P, F, x, vp, vv, vj, va, rhs}
    if (smem + arr <= SMEM_TOTAL_LIMIT) { smem += arr; mask |= 1 << a; }
  }
  if (hot_mask) *hot_mask = mask;
  return smem;
}

size_t slot_bytes(const scp::Layout& L) { return L.n_double * sizeof(double) + L.n_int * sizeof(int); }

int validate(const scp_b200_problem* p) {
  if (!p) return fail(1, "null problem");
  if (p->n_agents < 1) return fail(1, "n_agents must be >= 1");
  if (p->n_steps < 2) return fail(1, "n_steps must be >= 2");
  if (2 * p->n_steps > scp::RED) return fail(1, "n_steps must be <= 512");
  if (!(p->time_step > 0)) return fail(1, "time_step must be > 0");
  if (p->max_scp_iter < 0 || p->max_scp_iter > SCP_B200_MAX_SCP_ITER) return fail(1, "max_scp_iter out of range");
  if (p->check_every < 1 || p->max_admm_iter < 1) return fail(1, "bad ADMM iteration settings");
  return 0;
}

}  // namespace

// error slot shared with the other translation units of the library (scp_stream.cu)
int scp_b200_set_error(int code, const char* msg) { return fail(code, msg ? msg : ""); }

// ---------------------------------------------------------------------------------- solver kernel
__global__ void __launch_bounds__(SOLVE_THREADS, 1)
scp_solve_kernel(const __grid_constant__ scp::Params g, int B, const double* __restrict__ p0,
                 const double* __restrict__ v0, const double* __restrict__ pf, const double* __restrict__ vf,
                 double* ws_d, int* ws_i, double* acc, double* pos, double* vel, scp_b200_record* rec,
                 unsigned int* counter, int* queue, int qstride, int resumable, int nmat_in_smem, int hot_mask) {
  extern __shared__ double smem[];
  __shared__ int s_b, s_fresh;
  scp::Ctx c;
  c.nthreads = blockDim.x;
  c.team = 1; c.tid0 = 0; c.np = blockDim.x < 512 ? blockDim.x : 512; c.rs = scp::RED; c.sh = smem;
  c.N = g.pb.n_agents; c.K = g.pb.n_steps; c.Q = 2 * c.N;
  c.g = &g;
  c.wd = ws_d + (size_t)blockIdx.x * g.L.n_double;
  c.wi = ws_i + (size_t)blockIdx.x * g.L.n_int;
  c.sm = smem;
  c.nmat = nullptr;
  c.nmat_in_smem = nmat_in_smem;
  {
    // shared-memory plan (decided on the host, see plan_smem): [reductions | operator N | per-warp rhs rows | hot arrays]
    const size_t QK = (size_t)c.Q * c.K;
    double* cur = smem + scp::sh_doubles(scp::RED) + (nmat_in_smem ? (size_t)c.K * c.K : 0);
    c.fused_rows = cur;
    const int fused_ok = nmat_in_smem && c.K <= 64;
    if (fused_ok) {
      const size_t a = (size_t)(blockDim.x >> 5) * c.K, b = (size_t)2 * c.K + (size_t)4 * c.N;
      cur += ((a > b ? a : b) + 1) & ~(size_t)1;
    }
    double* glob[8] = {c.wd + g.L.P, c.wd + g.L.F, c.wd + g.L.x, c.wd + g.L.vp, c.wd + g.L.vv, c.wd + g.L.vj, c.wd + g.L.va, c.wd + g.L.rhs};
    double* ptr[8];
    for (int o = 0; o < 8; ++o) {
      const int a = o;                                        // same order as plan_smem
      if (hot_mask & (1 << a)) { ptr[a] = cur; cur += QK; } else ptr[a] = glob[a];
    }
    c.all_hot = (hot_mask & 0xFF) == 0xFF;
    c.mma_ok = c.all_hot;
    c.a_P = ptr[0]; c.a_F = ptr[1]; c.a_x = ptr[2]; c.a_vp = ptr[3]; c.a_vv = ptr[4]; c.a_vj = ptr[5]; c.a_va = ptr[6]; c.a_rhs = ptr[7];
    c.fused_epl = fused_ok ? 2 : 0;
    // region the polish factors in (hot arrays are assigned contiguously, in this order, rhs last and never parked)
    c.pol_smem = nullptr; c.pol_smem_doubles = 0;
    for (int a = 0; a < 7; ++a) {
      const bool hot = (hot_mask >> a) & 1;
      c.hot_s[a] = hot ? ptr[a] : nullptr; c.hot_g[a] = glob[a];
      if (hot) { if (!c.pol_smem || ptr[a] < c.pol_smem) c.pol_smem = ptr[a]; c.pol_smem_doubles += QK; }
    }
  }
  // Work queue of quanta (one SCP iteration each).  Fresh scenarios (tickets 0..B-1) come first, so that every scenario
  // has shown its first quantum early; a scenario that is not finished after a quantum is pushed to one of QLEVELS
  // FIFO queues chosen by the service it has attained so far relative to the mean first quantum (< 1x, 1-2x, 2-4x,
  // > 4x), and workers pop from the highest non-empty level: the scenarios that have already cost the most -- with the
  // heavy-tailed solve times of this workload the ones with the most left to do -- run without waiting, instead of
  // taking round-robin turns and finishing long after the rest (round 1: step 1096 ms against a balanced 905 ms on one
  // GPU, 1545 ms on a rank whose longest scenario started late).
  // header (32-bit words): [0] fresh tickets, [2] finished scenarios, [8+l] pops of level l, [16+l] pushes of level l;
  //         64-bit words at byte 128: sum of first-quantum cycles, count.
  unsigned long long* stat64 = (unsigned long long*)(counter + 32);
  for (;;) {
    if (threadIdx.x == 0) {
      int b = -1, fresh = 0;
      if (*(volatile unsigned*)&counter[0] < (unsigned)B) {
        const unsigned h = atomicAdd(&counter[0], 1u);
        if (h < (unsigned)B) { b = (int)h; fresh = 1; }
      }
      if (b < 0 && resumable) {
        for (;;) {
          for (int l = QLEVELS - 1; l >= 0 && b < 0; --l) {
            const unsigned hd = *(volatile unsigned*)&counter[8 + l], tl = *(volatile unsigned*)&counter[16 + l];
            if (hd < tl && atomicCAS(&counter[8 + l], hd, hd + 1u) == hd) {
              volatile int* slot = queue + (size_t)l * qstride + hd;
              int v;
              while ((v = *slot) < 0) __nanosleep(100);       // the pusher has reserved the slot and is writing it
              b = v;
            }
          }
          if (b >= 0) break;
          // nothing queued: this CTA retires.  A scenario still in flight is held by a CTA that pops again after
          // pushing it, so queued work is never orphaned; the SM goes to the next launch (consecutive batches on
          // different streams overlap: the tail of one batch runs under the head of the next)
          break;
        }
      }
      s_b = b; s_fresh = fresh;
    }
    __syncthreads();
    const int b = s_b, fresh = s_fresh;
    __syncthreads();
    if (b < 0) break;
    __threadfence();                                   // another SM may have written this scenario's record / iterate
    const size_t s2 = (size_t)b * c.N * 2, s3 = (size_t)b * c.N * c.K * 2;
    c.p0 = p0 + s2; c.v0 = v0 + s2; c.pf = pf + s2; c.vf = vf + s2;
    c.acc = acc + s3; c.pos = pos + s3; c.vel = vel + s3; c.rec = rec + b;
    const int done = scp::solve_scenario(c, resumable ? (fresh ? 2 : 1) : 0);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long att = (unsigned long long)c.rec->cycles_total;
      if (fresh) { atomicAdd(&stat64[0], att); atomicAdd(&stat64[1], 1ull); }
      if (done) atomicAdd(&counter[2], 1u);
      else {
        const unsigned long long n1 = *(volatile unsigned long long*)&stat64[1], s1 = *(volatile unsigned long long*)&stat64[0];
        const double mean1 = n1 ? (double)s1 / (double)n1 : 1.0, ratio = (double)att / mean1;
        const int l = ratio < 1.0 ? 0 : (ratio < 2.0 ? 1 : (ratio < 4.0 ? 2 : 3));
        const unsigned t = atomicAdd(&counter[16 + l], 1u);
        if (t < (unsigned)qstride) { __threadfence(); atomicExch(queue + (size_t)l * qstride + t, b); }
      }
    }
  }
}

// ---------------------------------------------------------------------------------- team kernel
// One large scenario at a time solved by the WHOLE cooperative grid (one CTA per SM): the same phase
// code, with team-global thread ids, grid-wide barriers and the reduction columns in global memory.
// Used when there are too few scenarios to fill the GPU with one CTA each (config 3: one 200-agent scenario).
__global__ void __launch_bounds__(SOLVE_THREADS, 1)
scp_solve_team_kernel(const __grid_constant__ scp::Params g, int B, const double* __restrict__ p0,
                      const double* __restrict__ v0, const double* __restrict__ pf, const double* __restrict__ vf,
                      double* ws_d, int* ws_i, double* team_scratch, double* acc, double* pos, double* vel,
                      scp_b200_record* rec) {
  extern __shared__ double smem[];
  scp::Ctx c;
  c.nthreads = gridDim.x * blockDim.x;
  c.team = gridDim.x; c.tid0 = blockIdx.x * blockDim.x; c.np = 512; c.rs = c.nthreads; c.sh = team_scratch;
  c.N = g.pb.n_agents; c.K = g.pb.n_steps; c.Q = 2 * c.N;
  c.g = &g;
  c.wd = ws_d; c.wi = ws_i;
  c.sm = smem;
  c.nmat = nullptr; c.nmat_in_smem = 0;
  c.fused_epl = 0; c.fused_rows = nullptr; c.all_hot = 0; c.mma_ok = 0;
  c.a_P = c.wd + g.L.P; c.a_F = c.wd + g.L.F; c.a_x = c.wd + g.L.x; c.a_vp = c.wd + g.L.vp;
  c.a_vv = c.wd + g.L.vv; c.a_vj = c.wd + g.L.vj; c.a_va = c.wd + g.L.va; c.a_rhs = c.wd + g.L.rhs;
  c.pol_smem = nullptr; c.pol_smem_doubles = 0;
  for (int a = 0; a < 7; ++a) { c.hot_s[a] = nullptr; c.hot_g[a] = nullptr; }
  for (int b = 0; b < B; ++b) {
    const size_t s2 = (size_t)b * c.N * 2, s3 = (size_t)b * c.N * c.K * 2;
    c.p0 = p0 + s2; c.v0 = v0 + s2; c.pf = pf + s2; c.vf = vf + s2;
    c.acc = acc + s3; c.pos = pos + s3; c.vel = vel + s3; c.rec = rec + b;
    scp::solve_scenario(c, 0);
    scp_team_sync(c.team);
  }
}

// ---------------------------------------------------------------------------------- reconstruct
// One warp per (scenario, agent); lanes stride over k with double2 (x,y) loads, the
// running sums c1 = cumsum(a), c2 = cumsum(c1) are carried by a 2-state warp scan:
//   v[k+1] = v0 + h c1[k],  p[k+1] = p0 + h (k+1) v0 + h^2 (c2[k] - c1[k]/2)   
__global__ void __launch_bounds__(256)
scp_reconstruct_kernel(const double2* __restrict__ acc, const double2* __restrict__ p0,
                       const double2* __restrict__ v0, int BN, int K, double h, double2* __restrict__ pos,
                       double2* __restrict__ vel) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= BN) return;
  const double2* a = acc + (size_t)warp * K;
  double2* po = pos + (size_t)warp * K;
  double2* ve = vel + (size_t)warp * K;
  const double2 P0 = p0[warp], V0 = v0[warp];
  if (lane == 0) { po[0] = P0; ve[0] = V0; }
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;   // carries
  for (int base = 0; base < K - 1; base += 32) {
    const int k = base + lane;
    double2 v = make_double2(0.0, 0.0);
    if (k < K - 1) v = a[k];
    double s1x = v.x, s1y = v.y, s2x = v.x, s2y = v.y;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      double o1x = __shfl_up_sync(0xffffffffu, s1x, d), o1y = __shfl_up_sync(0xffffffffu, s1y, d);
      double o2x = __shfl_up_sync(0xffffffffu, s2x, d), o2y = __shfl_up_sync(0xffffffffu, s2y, d);
      if (lane >= d) {
        s2x += o2x + (double)d * o1x; s2y += o2y + (double)d * o1y;
        s1x += o1x; s1y += o1y;
      }
    }
    const double n = (double)(lane + 1);
    const double t2x = c2x + n * c1x + s2x, t2y = c2y + n * c1y + s2y;
    const double t1x = c1x + s1x, t1y = c1y + s1y;
    if (k < K - 1) {
      const double kk = (double)(k + 1);
      ve[k + 1] = make_double2(V0.x + h * t1x, V0.y + h * t1y);
      po[k + 1] = make_double2(P0.x + h * kk * V0.x + h * h * (t2x - 0.5 * t1x),
                               P0.y + h * kk * V0.y + h * h * (t2y - 0.5 * t1y));
    }
    c1x = __shfl_sync(0xffffffffu, t1x, 31); c1y = __shfl_sync(0xffffffffu, t1y, 31);
    c2x = __shfl_sync(0xffffffffu, t2x, 31); c2y = __shfl_sync(0xffffffffu, t2y, 31);
  }
}

// ---------------------------------------------------------------------------------- linearize
// CTA = (scenario b, tile of KT time steps, chunk of pair indices).  The positions of
// all agents for the tile are staged in shared memory ([kk][i] double2); threads then
// stream over rows p (pair index, the reference's i<j lexicographic order) writing
// eta (double2) and the bound coalesced, and reduce min distance / first violation
// with warp shuffles before one atomic per CTA.
constexpr int LIN_THREADS = 256;

__device__ __forceinline__ void pair_from_index(long long p, int N, int& i, int& j) {
  // p = i (2N - i - 1)/2 + (j - i - 1)
  const double t = 2.0 * N - 1.0;
  int ii = (int)floor((t - sqrt(t * t - 8.0 * (double)p)) * 0.5);
  if (ii < 0) ii = 0;
  while ((long long)ii * (2 * N - ii - 1) / 2 > p) --ii;
  while ((long long)(ii + 1) * (2 * N - ii - 2) / 2 <= p) ++ii;
  i = ii;
  j = (int)(p - (long long)ii * (2 * N - ii - 1) / 2) + ii + 1;
}

__global__ void __launch_bounds__(LIN_THREADS)
scp_linearize_kernel(const double2* __restrict__ pos, int N, int K, int KT, int nchunks, long long p_begin, long long p_end,
                     double R, double thr,
                     double2* __restrict__ eta, double* __restrict__ bound,
                     unsigned long long* __restrict__ minsep_bits, unsigned long long* __restrict__ first_row) {
  extern __shared__ double2 sp[];                       // [KT][N]
  const int b = blockIdx.z, kt = blockIdx.y, chunk = blockIdx.x;
  const int k0 = kt * KT, kn = min(KT, K - k0);
  const long long P = (long long)N * (N - 1) / 2;
  const double2* src = pos + (size_t)b * N * K;
  for (int e = threadIdx.x; e < N * kn; e += blockDim.x) {
    const int i = e / kn, kk = e - i * kn;              // kn contiguous double2 per agent
    sp[kk * N + i] = src[(size_t)i * K + k0 + kk];
  }
  __syncthreads();
  const long long Pr = p_end - p_begin;               // rows of this call (a rank's share of the pair indices)
  const long long per = (Pr + nchunks - 1) / nchunks;
  const long long pa = p_begin + (long long)chunk * per, pb = min(p_end, pa + per);
  double mn = INFINITY;
  unsigned long long fr = ~0ull;
  // each thread walks its pair indices p = pa + tid, pa + tid + T, ...; (i, j) is advanced incrementally
  int i0 = 0, j0 = 0;
  const long long pfirst = pa + threadIdx.x;
  if (pfirst < pb) pair_from_index(pfirst, N, i0, j0);
  const int T = blockDim.x;
  for (int kk = 0; kk < kn; ++kk) {
    const int k = k0 + kk;
    const double2* row = sp + kk * N;
    const size_t obase = ((size_t)b * K + k) * (size_t)Pr - (size_t)p_begin;   // output row index is p - p_begin
    int i = i0, j = j0;
    for (long long p = pfirst; p < pb; p += T) {
      const double2 a = row[i], c = row[j];
      const double dx = a.x - c.x, dy = a.y - c.y;
      // one square root serves both np.hypot (scp.py:501) and np.linalg.norm (scp.py:609): same value to 1 ulp
      double dist = sqrt(dx * dx + dy * dy);
      mn = fmin(mn, dist);
      if (dist < thr) { unsigned long long r = (unsigned long long)k * (unsigned long long)P + (unsigned long long)p; fr = r < fr ? r : fr; }
      if (eta) {
        double ex, ey;
        if (dist < 1e-6) { ex = 1.0; ey = 0.0; dist = 1.0; }   // deterministic stand-in for scp.py:503-507
        else { const double inv = 1.0 / dist; ex = dx * inv; ey = dy * inv; }
        eta[obase + p] = make_double2(ex, ey);
        bound[obase + p] = R + ((ex * dx + ey * dy) - dist);   // scp.py:547-549 without the p0/v0 shift (T3)
      }
      // advance (i, j) by T pairs in lexicographic i<j order
      j += T;
      while (j >= N) { const int over = j - N; ++i; j = i + 1 + over; if (i >= N - 1) break; }
    }
  }
  // warp shuffle reduction, then one atomic per warp (positive doubles order like their bit patterns)
  unsigned long long mb = (unsigned long long)__double_as_longlong(mn);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, mb, d);
    mb = o < mb ? o : mb;
    unsigned long long f = __shfl_xor_sync(0xffffffffu, fr, d);
    fr = f < fr ? f : fr;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(minsep_bits + b, mb);
    if (fr != ~0ull) atomicMin(first_row + b, fr);
  }
}

__global__ void scp_linearize_init_kernel(unsigned long long* minsep_bits, unsigned long long* first_row, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { minsep_bits[b] = (unsigned long long)__double_as_longlong(INFINITY); first_row[b] = ~0ull; }
}

__global__ void scp_linearize_finish_kernel(const unsigned long long* minsep_bits, const unsigned long long* first_row,
                                            int B, int N, double* minsep, int32_t* first) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (minsep) minsep[b] = (minsep_bits[b] == ~0ull) ? INFINITY : __longlong_as_double((long long)minsep_bits[b]);
  if (first) {
    const unsigned long long P = (unsigned long long)N * (N - 1) / 2;
    if (first_row[b] == ~0ull || P == 0) { first[3 * b] = first[3 * b + 1] = first[3 * b + 2] = -1; }
    else {
      int i, j;
      pair_from_index((long long)(first_row[b] % P), N, i, j);
      first[3 * b] = (int)(first_row[b] / P); first[3 * b + 1] = i; first[3 * b + 2] = j;
    }
  }
}

// ---------------------------------------------------------------------------------- fp64 peak probe
// 8 independent FMA chains per thread, 512 threads per CTA, 4 CTAs per SM worth of blocks: the roofline denominator of
// the solver kernel (its arithmetic is fp64 FMA / ADD / MNMX on the same pipe).
__global__ void __launch_bounds__(512) scp_fp64_probe_kernel(double* out, int iters) {
  double a[8];
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) a[ch] = 1.0 + 1e-9 * (double)(threadIdx.x + ch);
  const double m = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) a[ch] = fma(a[ch], m, b);
  }
  double s = 0.0;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) s += a[ch];
  if (s == 12345.678) out[0] = s;      // never true: keeps the chains alive
}

// ---------------------------------------------------------------------------------- C ABI
extern "C" {

int scp_b200_measure_fp64_peak(double* tflops_out) {
  if (!tflops_out) return fail(1, "null argument");
  int dev = 0, sms = 148;
  CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d = nullptr;
  CUDA_OK(cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
  const int iters = 20000, blocks = sms * 4;
  scp_fp64_probe_kernel<<<blocks, 512>>>(d, 200);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CUDA_OK(cudaEventRecord(e0));
    scp_fp64_probe_kernel<<<blocks, 512>>>(d, iters);
    CUDA_OK(cudaEventRecord(e1));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * (double)iters * 512.0 * (double)blocks / ((double)ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  CUDA_OK(cudaGetLastError());
  *tflops_out = best;
  return 0;
}

int scp_b200_abi_version(void) { return SCP_B200_ABI_VERSION; }
size_t scp_b200_sizeof_problem(void) { return sizeof(scp_b200_problem); }
size_t scp_b200_sizeof_record(void) { return sizeof(scp_b200_record); }

const char* scp_b200_last_error(void) { return g_err.c_str(); }

void scp_b200_default_problem(scp_b200_problem* prob, int n_agents, double time_horizon, double time_step,
                              double min_distance) {
  scp_fill_default_problem(prob, n_agents, time_horizon, time_step, min_distance);
}

size_t scp_b200_tables_bytes(const scp_b200_problem* prob) {
  return scp::tables_doubles(prob->n_steps) * sizeof(double);
}

int scp_b200_build_tables(const scp_b200_problem* prob, void* d_tables, void* stream) {
  if (int rc = validate(prob)) return rc;
  scp::HostTables t = scp::build_host_tables(*prob);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaMemcpyAsync(d_tables, t.blob.data(), t.blob.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaStreamSynchronize(st));   // t.blob is pageable and dies with this frame
  return 0;
}

size_t scp_b200_workspace_bytes(const scp_b200_problem* prob, int slots) {
  scp::Layout L = scp::make_layout(prob->n_agents, prob->n_steps);
  return slot_bytes(L) * (size_t)slots + HEADER_BYTES + TEAM_SCRATCH_BYTES;
}

int scp_b200_default_slots(const scp_b200_problem* prob) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  scp::Layout L = scp::make_layout(prob->n_agents, prob->n_steps);
  const size_t smem = plan_smem(prob->n_agents, prob->n_steps, nullptr);
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;                        // 4 x 512 threads = the SM's 2048
  size_t slots = (size_t)sms * per_sm;
  const size_t budget = (size_t)48 << 30;            // keep scratch well inside 180 GB
  while (slots > 1 && slot_bytes(L) * slots > budget) slots /= 2;
  return (int)slots;
}

int scp_b200_solve_batch(const scp_b200_problem* prob, int B, const double* d_p0, const double* d_v0,
                         const double* d_pf, const double* d_vf, const void* d_tables, void* d_workspace,
                         size_t workspace_bytes, int slots, double* d_acc, double* d_pos, double* d_vel,
                         scp_b200_record* d_records, void* stream) {
  if (int rc = validate(prob)) return rc;
  if (B <= 0) return 0;
  if (slots < 1) return fail(1, "slots must be >= 1");
  scp::Params g;
  g.pb = *prob;
  g.L = scp::make_layout(prob->n_agents, prob->n_steps);
  const int K = prob->n_steps;
  const double* tb = (const double*)d_tables;
  g.tb.B1 = tb; g.tb.B2 = tb + (size_t)K * K; g.tb.rj = tb + 2 * (size_t)K * K;
  g.tb.ra = g.tb.rj + K; g.tb.rv = g.tb.ra + K; g.tb.rp = g.tb.rv + K; g.tb.rc = g.tb.rp + K;
  const size_t need = slot_bytes(g.L) * (size_t)slots + HEADER_BYTES + TEAM_SCRATCH_BYTES;
  if (workspace_bytes < need) return fail(2, "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)d_workspace;
  unsigned int* counter = (unsigned int*)base;
  int* queue = (int*)(base + 256);
  double* ws_d = (double*)(base + HEADER_BYTES);
  int* ws_i = (int*)(base + HEADER_BYTES + g.L.n_double * sizeof(double) * (size_t)slots);
  CUDA_OK(cudaMemsetAsync(counter, 0, 256, st));
  // warm_duals keeps the multipliers of the previous subproblem in the slot scratch: a re-queued scenario may resume in
  // another slot, so that option runs every scenario start to finish in one slot
  const int qstride = B * (prob->max_scp_iter + 2);        // a scenario is pushed at most once per SCP iteration
  const int resumable = (!prob->warm_duals && (long long)qstride * QLEVELS <= (long long)QUEUE_CAP) ? 1 : 0;
  if (resumable) CUDA_OK(cudaMemsetAsync(queue, 0xFF, (size_t)qstride * QLEVELS * sizeof(int), st));
  int hot_mask = 0;
  const size_t nm = nmat_smem_bytes(K);
  const size_t smem = plan_smem(prob->n_agents, K, &hot_mask);
  CUDA_OK(cudaFuncSetAttribute(scp_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // few large scenarios: one scenario at a time on the whole GPU (cooperative grid); otherwise one CTA each
  int dev = 0, sms = 148, coop = 0;
  CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  const bool team_mode = coop && prob->team_mode != 1 && prob->team_mode != 3 && (prob->team_mode == 2 || (B <= 8 && (size_t)2 * prob->n_agents * K >= 8192));
  if (team_mode) {
    double* team_scratch = (double*)(base + HEADER_BYTES + slot_bytes(g.L) * (size_t)slots);
    if (scp::sh_doubles((size_t)sms * SOLVE_THREADS) * sizeof(double) > TEAM_SCRATCH_BYTES) return fail(3, "team scratch too small");
    void* args[] = {(void*)&g, (void*)&B, (void*)&d_p0, (void*)&d_v0, (void*)&d_pf, (void*)&d_vf, (void*)&ws_d, (void*)&ws_i,
                    (void*)&team_scratch, (void*)&d_acc, (void*)&d_pos, (void*)&d_vel, (void*)&d_records};
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)scp_solve_team_kernel, dim3(sms), dim3(SOLVE_THREADS), args, 1024, st));
  } else {
    const int grid = B < slots ? B : slots;
    scp_solve_kernel<<<grid, SOLVE_THREADS, smem, st>>>(g, B, d_p0, d_v0, d_pf, d_vf, ws_d, ws_i, d_acc, d_pos,
                                                          d_vel, d_records, counter, queue, qstride, resumable, nm ? 1 : 0, hot_mask);
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

namespace {
constexpr int MAX_DEVICES = 64;
constexpr int HOST_LANES = 2;
// Device buffers reused across scp_b200_solve_batch_host calls, per device.  Two LANES (stream + workspace + staging
// buffers each): two host threads calling at the same time run on different streams, so the tail of one batch (a few
// long scenarios, most SMs already retired) overlaps the head of the next; a single caller always gets lane 0.
struct HostLane {
  std::mutex mu;
  void* ws = nullptr; size_t ws_bytes = 0;
  void* io = nullptr; size_t io_bytes = 0;
  cudaStream_t stream = nullptr;
};
struct HostCache {
  std::mutex mu;        // guards the tables (shared by the lanes)
  void* tables = nullptr; size_t tables_bytes = 0; scp_b200_problem tables_for{}; bool tables_valid = false;
  std::vector<void*> retired;   // replaced tables another lane may still be reading
  HostLane lane[HOST_LANES];
};
HostCache g_cache[MAX_DEVICES];
}  // namespace

int scp_b200_solve_batch_host(const scp_b200_problem* prob, int B, const double* h_p0, const double* h_v0,
                              const double* h_pf, const double* h_vf, double* h_acc, double* h_pos,
                              double* h_vel, scp_b200_record* h_records, int device) {
  if (int rc = validate(prob)) return rc;
  if (B <= 0) return 0;
  if (device < 0 || device >= MAX_DEVICES) return fail(1, "device index out of range");
  HostCache& hc = g_cache[device];
  CUDA_OK(cudaSetDevice(device));
  // a free lane, else wait for lane 0
  int li = -1;
  for (int l = 0; l < HOST_LANES && li < 0; ++l) if (hc.lane[l].mu.try_lock()) li = l;
  if (li < 0) { hc.lane[0].mu.lock(); li = 0; }
  HostLane& ln = hc.lane[li];
  std::lock_guard<std::mutex> lane_lock(ln.mu, std::adopt_lock);
  if (!ln.stream) CUDA_OK(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
  const int N = prob->n_agents, K = prob->n_steps;
  {
    std::lock_guard<std::mutex> lock(hc.mu);
    const size_t tb = scp_b200_tables_bytes(prob);
    if (!hc.tables_valid || tb != hc.tables_bytes || memcmp(&hc.tables_for, prob, sizeof(*prob)) != 0) {
      // the other lane may still be reading the old tables: build the new ones in a fresh allocation and leave the old
      // one to be freed when it is replaced the next time (a problem change between calls is rare)
      std::vector<void*>& retired = hc.retired;
      void* fresh = nullptr;
      CUDA_OK(cudaMalloc(&fresh, tb));
      if (int rc = scp_b200_build_tables(prob, fresh, ln.stream)) { cudaFree(fresh); return rc; }
      if (hc.tables) retired.push_back(hc.tables);
      while (retired.size() > 2) { cudaFree(retired.front()); retired.erase(retired.begin()); }
      hc.tables = fresh; hc.tables_bytes = tb; hc.tables_for = *prob; hc.tables_valid = true;
    }
  }
  void* tables = hc.tables;
  int slots = scp_b200_default_slots(prob);
  if (slots > B) slots = B;
  const size_t wb = scp_b200_workspace_bytes(prob, slots);
  if (wb > ln.ws_bytes) { if (ln.ws) cudaFree(ln.ws); ln.ws = nullptr; ln.ws_bytes = 0; CUDA_OK(cudaMalloc(&ln.ws, wb)); ln.ws_bytes = wb; }
  const size_t n2 = (size_t)B * N * 2 * sizeof(double), n3 = (size_t)B * N * K * 2 * sizeof(double);
  const size_t nr = (size_t)B * sizeof(scp_b200_record);
  const size_t iob = 4 * n2 + 3 * n3 + nr + 1024;
  if (iob > ln.io_bytes) { if (ln.io) cudaFree(ln.io); ln.io = nullptr; ln.io_bytes = 0; CUDA_OK(cudaMalloc(&ln.io, iob)); ln.io_bytes = iob; }
  char* io = (char*)ln.io;
  double *d_p0 = (double*)io, *d_v0 = (double*)(io + n2), *d_pf = (double*)(io + 2 * n2), *d_vf = (double*)(io + 3 * n2);
  double *d_acc = (double*)(io + 4 * n2), *d_pos = (double*)(io + 4 * n2 + n3), *d_vel = (double*)(io + 4 * n2 + 2 * n3);
  scp_b200_record* d_rec = (scp_b200_record*)(io + 4 * n2 + 3 * n3);
  cudaStream_t st = ln.stream;
  CUDA_OK(cudaMemcpyAsync(d_p0, h_p0, n2, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_v0, h_v0, n2, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_pf, h_pf, n2, cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemcpyAsync(d_vf, h_vf, n2, cudaMemcpyHostToDevice, st));
  if (int rc = scp_b200_solve_batch(prob, B, d_p0, d_v0, d_pf, d_vf, tables, ln.ws, ln.ws_bytes, slots, d_acc,
                                    d_pos, d_vel, d_rec, st))
    return rc;
  if (h_acc) CUDA_OK(cudaMemcpyAsync(h_acc, d_acc, n3, cudaMemcpyDeviceToHost, st));
  if (h_pos) CUDA_OK(cudaMemcpyAsync(h_pos, d_pos, n3, cudaMemcpyDeviceToHost, st));
  if (h_vel) CUDA_OK(cudaMemcpyAsync(h_vel, d_vel, n3, cudaMemcpyDeviceToHost, st));
  if (h_records) CUDA_OK(cudaMemcpyAsync(h_records, d_rec, nr, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int scp_b200_reconstruct(const double* d_acc, const double* d_p0, const double* d_v0, int B, int N, int K,
                         double h, double* d_pos, double* d_vel, void* stream) {
  if (B <= 0 || N <= 0) return 0;
  if (K < 1) return fail(1, "n_steps must be >= 1");
  const long long BN = (long long)B * N;
  const int wpb = 256 / 32;
  const long long blocks = (BN + wpb - 1) / wpb;
  if (blocks > 0x7fffffffLL) return fail(1, "too many agents for one launch");
  scp_reconstruct_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const double2*)d_acc, (const double2*)d_p0, (const double2*)d_v0, (int)BN, K, h, (double2*)d_pos,
      (double2*)d_vel);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int scp_b200_linearize(const double* d_pos, int B, int N, int K, double R, double feas_margin, double* d_eta,
                       double* d_bound, double* d_minsep, int32_t* d_first, void* stream) {
  return scp_b200_linearize_range(d_pos, B, N, K, R, feas_margin, 0, (int64_t)N * (N - 1) / 2, d_eta, d_bound, d_minsep,
                                  d_first, stream);
}

int scp_b200_linearize_range(const double* d_pos, int B, int N, int K, double R, double feas_margin, int64_t pair_begin,
                             int64_t pair_end, double* d_eta, double* d_bound, double* d_minsep, int32_t* d_first,
                             void* stream) {
  if (B <= 0) return 0;
  if (N < 1 || K < 1) return fail(1, "bad sizes");
  if ((d_eta == nullptr) != (d_bound == nullptr)) return fail(1, "d_eta and d_bound must both be given or both NULL");
  if (pair_begin < 0 || pair_end < pair_begin || pair_end > (int64_t)N * (N - 1) / 2) return fail(1, "bad pair range");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch for the two 64-bit reductions, allocated stream-ordered
  {  // keep the stream-ordered pool's memory between calls (the default threshold returns it at every sync)
    static std::once_flag once;
    std::call_once(once, [] {
      int dev = 0; cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    });
  }
  unsigned long long* red = nullptr;
  CUDA_OK(cudaMallocAsync((void**)&red, 2 * (size_t)B * sizeof(unsigned long long), st));
  CUDA_OK(cudaMemsetAsync(red, 0xFF, 2 * (size_t)B * sizeof(unsigned long long), st));   // +inf / no row, as ordered bit patterns
  if (N >= 2 && pair_end > pair_begin) {
    // position tile of KT steps in shared memory; kept <= 32 KB so that 6-7 CTAs share an SM (streaming kernel)
    int KT = 8;
    while (KT > 1 && (size_t)KT * N * sizeof(double2) > 32 * 1024) KT >>= 1;
    if ((size_t)KT * N * sizeof(double2) > 200 * 1024) { cudaFreeAsync(red, st); return fail(1, "n_agents too large for the position tile"); }
    const int ktiles = (K + KT - 1) / KT;
    const long long P = pair_end - pair_begin;
    long long want = (8LL * 148 + (long long)B * ktiles - 1) / ((long long)B * ktiles);   // >= 8 CTAs per SM in total
    long long maxc = (P + 511) / 512;
    int nchunks = (int)(want < 1 ? 1 : (want > maxc ? maxc : want));
    if (nchunks < 1) nchunks = 1;
    const size_t smem = (size_t)KT * N * sizeof(double2);
    CUDA_OK(cudaFuncSetAttribute(scp_linearize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(nchunks, ktiles, B);
    scp_linearize_kernel<<<grid, LIN_THREADS, smem, st>>>((const double2*)d_pos, N, K, KT, nchunks, (long long)pair_begin,
                                                          (long long)pair_end, R,
                                                          R - feas_margin, (double2*)d_eta, d_bound, red, red + B);
  }
  scp_linearize_finish_kernel<<<(B + 255) / 256, 256, 0, st>>>(red, red + B, B, N, d_minsep, d_first);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(red, st);
  CUDA_OK(e);
  return 0;
}

}  // extern "C"
