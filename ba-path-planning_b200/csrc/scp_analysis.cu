// scp_analysis.cu -- the two steps either side of the SCP solve, on the device:
//
//   scp_generate_kernel   batched rejection-sampling scenario generator.  Same construction and acceptance rules as the
//                         host generators (scenarios/position_generator.py): the reference's layout (reference
//                         position_generator.py:18-40, 44-75: starts on four corner circles, goals on a central diamond
//                         or the circles, pairwise spacing >= min_distance, at most max_attempts draws per set) and the bounded-travel
//                         layout used for more than 50 agents.  The random stream is a counter-based hash (the host
//                         generators use Python's Mersenne twister, which is not worth restating on a GPU): scenarios are
//                         reproducible from (seed, scenario index), not equal to the host generator's draws.
//   scp_check_kernel      post-solve analysis of a batch of trajectories: minimum separation at the samples (the quantity
//                         of scp.py:597-615 and of position_generator.py:173-205's distance report), minimum separation
//                         in CONTINUOUS time (between samples the relative motion is a quadratic in t -- piecewise
//                         constant acceleration, scp.py:371-397 -- so |d(t)|^2 is a quartic whose stationary points are
//                         the roots of a cubic, solved in closed form), and the dynamics residual of SURVEY.md 8(c).
//
// One CTA per scenario in both kernels; every reduction is a warp shuffle + one shared-memory pass, no atomics.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "../../include/scp_b200.h"

int scp_b200_set_error(int code, const char* msg);   // scp_b200.cu

namespace {

#define AN_CUDA_OK(expr)                                                          \
  do {                                                                            \
    cudaError_t e_ = (expr);                                                      \
    if (e_ != cudaSuccess) return scp_b200_set_error(100 + (int)e_, cudaGetErrorString(e_)); \
  } while (0)

constexpr int AN_THREADS = 256;

// ---------------------------------------------------------------------------------- counter-based random numbers
__device__ __forceinline__ uint64_t mix64(uint64_t z) {      // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct Rng {
  uint64_t key, ctr;
  __device__ double uniform() {                               // [0, 1), 53 bits
    return (double)(mix64(key ^ mix64(ctr++)) >> 11) * (1.0 / 9007199254740992.0);
  }
  __device__ double uniform(double a, double b) { return a + (b - a) * uniform(); }
};

// true when candidate (x, y) keeps `gap` to every point placed so far (block-wide test, result to all threads)
__device__ bool spaced(const double2* pts, int n, double x, double y, double gap) {
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dx = pts[i].x - x, dy = pts[i].y - y;
    if (dx * dx + dy * dy < gap * gap) bad = 1;
  }
  return __syncthreads_or(bad) == 0;
}

// layout 0: the reference's (position_generator.py:18-40, 44-75): 20 x 20 m box, starts ON four corner circles (radius 2.5),
//           goals on the border of the central diamond (90 %) or on the circles (10 %), spacing >= min_distance
// layout 1: bounded travel (generate_positions_large): arena side sqrt(16 N), starts uniform with spacing >= 1.25 R, goal =
//           start + U(0.5,1) * 0.4 v_max T in a uniform direction, inside the arena, goal spacing >= 1.25 R
__global__ void __launch_bounds__(AN_THREADS)
scp_generate_kernel(int N, int layout, double min_distance, double time_horizon, double vel_limit, uint64_t seed,
                    int first_scenario, int max_attempts, double2* __restrict__ p0, double2* __restrict__ pf,
                    int* __restrict__ status) {
  const int b = blockIdx.x;
  double2* S = p0 + (size_t)b * N;
  double2* G = pf + (size_t)b * N;
  Rng rng{mix64(seed) ^ mix64((uint64_t)(first_scenario + b) * 0xD1B54A32D192ED03ull + 1), 0};
  int ok = 1;
  if (layout == 0) {
    // reference layout constants (position_generator.py:18-40)
    const double cxs[4] = {3.5, 16.5, 3.5, 16.5}, cys[4] = {3.5, 3.5, 16.5, 16.5};
    const double rad = 2.5, dsz = 6.0 / sqrt(2.0);
    const double vx[4] = {10.0, 10.0 + dsz, 10.0, 10.0 - dsz}, vy[4] = {10.0 + dsz, 10.0, 10.0 - dsz, 10.0};
    for (int phase = 0; phase < 2 && ok; ++phase) {
      double2* dst = phase ? G : S;
      int n = 0;
      for (int a = 0; a < max_attempts && n < N; ++a) {       // max_attempts draws for the whole set (:52-58, :63-72)
        double x, y;
        const bool circle = phase == 0 || !(rng.uniform() < 0.9);
        if (circle) {                                         // ON one of the four corner circles (:235-237)
          const int ci = (int)(rng.uniform() * 4.0) & 3;
          const double th = rng.uniform(0.0, 2.0 * M_PI);
          x = cxs[ci] + rad * cos(th); y = cys[ci] + rad * sin(th);
        } else {                                              // on the border of the central diamond (:240-244)
          const int e = (int)(rng.uniform() * 4.0) & 3;
          const double t = rng.uniform();
          x = vx[e] + t * (vx[(e + 1) & 3] - vx[e]); y = vy[e] + t * (vy[(e + 1) & 3] - vy[e]);
        }
        const bool placed = spaced(dst, n, x, y, min_distance);
        if (placed) { if (threadIdx.x == 0) dst[n] = make_double2(x, y); ++n; }
        __syncthreads();
      }
      if (n < N) ok = 0;
    }
  } else {
    const double side = sqrt(16.0 * (double)N), gap = 1.25 * min_distance, dmax = 0.4 * vel_limit * time_horizon;
    int n = 0;
    for (int a = 0; a < max_attempts && n < N; ++a) {
      const double x = rng.uniform(1.0, side - 1.0), y = rng.uniform(1.0, side - 1.0);
      double gx = 0.0, gy = 0.0;
      bool placed = spaced(S, n, x, y, gap);
      if (placed) {
        placed = false;
        for (int t = 0; t < 20 && !placed; ++t) {
          const double th = rng.uniform(0.0, 2.0 * M_PI), d = rng.uniform(0.5, 1.0) * dmax;
          gx = x + d * cos(th); gy = y + d * sin(th);
          const bool inside = gx >= 1.0 && gx <= side - 1.0 && gy >= 1.0 && gy <= side - 1.0;   // uniform over the block
          placed = inside && spaced(G, n, gx, gy, gap);
        }
      }
      if (placed) {
        if (threadIdx.x == 0) { S[n] = make_double2(x, y); G[n] = make_double2(gx, gy); }
        ++n;
      }
      __syncthreads();
    }
    if (n < N) ok = 0;
  }
  if (threadIdx.x == 0) status[b] = ok;
}

// ---------------------------------------------------------------------------------- closed-form cubic
// min over t in [0, h] of |d + w t + a t^2 / 2|^2 ; returns the squared distance and the minimiser
__device__ double min_quartic(double dx, double dy, double wx, double wy, double ax, double ay, double h, double* tmin) {
  auto f = [&](double t) { const double x = dx + t * (wx + 0.5 * t * ax), y = dy + t * (wy + 0.5 * t * ay); return x * x + y * y; };
  double best = f(0.0), tb = 0.0;
  { const double v = f(h); if (v < best) { best = v; tb = h; } }
  // f'(t)/2 = (d + w t + a t^2/2).(w + a t) = c3 t^3 + c2 t^2 + c1 t + c0
  const double c3 = 0.5 * (ax * ax + ay * ay), c2 = 1.5 * (wx * ax + wy * ay), c1 = wx * wx + wy * wy + dx * ax + dy * ay,
               c0 = dx * wx + dy * wy;
  double r[3]; int nr = 0;
  const double scale = fabs(c3) * h * h * h + fabs(c2) * h * h + fabs(c1) * h + fabs(c0);
  if (fabs(c3) * h * h * h > 1e-14 * scale) {
    const double A = c2 / c3, B = c1 / c3, C = c0 / c3;
    const double p = B - A * A / 3.0, q = 2.0 * A * A * A / 27.0 - A * B / 3.0 + C, sh = -A / 3.0;
    const double disc = 0.25 * q * q + p * p * p / 27.0;
    if (disc > 0.0) {
      const double sq = sqrt(disc);
      r[nr++] = cbrt(-0.5 * q + sq) + cbrt(-0.5 * q - sq) + sh;
    } else {
      const double m = 2.0 * sqrt(-p / 3.0);
      const double arg = m > 0.0 ? fmin(1.0, fmax(-1.0, 3.0 * q / (p * m))) : 0.0;
      const double th = acos(arg) / 3.0;
      for (int k = 0; k < 3; ++k) r[nr++] = m * cos(th - 2.0 * M_PI * k / 3.0) + sh;
    }
  } else if (fabs(c2) * h * h > 1e-14 * scale) {
    const double disc = c1 * c1 - 4.0 * c2 * c0;
    if (disc >= 0.0) { const double sq = sqrt(disc); r[nr++] = (-c1 + sq) / (2.0 * c2); r[nr++] = (-c1 - sq) / (2.0 * c2); }
  } else if (fabs(c1) > 0.0) {
    r[nr++] = -c0 / c1;
  }
  for (int k = 0; k < nr; ++k) {
    double t = r[k];
    if (!(t > 0.0 && t < h)) continue;
    for (int it = 0; it < 2; ++it) {                          // two Newton steps polish the closed-form root
      const double g = ((c3 * t + c2) * t + c1) * t + c0, gp = (3.0 * c3 * t + 2.0 * c2) * t + c1;
      if (gp != 0.0) { const double tn = t - g / gp; if (tn > 0.0 && tn < h) t = tn; }
    }
    const double v = f(t);
    if (v < best) { best = v; tb = t; }
  }
  *tmin = tb;
  return best;
}

struct MinLoc { double v; double where; };
__device__ __forceinline__ MinLoc min_loc(MinLoc a, MinLoc b) { return (b.v < a.v || (b.v == a.v && b.where < a.where)) ? b : a; }

__device__ MinLoc block_min(MinLoc m, MinLoc* sh) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    MinLoc o{__shfl_xor_sync(0xffffffffu, m.v, d), __shfl_xor_sync(0xffffffffu, m.where, d)};
    m = min_loc(m, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = m;
  __syncthreads();
  MinLoc r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = min_loc(r, sh[w]);
  return r;
}
__device__ double block_max(double v, double* sh) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, sh[w]);
  return r;
}

// out[b] = scp_b200_check (8 doubles)
__global__ void __launch_bounds__(AN_THREADS)
scp_check_kernel(int N, int K, double h, double xmin, double ymin, double xmax, double ymax, double vlim, double alim,
                 double jlim, const double2* __restrict__ acc, const double2* __restrict__ pos,
                 const double2* __restrict__ vel, const double2* __restrict__ p0, const double2* __restrict__ v0,
                 const double2* __restrict__ pf, const double2* __restrict__ vf, scp_b200_check* __restrict__ out) {
  __shared__ MinLoc shm[AN_THREADS / 32];
  __shared__ double shd[AN_THREADS / 32];
  const int b = blockIdx.x;
  const double2* A = acc + (size_t)b * N * K;
  const double2* P = pos + (size_t)b * N * K;
  const double2* V = vel + (size_t)b * N * K;
  const long long pairs = (long long)N * (N - 1) / 2;
  // ---- separation: item = (pair, k); sampled distance at k, continuous minimum over [t_k, t_k+1] for k < K-1
  MinLoc ms{INFINITY, 0.0}, mc{INFINITY, 0.0};
  for (long long e = threadIdx.x; e < pairs * K; e += blockDim.x) {
    const long long pr = e / K;
    const int k = (int)(e - pr * K);
    // pair index -> (i, j), i < j, lexicographic (scp.py:487-496 order)
    int i = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * (double)pr)) * 0.5);
    while ((long long)i * (2 * N - i - 1) / 2 > pr) --i;
    while ((long long)(i + 1) * (2 * N - i - 2) / 2 <= pr) ++i;
    const int j = (int)(pr - (long long)i * (2 * N - i - 1) / 2) + i + 1;
    const double2 pi = P[(size_t)i * K + k], pj = P[(size_t)j * K + k];
    const double dx = pi.x - pj.x, dy = pi.y - pj.y;
    const double d2 = dx * dx + dy * dy;
    ms = min_loc(ms, MinLoc{d2, (double)k});
    if (k < K - 1) {
      const double2 vi = V[(size_t)i * K + k], vj = V[(size_t)j * K + k], ai = A[(size_t)i * K + k], aj = A[(size_t)j * K + k];
      double t;
      const double c2 = min_quartic(dx, dy, vi.x - vj.x, vi.y - vj.y, ai.x - aj.x, ai.y - aj.y, h, &t);
      mc = min_loc(mc, MinLoc{c2, (double)k + t / h});
    } else {
      mc = min_loc(mc, MinLoc{d2, (double)k});
    }
  }
  ms = block_min(ms, shm);
  mc = block_min(mc, shm);
  // ---- dynamics residual: item = (agent, k)
  double box = 0.0, dyn = 0.0, term = 0.0;
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    const int i = e / K, k = e - i * K;
    const double2 a = A[e], p = P[e], v = V[e];
    box = fmax(box, fmax(fabs(a.x), fabs(a.y)) - alim);
    if (k == 0) {
      dyn = fmax(dyn, fmax(fmax(fabs(p.x - p0[(size_t)b * N + i].x), fabs(p.y - p0[(size_t)b * N + i].y)),
                           fmax(fabs(v.x - v0[(size_t)b * N + i].x), fabs(v.y - v0[(size_t)b * N + i].y))));
    } else {                                                   // box rows bind states 1..K-1 (scp.py:212-257)
      box = fmax(box, fmax(fabs(v.x), fabs(v.y)) - vlim);
      box = fmax(box, fmax(fmax(xmin - p.x, p.x - xmax), fmax(ymin - p.y, p.y - ymax)));
    }
    // state k+1 from state k under constant acceleration (scp.py:371-397)
    const double nvx = v.x + h * a.x, nvy = v.y + h * a.y;
    const double npx = p.x + h * v.x + 0.5 * h * h * a.x, npy = p.y + h * v.y + 0.5 * h * h * a.y;
    if (k < K - 1) {
      const double2 a1 = A[e + 1], p1 = P[e + 1], v1 = V[e + 1];
      box = fmax(box, fmax(fabs(a1.x - a.x), fabs(a1.y - a.y)) / h - jlim);
      dyn = fmax(dyn, fmax(fmax(fabs(p1.x - npx), fabs(p1.y - npy)), fmax(fabs(v1.x - nvx), fabs(v1.y - nvy))));
    } else {                                                   // terminal equalities bind state K (scp.py:219-224, 250-257)
      term = fmax(term, fmax(fmax(fabs(npx - pf[(size_t)b * N + i].x), fabs(npy - pf[(size_t)b * N + i].y)),
                             fmax(fabs(nvx - vf[(size_t)b * N + i].x), fabs(nvy - vf[(size_t)b * N + i].y))));
    }
  }
  box = block_max(fmax(box, 0.0), shd);
  dyn = block_max(dyn, shd);
  term = block_max(term, shd);
  if (threadIdx.x == 0) {
    scp_b200_check r;
    r.min_separation = pairs ? sqrt(ms.v) : INFINITY;
    r.min_separation_step = ms.where;
    r.min_separation_continuous = pairs ? sqrt(mc.v) : INFINITY;
    r.min_separation_continuous_time = mc.where * h;
    r.box_violation = box;
    r.dynamics_violation = dyn;
    r.terminal_violation = term;
    r.dynamics_residual = fmax(box, fmax(dyn, term));
    out[b] = r;
  }
}

}  // namespace

extern "C" {

int scp_b200_generate_scenarios(int n_scenarios, int n_agents, int layout, double min_distance, double time_horizon,
                                double vel_limit, uint64_t seed, int first_scenario, int max_attempts, double* d_p0,
                                double* d_pf, int32_t* d_status, void* stream) {
  if (n_scenarios < 0 || n_agents < 1 || (layout != 0 && layout != 1) || !d_p0 || !d_pf || !d_status)
    return scp_b200_set_error(1, "scp_b200_generate_scenarios: bad arguments");
  if (n_scenarios == 0) return 0;
  if (max_attempts <= 0) max_attempts = layout == 0 ? 1000 : 200 * n_agents;
  scp_generate_kernel<<<n_scenarios, AN_THREADS, 0, (cudaStream_t)stream>>>(
      n_agents, layout, min_distance, time_horizon, vel_limit, seed, first_scenario, max_attempts, (double2*)d_p0,
      (double2*)d_pf, d_status);
  AN_CUDA_OK(cudaGetLastError());
  return 0;
}

int scp_b200_check_batch(const scp_b200_problem* prob, int n_scenarios, const double* d_acc, const double* d_pos,
                         const double* d_vel, const double* d_p0, const double* d_v0, const double* d_pf,
                         const double* d_vf, scp_b200_check* d_out, void* stream) {
  if (!prob || n_scenarios < 0 || !d_acc || !d_pos || !d_vel || !d_p0 || !d_v0 || !d_pf || !d_vf || !d_out)
    return scp_b200_set_error(1, "scp_b200_check_batch: bad arguments");
  if (n_scenarios == 0) return 0;
  scp_check_kernel<<<n_scenarios, AN_THREADS, 0, (cudaStream_t)stream>>>(
      prob->n_agents, prob->n_steps, prob->time_step, prob->space[0], prob->space[1], prob->space[2], prob->space[3],
      prob->vel_limit, prob->acc_limit, prob->jerk_limit, (const double2*)d_acc, (const double2*)d_pos,
      (const double2*)d_vel, (const double2*)d_p0, (const double2*)d_v0, (const double2*)d_pf, (const double2*)d_vf, d_out);
  AN_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
