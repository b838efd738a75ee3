// scp_tables.h -- host-side construction of the constant operator tables.
//
// Replaces the constant part of SCP._precompute_constraint_matrices
// (reference src/path_planning/solvers/scp.py:182-232): instead of the sparse
// C_jerk / C_acc / C_vel / C_pos matrices the solver needs, per (K, h, weights),
//   B1 = D'RjD + Ra + V'RvV + S'RpS      (K x K, box rows, unit rho)
//   B2 = S'RcS                           (K x K, one collision copy, unit rho)
// with D the first difference /h (scp.py:10-28), V = h T (scp.py:198-203),
// S[k][j] = h^2 (k-j+1/2) (scp.py:227-232), rows k = 0..K-2 (row K-1 of V and S
// is the terminal equality, scp.py:222-224 and 255-257, handled exactly).
// Row weights follow an inf-norm row equilibration (what OSQP's Ruiz scaling
// converges to for these rows): r = w_class / max|row|.
#ifndef SCP_TABLES_H
#define SCP_TABLES_H

#include <vector>

#include "../../include/scp_b200.h"

namespace scp {

struct HostTables {
  int K;
  std::vector<double> blob;  // [B1 K*K | B2 K*K | rj K | ra K | rv K | rp K | rc K]
  size_t oB1, oB2, orj, ora, orv, orp, orc;
};

inline size_t tables_doubles(int K) { return 2 * (size_t)K * K + 5 * (size_t)K; }

inline HostTables build_host_tables(const scp_b200_problem& pb) {
  HostTables t;
  const int K = pb.n_steps;
  const double h = pb.time_step;
  t.K = K;
  t.blob.assign(tables_doubles(K), 0.0);
  t.oB1 = 0; t.oB2 = (size_t)K * K; t.orj = 2 * (size_t)K * K;
  t.ora = t.orj + K; t.orv = t.ora + K; t.orp = t.orv + K; t.orc = t.orp + K;
  double* B1 = t.blob.data() + t.oB1;
  double* B2 = t.blob.data() + t.oB2;
  double *rj = t.blob.data() + t.orj, *ra = t.blob.data() + t.ora, *rv = t.blob.data() + t.orv;
  double *rp = t.blob.data() + t.orp, *rc = t.blob.data() + t.orc;
  for (int k = 0; k < K; ++k) {
    rj[k] = pb.w_jerk * h;                       // max|D row| = 1/h
    ra[k] = pb.w_acc;
    rv[k] = pb.w_vel / h;                        // max|V row| = h
    rp[k] = pb.w_pos / (h * h * (k + 0.5));      // max|S row k| = h^2 (k+1/2)
    rc[k] = pb.w_col / (h * h * (k + 0.5));
  }
  // suffix sums over rows k = m..K-2 of rv, and of rp/rc times 1, k, k^2
  // S'RS[m][n] = h^4 sum_k r[k] (k-m+.5)(k-n+.5) = h^4 (s2 - (m+n-1) s1 + (m-.5)(n-.5) s0), k >= max(m,n)
  std::vector<double> sv(K + 1, 0.0), p0(K + 1, 0.0), p1(K + 1, 0.0), p2(K + 1, 0.0), c0(K + 1, 0.0),
      c1(K + 1, 0.0), c2(K + 1, 0.0);
  for (int k = K - 2; k >= 0; --k) {
    sv[k] = sv[k + 1] + rv[k];
    p0[k] = p0[k + 1] + rp[k]; p1[k] = p1[k + 1] + rp[k] * k; p2[k] = p2[k + 1] + rp[k] * (double)k * k;
    c0[k] = c0[k + 1] + rc[k]; c1[k] = c1[k + 1] + rc[k] * k; c2[k] = c2[k + 1] + rc[k] * (double)k * k;
  }
  const double h2 = h * h, h4 = h2 * h2;
  for (int m = 0; m < K; ++m)
    for (int n = 0; n < K; ++n) {
      int a = m > n ? m : n;
      double am = m - 0.5, an = n - 0.5;
      double v = h2 * sv[a] + h4 * (p2[a] - (am + an) * p1[a] + am * an * p0[a]);
      if (m == n) {
        v += ra[m];
        if (m >= 1) v += rj[m - 1] / h2;
        if (m < K - 1) v += rj[m] / h2;
      } else if (n == m + 1) {
        v -= rj[m] / h2;
      } else if (m == n + 1) {
        v -= rj[n] / h2;
      }
      B1[(size_t)m * K + n] = v;
      B2[(size_t)m * K + n] = h4 * (c2[a] - (am + an) * c1[a] + am * an * c0[a]);
    }
  return t;
}

}  // namespace scp
#endif
